"""Per-source-line stall samples of one kernel from an ncu report (development aid).

    python scripts/ncu_hot_lines.py REPORT.ncu-rep KERNEL_REGEX [TOP_N]
"""
import csv, io, subprocess, sys, collections

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if not hi:
    hi = [i for i, r in enumerate(rows) if "# Samples" in r]
h = rows[hi[0]]
end = hi[1] - 1 if len(hi) > 1 else len(rows)
si = h.index("# Samples")
srci = h.index("Source")
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.OrderedDict()
tot = 0
for r in rows[hi[0] + 1:end]:
    if len(r) <= si:
        continue
    try:
        n = int(r[si])
    except ValueError:
        continue
    tot += n
    key = r[srci][:110]
    a = agg.setdefault(key, [0, collections.Counter()])
    a[0] += n
    for i, nm in stall_cols:
        try:
            a[1][nm] += int(r[i])
        except ValueError:
            pass
print(f"total samples {tot}")
for k, (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    s = ", ".join(f"{a[6:]}={b}" for a, b in st.most_common(3) if b)
    print(f"{n:6d} {100*n/max(1,tot):5.1f}%  {k}   [{s}]")
