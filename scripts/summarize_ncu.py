"""Turn ncu reports / launch lists under gpurun_out/ into the small text summaries kept under profiles/.

    python scripts/summarize_ncu.py raw  gpurun_out/X.ncu-rep  > profiles/X.md
    python scripts/summarize_ncu.py list gpurun_out/launches.csv [N_LAST] > profiles/launches.md
"""
import csv, io, re, subprocess, sys, collections

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of `{path}`\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## {name[:110]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k or h.endswith("." + k):
                    print(f"| {k} | {r[i]} | {units[i]} |")
                    break
        print()


def launches(path, n_last=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, gi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
    seq = []
    for r in rows[hi + 1:]:
        if len(r) > vi:
            seq.append((re.sub(r"\(.*", "", r[ki].replace("void ", "")).strip()[:64], r[gi], float(r[vi]) / 1e3))
    if n_last:
        seq = seq[-int(n_last):]
    agg = collections.OrderedDict()
    for n, g, t in seq:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1; a[1] += t
    tot = sum(a[1] for a in agg.values())
    print(f"# per-launch device times (`ncu --metrics gpu__time_duration.sum --clock-control none`) from `{path}`\n")
    print("Cold-cache, serialised launches: compare SHARES, not absolutes.\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    print("\n## last launches in order\n\n| kernel | grid | us |\n|---|---|---|")
    for n, g, t in seq[-24:]:
        print(f"| {n} | {g} | {t:.2f} |")


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2])
    else:
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
