"""Pool mode (32 < k <= 128) on large shards: time, fall-backs, pool size, per-CTA phase skew, parity against the exact
fp64 path on a sample of the batch.  Development aid (BASELINE config 4 per-GPU shapes: 12.5M / 25M / 50M rows).

    python scripts/pool_large.py ROWS [NQ K ITERS CHECK DIM]      env: B2R_TRACE=1 B2R_DELAY_US=.. B2R_SEED_WAIT_NS=.. B2R_NO_SEED=1
"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multimodal_rag_b200 import _lib
from multimodal_rag_b200.sharded import DeviceShard


def build(rows, dim, seed=0xC4):
    sh = DeviceShard(dim, "cosine", capacity=rows, row_base=0, device=0)
    g = torch.Generator(device="cuda").manual_seed(seed)
    for s0 in range(0, rows, 1 << 18):
        m = min(1 << 18, rows - s0)
        sh.ingest(torch.nn.functional.normalize(torch.randn(m, dim, generator=g, device="cuda"), dim=1))
    torch.cuda.synchronize()
    return sh


def trace_summary(lib, sh):
    buf = np.zeros((256, 8), dtype=np.uint64)
    n = ctypes.c_int()
    _lib.check(lib.b2r_debug_trace(sh.h, buf.ctypes.data, 256, ctypes.byref(n)))
    if n.value == 0:
        return None
    t = buf[: n.value].astype(np.int64)
    t0 = t[:, 0].min()
    rel = (t[:, :4] - t0) / 1e3
    wait = t[:, 4:]
    waits = {"epi_wait_full_acc_kcyc": float(np.median(wait[:, 2]) / 1e3)}
    if os.environ.get("B2R_TRACE") == "2":
        ft = (t[:, 4] - t0) / 1e3; sd = (t[:, 5] - t0) / 1e3
        ft = ft[t[:, 4] > 0]; sd = sd[t[:, 5] > 0]
        waits = {"first_full_acc_us": [float(ft.min()), float(np.median(ft)), float(ft.max())],
                 "sampling_done_us": [float(sd.min()), float(np.median(sd)), float(sd.max())]}
    if os.environ.get("B2R_TRACE") == "3":
        def rel_(c):
            v = (t[:, c] - t0) / 1e3
            v = v[t[:, c] > 0]
            return [float(v.min()), float(np.median(v)), float(v.max())] if v.size else None
        waits = {"mma_past_launch_wait_us": rel_(4), "queries_resident_us": rel_(5), "first_full_acc_us": rel_(7)}
    return {"waits": waits, "ctas": n.value, "start_spread_us": float(rel[:, 0].max()), "posted_us": [float(rel[:, 1].min()), float(rel[:, 1].max())],
            "seeded_us": [float(rel[:, 2].min()), float(rel[:, 2].max())], "done_us": [float(rel[:, 3].min()), float(np.median(rel[:, 3])), float(rel[:, 3].max())]}


def main():
    a = [int(x) for x in sys.argv[1:]] + [None] * 6
    rows, nq, k, iters, check = a[0] or 25_000_000, a[1] or 1024, a[2] or 100, a[3] or 3, a[4] if a[4] is not None else 8
    dim = a[5] or 384
    lib = _lib.load()
    t0 = time.time()
    sh = build(rows, dim)
    print(f"built {rows} x {dim} in {time.time() - t0:.1f} s", flush=True)
    g = torch.Generator(device="cuda").manual_seed(0xBEEF4)
    Q = [torch.nn.functional.normalize(torch.randn(nq, dim, generator=g, device="cuda"), dim=1) for _ in range(2)]
    o = sh.alloc_out(nq, k)
    for i in range(2):
        sh.query_local(Q[i % 2], k, o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        sh.query_local(Q[i % 2], k, o)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    st = _lib.B2RStats()
    _lib.check(lib.b2r_get_stats(sh.h, ctypes.byref(st)))
    print(f"rows={rows} nq={nq} k={k}: {ms:.3f} ms/batch = {nq / ms * 1e3:.0f} q/s, {2.0 * nq * rows * dim / ms / 1e9:.0f} TFLOP/s, "
          f"fallbacks={st.n_exact_fallbacks} pool/query={st.n_pool_entries / max(1, st.n_pool_queries):.0f}", flush=True)
    tr = trace_summary(lib, sh)
    if tr:
        print("trace:", tr, flush=True)
    if check:
        # the fast path against the exact fp64 scan (K5) on the first `check` queries of batch 0
        sh.query_local(Q[0], k, o)
        fast_rows = o["rows"][:check].clone(); fast_d = o["d64"][:check].clone()
        _lib.check(lib.b2r_set_path(sh.h, 3))
        o2 = sh.alloc_out(check, k)
        t0 = time.time()
        sh.query_local(Q[0][:check].contiguous(), k, o2)
        torch.cuda.synchronize()
        print(f"exact path: {check} queries in {(time.time() - t0) * 1e3:.1f} ms", flush=True)
        _lib.check(lib.b2r_set_path(sh.h, 0))
        same = torch.equal(fast_rows, o2["rows"])
        print("parity vs exact path:", "OK" if same else "MISMATCH", "max |d64 diff| =", float((fast_d - o2["d64"]).abs().max()), flush=True)
        if not same:
            sys.exit(1)
    sh.close()


if __name__ == "__main__":
    main()
