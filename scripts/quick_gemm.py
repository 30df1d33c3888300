"""Ad-hoc device timing of the K3 path (development aid, not the bench)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_rag_b200 import B200Collection, _lib

def run(n, d, nq, k, space="cosine", iters=20, path=0):
    g = torch.Generator(device="cuda").manual_seed(1)
    c = B200Collection("b", {"hnsw:space": space}, capacity=n, dimension=d)
    lib = _lib.load()
    step = 1 << 18
    for s in range(0, n, step):
        m = min(step, n - s)
        x = torch.randn(m, d, generator=g, device="cuda")
        first = ctypes.c_int64()
        _lib.check(lib.b2r_ingest_f32(c.handle, x.data_ptr(), m, None, ctypes.byref(first), 0))
    q = torch.nn.functional.normalize(torch.randn(nq, d, generator=g, device="cuda"), dim=1)
    rows = torch.empty(nq, k, dtype=torch.int64, device="cuda"); dist = torch.empty(nq, k, device="cuda")
    cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
    if path: c.set_path(path)
    st = torch.cuda.current_stream().cuda_stream
    def once():
        _lib.check(lib.b2r_query(c.handle, q.data_ptr(), nq, k, None, rows.data_ptr(), dist.data_ptr(), cnt.data_ptr(), st))
    for _ in range(3): once()
    torch.cuda.synchronize()
    tot, cn = ctypes.c_double(), ctypes.c_int64()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): once()                  # whole-call time: no events between the kernels (they would break PDL)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    lib.b2r_set_kernel_timing(c.handle, 1)
    for _ in range(iters): once()                  # second loop: CUDA events around the scoring kernel
    torch.cuda.synchronize()
    lib.b2r_kernel_time_ms(c.handle, ctypes.byref(tot), ctypes.byref(cn), 1)
    lib.b2r_set_kernel_timing(c.handle, 0)
    kms = tot.value / max(1, cn.value)
    gb = n * c.stats()["dim_padded"] * 2 / 1e9
    fl = 2.0 * nq * n * d
    print(f"n={n} d={d} nq={nq} k={k} {space} path={path}: {ms*1e3:.1f} us/batch qps={nq/ms*1e3:.0f} | scoring kernel {kms*1e3:.1f} us x{cn.value/iters:.0f}"
          f" -> {gb/kms*1e3:.0f} GB/s {fl/kms/1e9:.0f} TFLOP/s | fallbacks={c.stats()['n_exact_fallbacks']} pool/query={c.stats()['n_pool_entries']/max(1,c.stats()['n_pool_queries']):.0f}", flush=True)
    c.close()

if __name__ == "__main__" and len(sys.argv) > 1:
    # python scripts/quick_gemm.py NQ [D K ITERS N PATH]: one configuration (used under ncu)
    a = [int(x) for x in sys.argv[1:]] + [None] * 6
    run(a[4] or 1_000_000, a[1] or 384, a[0], a[2] or 5, iters=a[3] or 5, path=a[5] or 0)
elif __name__ == "__main__":
    run(1_000_000, 384, 256, 5)
    run(1_000_000, 384, 128, 5)
    run(1_000_000, 384, 64, 5)
    run(1_000_000, 384, 16, 5)
    run(1_000_000, 384, 1024, 5)
    run(1_000_000, 512, 256, 10)
    run(1_000_000, 768, 64, 20)
    run(1_000_000, 384, 1024, 100)
    run(1_000_000, 384, 256, 100)
    run(1_000_000, 384, 1, 100)
    run(4_000_000, 384, 1024, 100, iters=5)
    run(1_000_000, 384, 1, 5)
    run(10_000, 384, 1, 5)
    for nq in (1, 2, 4, 8, 16):
        run(1_000_000, 384, nq, 5, path=2, iters=50)
    run(10_000, 384, 1, 5, path=2, iters=50)
    run(1_000_000, 512, 1, 10, path=2, iters=50)
    run(1_000_000, 768, 1, 20, path=2, iters=50)
    run(1_000_000, 512, 1, 10, path=1, iters=50)
    run(1_000_000, 768, 1, 20, path=1, iters=50)
