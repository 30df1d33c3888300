"""Device timing of K1 (fused normalise + bf16 pack ingest), device-resident input (development aid).

    python scripts/quick_ingest.py [D BATCH_ROWS N_BATCHES]
"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_rag_b200 import B200Collection, _lib


def run(d, rows, nb, space="cosine", master=True):
    lib = _lib.load()
    c = B200Collection("i", {"hnsw:space": space}, capacity=rows * (nb + 2), dimension=d, keep_f32_master=master)
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(rows, d, generator=g, device="cuda")
    first = ctypes.c_int64()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        _lib.check(lib.b2r_ingest_f32(c.handle, x.data_ptr(), rows, None, ctypes.byref(first), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nb):
        _lib.check(lib.b2r_ingest_f32(c.handle, x.data_ptr(), rows, None, ctypes.byref(first), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nb
    dp = c.stats()["dim_padded"]
    byt = rows * (d * 4 + dp * 2 + (dp * 4 if master else 0) + 1 + (4 if space == "l2" else 0))
    print(f"ingest {rows} x {d} {space} master={master}: {ms * 1e3:.1f} us/batch = {rows / ms / 1e3:.1f} M rows/s, "
          f"{byt / ms / 1e6:.0f} GB/s algorithmic ({byt / rows} B/row)", flush=True)
    c.close()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        a = [int(v) for v in sys.argv[1:]] + [None] * 3
        run(a[0], a[1] or 262144, a[2] or 10)
    else:
        run(384, 262144, 10)
        run(384, 8192, 50)
        run(768, 262144, 6)
        run(768, 8192, 50)
        run(512, 262144, 8, "l2")
        run(384, 262144, 10, "cosine", False)
