"""What the fused exchange costs on ONE GPU (a world of one rank: the mailbox is local, no NVLink): plain b2r_query_ex against
b2r_query_push with the merge enqueued behind the next batch, and against b2r_query_push with the merges bunched four at a time
(which separates the price of the push + publish from the price of the merge launch).  Round-robin, CUDA events."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_rag_b200 import _lib
from multimodal_rag_b200.sharded import DeviceShard

lib = _lib.load()
n, d, nq, k = 1_000_000, 384, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 5
dev = torch.device("cuda", 0)
F = torch.nn.functional
sh = DeviceShard(d, "cosine", capacity=n, row_base=0, device=0, world=1)
g = torch.Generator(device=dev).manual_seed(1)
for s in range(0, n, 1 << 18):
    sh.ingest(F.normalize(torch.randn(min(1 << 18, n - s), d, generator=g, device=dev), dim=1))
Q = [F.normalize(torch.randn(nq, d, generator=g, device=dev), dim=1) for _ in range(4)]
x = ctypes.c_void_p()
_lib.check(lib.b2r_xchg_create(0, 0, 1, nq, 8, ctypes.byref(x)))
outs = [sh.alloc_out(nq, k) for _ in range(4)]
st = torch.cuda.current_stream().cuda_stream
K = 100

def push(i, ride=None):
    o = outs[i % 4]
    m = outs[ride % 4] if ride is not None else None
    _lib.check(lib.b2r_query_push(sh.h, x, Q[i % 4].data_ptr(), nq, k, None, o["rows"].data_ptr(), o["dist"].data_ptr(), o["cnt"].data_ptr(),
                                  m["m_rows"].data_ptr() if m else None, m["m_dist"].data_ptr() if m else None, m["m_cnt"].data_ptr() if m else None, st))
def merge(i):
    o = outs[i % 4]
    _lib.check(lib.b2r_xchg_merge(x, nq, k, o["m_rows"].data_ptr(), o["m_dist"].data_ptr(), o["m_cnt"].data_ptr(), st))
def v_plain():
    for i in range(K): sh.query_local(Q[i % 4], k, outs[i % 4])
def v_fused():
    for i in range(K):
        push(i)
        if i: merge(i - 1)
    merge(K - 1)
def v_rider():
    for i in range(K): push(i, i - 1 if i else None)
    merge(K - 1)
def v_bunched():
    for i in range(K):
        push(i)
        if i % 4 == 3:
            for j in range(i - 3, i + 1): merge(j)
variants = [("plain b2r_query_ex", v_plain), ("b2r_query_push, merge rides in the next batch's last kernel", v_rider), ("b2r_query_push, merge behind the next batch", v_fused), ("b2r_query_push, merges four at a time", v_bunched)]
times = {nme: [] for nme, _ in variants}
for rnd in range(8):
    for nme, fn in variants:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if rnd >= 2: times[nme].append(e0.elapsed_time(e1) / K)
base = sorted(times["plain b2r_query_ex"])[3]
for nme, _ in variants:
    ts = sorted(times[nme])
    print(f"nq={nq}: {nme:62s} median {ts[len(ts)//2]*1e3:7.1f} us/step  min {ts[0]*1e3:7.1f}  (+{(ts[len(ts)//2]-base)*1e3:5.1f} us)", flush=True)
