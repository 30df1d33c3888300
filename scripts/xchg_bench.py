"""A/B timing of the cross-shard exchange variants on N GPUs (torchrun): local scan only, NCCL all_gather + merge, the library's own
exchange kernel -- each on one stream (latency of a batch) and pipelined on a side stream (throughput).  Variants are
interleaved round-robin so that clock drift under the power cap hits all of them alike."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from multimodal_rag_b200.sharded import DeviceShard

def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    n, d, nq, k = 1_000_000, 384, int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 5
    F = torch.nn.functional
    def mk():
        sh = DeviceShard(d, "cosine", capacity=n, row_base=rank * n, device=dev.index)
        g = torch.Generator(device=dev).manual_seed(1 + rank)
        for s in range(0, n, 1 << 18):
            sh.ingest(F.normalize(torch.randn(min(1 << 18, n - s), d, generator=g, device=dev), dim=1))
        return sh
    sh_n, sh_p = mk(), mk()                # same rows; one keeps NCCL, the other gets the p2p exchange
    sh_p.enable_p2p_exchange(nq_max=max(nq, 256), k_max=max(k, 8))
    gq = torch.Generator(device=dev).manual_seed(7)
    Q = [F.normalize(torch.randn(nq, d, generator=gq, device=dev), dim=1) for _ in range(4)]
    on, op = sh_n.alloc_out(nq, k), sh_p.alloc_out(nq, k)
    on2, op2 = [sh_n.alloc_out(nq, k) for _ in range(2)], [sh_p.alloc_out(nq, k) for _ in range(2)]
    K = 100
    def v_local(i): sh_n.query_local(Q[i % 4], k, on)
    def v_nccl(i): sh_n.query_device(Q[i % 4], k, on)
    def v_p2p(i): sh_p.query_device(Q[i % 4], k, op)
    def v_nccl_pipe(i):
        sh_n.query_device_pipelined(Q[i % 4], k, on2[i % 2])
        if i == K - 1: sh_n.drain()
    def v_p2p_pipe(i):
        sh_p.query_device_pipelined(Q[i % 4], k, op2[i % 2])
        if i == K - 1: sh_p.drain()
    sh_f = mk()                            # a third shard with mailboxes of its own for the fused form (b2r_query_push)
    sh_f.enable_p2p_exchange(nq_max=max(nq, 256), k_max=max(k, 8), default=False)
    of2 = [sh_f.alloc_out(nq, k) for _ in range(2)]
    def v_fused(i):
        sh_f.query_device_fused(Q[i % 4], k, of2[i % 2])
        if i == K - 1: sh_f.drain()
    variants = [("local scan only", v_local), ("fused into the query's kernels (b2r_query_push), merge rides in the next call", v_fused), ("nccl all_gather + merge, one stream", v_nccl), ("b2r_xchg_merge, one stream", v_p2p),
                ("nccl, side stream (pipelined)", v_nccl_pipe), ("b2r_xchg_merge, side stream (pipelined)", v_p2p_pipe)]
    times = {name: [] for name, _ in variants}
    for rnd in range(8):
        for name, fn in variants:
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K): fn(i)
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rnd >= 2: times[name].append(float(t.item()))
    if rank == 0:
        base = sorted(times["local scan only"])[len(times["local scan only"]) // 2]
        for name, _ in variants:
            ts = sorted(times[name]); med = ts[len(ts) // 2]
            print(f"N={world} nq={nq} k={k}: {name:80s} median {med * 1e3:7.1f} us/step  min {ts[0] * 1e3:7.1f}  (+{(med - base) * 1e3:5.1f} us over the local scan)", flush=True)
    sh_n.close(); sh_p.close(); sh_f.close()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
