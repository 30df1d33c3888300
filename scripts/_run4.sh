mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/xchg_bench.py 256 5 2>/dev/null | tee gpurun_out/xchg_ab2.log
