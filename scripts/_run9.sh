timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/xchg_bench.py 256 5 2>&1 | grep "^N=\|rror"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/xchg_bench.py 1024 100 2>&1 | grep "^N=\|rror"
