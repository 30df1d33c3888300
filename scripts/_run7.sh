mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
python scripts/fused_cost.py 256 2>&1 | tee gpurun_out/fused_cost3.log
python scripts/fused_cost.py 1 2>&1 | tee -a gpurun_out/fused_cost3.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_seeding.py tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -3
