mkdir -p gpurun_out
python scripts/fused_cost.py 256 2>&1 | tee gpurun_out/fused_cost.log
python scripts/fused_cost.py 1 2>&1 | tee -a gpurun_out/fused_cost.log
