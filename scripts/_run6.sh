mkdir -p gpurun_out
./scripts/dev/sys_scope_ubench 2>&1 | tee gpurun_out/sys_scope_ubench.log
python scripts/fused_cost.py 256 2>&1 | tee gpurun_out/fused_cost2.log
python scripts/fused_cost.py 1 2>&1 | tee -a gpurun_out/fused_cost2.log
