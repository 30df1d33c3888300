bash scripts/gpu_validate.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/r2q_ncu_bench.log 2>&1; echo "ncu rc=$?"
