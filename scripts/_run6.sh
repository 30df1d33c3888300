timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_pool_large.py tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2e_pytest.log
