timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_property.py tests/test_gpu_pool_large.py -x -q 2>&1 | tail -3
timeout 300 python scripts/pool_large.py 25000000 1024 100 2 8 2>&1 | grep "exact\|parity"
timeout 300 python scripts/pool_large.py 4000000 256 5 2 8 2>&1 | grep "exact\|parity"
