set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2a_pytest.log
B2R_TRACE=1 timeout 300 python scripts/pool_large.py 4000000 1024 100 5 8 > gpurun_out/r2a_pool4m.log 2>&1; echo rc=$?
B2R_TRACE=1 B2R_DELAY_US=200 timeout 300 python scripts/pool_large.py 4000000 1024 100 5 8 > gpurun_out/r2a_pool4m_delay.log 2>&1; echo rc=$?
B2R_TRACE=1 B2R_SEED_WAIT_NS=1 timeout 300 python scripts/pool_large.py 4000000 1024 100 5 8 > gpurun_out/r2a_pool4m_nowait.log 2>&1; echo rc=$?
B2R_TRACE=1 B2R_NO_SEED=1 timeout 300 python scripts/pool_large.py 4000000 1024 100 5 8 > gpurun_out/r2a_pool4m_noseed.log 2>&1; echo rc=$?
B2R_TRACE=1 B2R_NO_SEED=1 timeout 300 python scripts/pool_large.py 1000000 64 100 5 8 > gpurun_out/r2a_pool1m_b64_noseed.log 2>&1; echo rc=$?
B2R_TRACE=1 timeout 600 python scripts/pool_large.py 25000000 1024 100 3 8 > gpurun_out/r2a_pool25m.log 2>&1; echo rc=$?
B2R_TRACE=1 timeout 900 python scripts/pool_large.py 50000000 1024 100 3 4 > gpurun_out/r2a_pool50m.log 2>&1; echo rc=$?
tail -n 6 gpurun_out/r2a_pool*.log
