set -x
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_seeding.py tests/test_gpu_configs.py -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2d_pytest.log
cat > /tmp/q.py <<'PY'
import sys; sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
from quick_gemm import run
run(1_000_000, 384, 256, 5, iters=200)
run(1_000_000, 384, 1024, 5, iters=50)
run(1_000_000, 512, 256, 10, iters=100)
run(1_000_000, 768, 256, 20, iters=50)
run(1_000_000, 384, 1024, 100, iters=50)
run(4_000_000, 384, 1024, 100, iters=20)
PY
timeout 300 python /tmp/q.py > gpurun_out/r2d_pair.log 2>&1; echo rc=$?
B2R_NO_PAIR=1 timeout 300 python /tmp/q.py > gpurun_out/r2d_nopair.log 2>&1; echo rc=$?
cat gpurun_out/r2d_pair.log gpurun_out/r2d_nopair.log
