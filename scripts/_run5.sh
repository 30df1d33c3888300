timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_seeding.py tests/test_gpu_fullsize.py tests/test_gpu_pool_large.py -x -q 2>&1 | tail -3
for cfg in "1000000 256 5 200 0 384" "1000000 1 5 200 0 384" "1000000 1 10 200 0 512" "1000000 256 10 100 0 512" "1000000 1024 100 50 0 384"; do
echo "== $cfg"
B2R_TRACE=2 timeout 200 python scripts/pool_large.py $cfg 2>&1 | grep -v "^built"
done
