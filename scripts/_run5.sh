timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_seeding.py tests/test_gpu_fullsize.py tests/test_gpu_property.py tests/test_gpu_where.py -x -q 2>&1 | tail -3
for nd in 0 1; do
for cfg in "1000000 1 5 200 0 384" "1000000 64 5 200 0 384" "1000000 1 10 200 0 512" "1000000 64 20 100 0 768" "1000000 128 5 200 0 384"; do
echo "== NO_DYN=$nd $cfg"
B2R_NO_DYN=$nd B2R_TRACE=1 timeout 200 python scripts/pool_large.py $cfg 2>&1 | grep -v "^built" | sed 's/waits.*ctas/ctas/'
done; done
