timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_seeding.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
for rep in 1 2; do
for lib in scripts/_ab/libb2r_before.so multimodal_rag_b200/libb2r.so; do
echo "== $lib"
B2R_LIB=$PWD/$lib timeout 200 python scripts/pool_large.py 1000000 256 5 400 0 384 2>&1 | grep "^rows"
B2R_LIB=$PWD/$lib timeout 200 python scripts/pool_large.py 1000000 1024 5 150 0 384 2>&1 | grep "^rows"
B2R_LIB=$PWD/$lib timeout 200 python scripts/pool_large.py 1000000 256 10 200 0 512 2>&1 | grep "^rows"
done; done
