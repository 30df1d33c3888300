timeout 600 python -m pytest tests/test_gpu_seeding.py tests/test_gpu_gemm.py tests/test_gpu_fullsize.py tests/test_gpu_property.py -m gpu -x -q 2>&1 | tail -3
for r in 1 2; do python scripts/quick_gemm.py 256 384 5 1000 2>&1 | tail -1 | cut -c1-140; B2R_SEED_TILES=2 python scripts/quick_gemm.py 256 384 5 1000 2>&1 | tail -1 | cut -c1-140; done
