timeout 600 python -m pytest tests/test_gpu_seeding.py tests/test_gpu_gemm.py tests/test_gpu_fullsize.py tests/test_gpu_property.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
for r in 1 2; do
python scripts/quick_gemm.py 256 384 5 1000 2>&1 | tail -1 | cut -c1-175; B2R_SEED_RANK_L=1 python scripts/quick_gemm.py 256 384 5 1000 2>&1 | tail -1 | cut -c1-175
python scripts/quick_gemm.py 256 512 10 600 2>&1 | tail -1 | cut -c1-175; B2R_SEED_RANK_L=1 python scripts/quick_gemm.py 256 512 10 600 2>&1 | tail -1 | cut -c1-175
done
python scripts/quick_gemm.py 1 384 5 1000 2>&1 | tail -1 | cut -c1-175; B2R_SEED_RANK_L=1 python scripts/quick_gemm.py 1 384 5 1000 2>&1 | tail -1 | cut -c1-175
