mkdir -p gpurun_out
for r in 1 2 3; do
for t in 0 3 4 6; do
  if [ $t = 0 ]; then unset B2R_SEED_TILES; else export B2R_SEED_TILES=$t; fi
  echo "== B2R_SEED_TILES=$t"; python scripts/quick_gemm.py 256 384 5 1000 2>&1 | tail -1
done; done 2>&1 | tee gpurun_out/seed_tiles2.log
