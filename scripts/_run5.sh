for st in 0 1 2 3; do
echo "== B2R_SEED_TILES=$st  (0 = default)"
B2R_SEED_TILES=$st timeout 200 python scripts/pool_large.py 1000000 1024 5 100 0 384 2>&1 | grep "^rows"
B2R_SEED_TILES=$st timeout 200 python scripts/pool_large.py 1000000 512 5 150 0 384 2>&1 | grep "^rows"
B2R_SEED_TILES=$st timeout 200 python scripts/pool_large.py 1000000 1024 20 100 0 384 2>&1 | grep "^rows"
done
