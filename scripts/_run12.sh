for div in 32 64 128; do
for rows in 12500000 50000000; do
echo "== POOL_SAMPLE_DIV=$div rows=$rows"
B2R_POOL_SAMPLE_DIV=$div timeout 600 python scripts/pool_large.py $rows 1024 100 6 4 2>&1 | grep -v "^built\|^exact"
done; done
