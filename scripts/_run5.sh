python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest.log
for nb in 0 1; do
for cfg in "10000000 64 20 200 0 768" "1000000 64 5 300 0 384" "1000000 1 5 300 0 384" "1000000 1 10 300 0 512" "1000000 16 5 300 0 384"; do
echo "== NO_BM64=$nb $cfg"
B2R_NO_BM64=$nb timeout 300 python scripts/pool_large.py $cfg 2>&1 | grep "^rows"
done; done
