for cfg in "1000000 256 5 200 0 384" "1000000 64 5 200 0 384" "1000000 1 5 200 0 384" "1000000 1024 5 50 0 384"; do
echo "== $cfg"
B2R_TRACE=1 timeout 200 python scripts/pool_large.py $cfg 2>&1 | grep -v "^built" | sed 's/waits.*ctas/ctas/'
done
