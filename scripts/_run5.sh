for rep in 1 2; do
for lib in scripts/_ab/libb2r_head_with_trace.so multimodal_rag_b200/libb2r.so; do
echo "== $lib"
B2R_LIB=$PWD/$lib timeout 200 python scripts/pool_large.py 1000000 256 5 400 0 384 2>&1 | grep "^rows"
B2R_LIB=$PWD/$lib timeout 200 python scripts/pool_large.py 1000000 1024 5 150 0 384 2>&1 | grep "^rows"
B2R_LIB=$PWD/$lib timeout 200 python scripts/pool_large.py 1000000 1 5 400 0 384 2>&1 | grep "^rows"
done; done
B2R_TRACE=1 python scripts/pool_large.py 1000000 256 5 50 0 384 2>&1 | grep "^trace"
