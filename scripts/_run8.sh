set -x
MET="gpu__time_duration.sum,sm__cycles_elapsed.max,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum"
python scripts/quick_gemm.py 64 768 20 3 10000000 > /dev/null 2>&1 && \
ncu --metrics $MET --clock-control none -k regex:gemm_topk -s 2 -c 2 --csv --log-file gpurun_out/r2p_b64_768_10m.csv python scripts/quick_gemm.py 64 768 20 3 10000000 > /dev/null 2>&1
B2R_NO_BM64=1 ncu --metrics $MET --clock-control none -k regex:gemm_topk -s 2 -c 2 --csv --log-file gpurun_out/r2p_b64_768_10m_bm128.csv python scripts/quick_gemm.py 64 768 20 3 10000000 > /dev/null 2>&1
ncu --metrics $MET --clock-control none -k regex:gemm_topk -s 4 -c 2 --csv --log-file gpurun_out/r2p_b1_384.csv python scripts/quick_gemm.py 1 384 5 5 > /dev/null 2>&1
B2R_NO_DYN=1 ncu --metrics $MET --clock-control none -k regex:gemm_topk -s 4 -c 2 --csv --log-file gpurun_out/r2p_b1_384_static.csv python scripts/quick_gemm.py 1 384 5 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_topk" -s 4 -c 1 -o gpurun_out/r2p_b1_full python scripts/quick_gemm.py 1 384 5 5 > gpurun_out/r2p_ncu_full.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/r2p_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/r2p_ncu_bench.log 2>&1
echo done
