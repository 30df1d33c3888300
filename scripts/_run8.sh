set -x
MET="gpu__time_duration.sum,sm__cycles_elapsed.max,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum.per_cycle_elapsed,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__inst_executed_pipe_tmem.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_2cta.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum,lts__t_sectors_srcunit_tex_op_read.sum,sm__mem_tensor_reads_op_ldt.sum"
# 1) launch list of the bench command
python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/r2g_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/r2g_ncu_bench.log 2>&1
echo rc=$?
# 2) tensor-pipe counters of K3 at batch 256 (pair), batch 1024 list mode, batch 1024 pool mode
for cfg in "256 384 5" "1024 384 5" "1024 384 100" "256 512 10"; do
  set -- $cfg
  ncu --metrics $MET --clock-control none -k regex:gemm_topk -s 4 -c 2 --csv --log-file gpurun_out/r2g_tensor_$1_$2_$3.csv python scripts/quick_gemm.py $1 $2 $3 3 > /dev/null 2>&1
  echo rc=$?
done
B2R_NO_PAIR=1 ncu --metrics $MET --clock-control none -k regex:gemm_topk -s 4 -c 2 --csv --log-file gpurun_out/r2g_tensor_256_384_5_nopair.csv python scripts/quick_gemm.py 256 384 5 3 > /dev/null 2>&1
# 3) full capture of the batch-256 query's kernels
ncu --set full --clock-control none --import-source on -k regex:"gemm_topk|finalize_union|exact_topk|ingest_kernel" -s 12 -c 4 -o gpurun_out/r2g_full python scripts/quick_gemm.py 256 384 5 3 > gpurun_out/r2g_ncu_full.log 2>&1
echo rc=$?
ls -la gpurun_out/ | grep r2g
