python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2o_bench8.json 2> gpurun_out/r2o_bench8.err; echo rc=$?
tail -c 1500 gpurun_out/r2o_bench8.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2o_bench8.json'))
for k in ('value','merged_queries_per_s','ms_per_step','ms_per_step_min','ms_per_step_p99','repeats','gpu_launches','clocks','ms_per_step_per_rank','verified','verified_how'):
    print(k, j.get(k))
print('comm', j['comm'])
print('e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
print('replicas', j['replicas']['qps'], j['replicas']['ms_per_step'])
c=j.get('sharded_c4'); print('c4', {k:c[k] for k in ('qps','ms_per_step','ms_per_step_min','tflops_per_gpu','exact_fallbacks_all_ranks','verified','rows_per_gpu')}, c['roofline'])
PY
