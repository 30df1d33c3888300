python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2o_bench4.json 2> gpurun_out/r2o_bench4.err; echo rc=$?
tail -c 800 gpurun_out/r2o_bench4.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2o_bench4.json'))
for k in ('value','merged_queries_per_s','ms_per_step','ms_per_step_min','repeats','gpu_launches','clocks','verified'):
    print(k, j.get(k))
print('comm', j['comm']['form_timed'], j['comm']['trial_ms_per_step'], j['comm']['nccl_all_gather_ms_per_step'])
print('roofline', j['roofline']['bound'], j['roofline']['frac'], j['roofline']['kernel_ms_per_step'])
print('e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
print('replicas', j['replicas']['qps'], j['replicas']['ms_per_step'])
c=j.get('sharded_c4'); print('c4', {k:c[k] for k in ('qps','ms_per_step','ms_per_step_min','tflops_per_gpu','exact_fallbacks_all_ranks','verified','rows_per_gpu')}, c['roofline']['frac'])
PY
