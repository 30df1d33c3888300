set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2o_bench2.json 2> gpurun_out/r2o_bench2.err; echo rc=$?
tail -c 2500 gpurun_out/r2o_bench2.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2o_bench2.json'))
for k in ('value','merged_queries_per_s','ms_per_step','ms_per_step_min','ms_per_step_p99','repeats','gpu_launches','clocks','ms_per_step_per_rank','verified','verified_how','comm'):
    print(k, j.get(k))
print('e2e', j['e2e'])
print('replicas', j['replicas'])
print('roofline', j['roofline'])
print('c4', json.dumps(j.get('sharded_c4'), indent=1))
PY
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
