set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err; echo rc=$?
tail -c 1500 gpurun_out/r2b_bench1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2b_bench1.json'))
for k in ('value','ms_per_step','ms_per_step_min','ms_per_step_p99','repeats','gpu_launches','clocks'):
    print(k, j.get(k))
print('e2e', {k:v for k,v in j['e2e'].items() if k!='blocking_call'})
print('blocking', j['e2e']['blocking_call'])
print('roofline', j['roofline'])
print('batch1', j['batch1'])
print('config0', j['config0'])
print('config3', json.dumps(j.get('config3'), indent=1))
print('config5', json.dumps(j.get('config5'), indent=1))
print('cpu', json.dumps(j.get('cpu_baseline'), indent=1)[:1500])
PY
