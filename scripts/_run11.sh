python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench1.json 2> gpurun_out/r2n_bench1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2n_bench1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2n_ref1.json 2> gpurun_out/r2n_ref1.err; echo "ref rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2n_bench1.json')); r=json.load(open('gpurun_out/r2n_ref1.json'))
print('value',j['value'],'ms',j['ms_per_step'],'min',j['ms_per_step_min'],'p99',j['ms_per_step_p99'],'e2e',j['e2e']['value'],'blocking',j['e2e']['blocking_call']['value'],j['e2e']['blocking_call']['latency_us_per_call'])
print('ref',r['value'],r['cpu_baseline']['cores'], 'ratio', j['value']/r['value'], j['e2e']['value']/r['value'])
rf=j['roofline']; print('roofline',rf['bound'],rf['frac'],rf['kernel_ms_per_step'],'hbm',rf['hbm']['frac'],'tensor_burst',rf['tensor']['frac_of_burst'])
print('clocks',j['clocks'])
print('b1',j['batch1']['qps'],j['batch1']['us_per_query'],j['batch1']['roofline']['frac'],'ingest',j['ingest']['rows_per_s'],j['ingest']['roofline']['frac'],'c0',j['config0'])
c3=j['config3']['filters']
for f in c3: print(f, {b: (round(c3[f][b]['qps']), round(c3[f][b]['e2e_qps']), round(c3[f][b]['roofline']['frac'],3)) for b in ('batch1','batch256')})
c5=j['config5']; print('c5', c5['qps'], c5['ms_per_step'], c5['ms_per_step_min'], c5['roofline']['frac'], c5['ingest']['rows_per_s'], c5['collection_api'])
print('cpu', {k:v for k,v in j['cpu_baseline'].items() if k not in ('hnsw','hnsw_config0')}, j['cpu_baseline']['hnsw']['ef10'], j['cpu_baseline']['hnsw']['ef100'])
PY
python __graft_entry__.py smoke 2>&1 | tail -2
