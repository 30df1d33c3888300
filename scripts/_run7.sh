timeout 900 python -m pytest tests/test_gpu_host_tables.py tests/test_reference_host_logic.py tests/test_gpu_manager.py tests/test_gpu_persist.py tests/test_gpu_where.py tests/test_chroma_import.py -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2f_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2f_bench1.json 2> gpurun_out/r2f_bench1.err; echo rc=$?
tail -c 1500 gpurun_out/r2f_bench1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2f_bench1.json'))
for k in ('value','ms_per_step','ms_per_step_min','repeats','gpu_launches','clocks'):
    print(k, j.get(k))
print('e2e', j['e2e']['value'], j['e2e']['blocking_call']['value'])
print('roofline', j['roofline']['frac'], j['roofline']['kernel_ms_per_step'], j['roofline']['tensor'])
print('batch1', j['batch1']['qps'], j['batch1']['roofline']['frac'])
c3=j['config3']['filters']
for f in c3: print(f, {b: (c3[f][b]['qps'], c3[f][b]['roofline']['frac']) for b in ('batch1','batch256')})
c5=j['config5']; print('c5', c5['qps'], c5['ms_per_step'], c5['roofline']['frac'], c5['collection_api'])
PY
