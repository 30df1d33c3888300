#!/bin/bash
# What the driver does at round end, in one gpurun call:  gpurun --timeout 1800 -- 'bash scripts/gpu_validate.sh'
# GPU tests, smoke(), the N = 1 bench and the CPU (reference) arm; outputs under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/validate_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/validate_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/validate_bench1.json 2> gpurun_out/validate_bench1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/validate_ref1.json 2> gpurun_out/validate_ref1.err; echo "reference rc=$?"
python - <<'PY'
import json
j = json.load(open('gpurun_out/validate_bench1.json')); r = json.load(open('gpurun_out/validate_ref1.json'))
print('value', round(j['value']), 'first region', round(j['value_first_region']), 'ms/step', j['ms_per_step'], 'e2e', round(j['e2e']['value']),
      'reference', round(r['value'], 1), 'on', r['cpu_baseline']['cores'], 'cores')
rf = j['roofline']; print('roofline', rf['bound'], round(rf['frac'], 3), 'kernel ms', rf['kernel_ms_per_step'], 'clocks', j['clocks'])
print('batch1', round(j['batch1']['qps']), 'config3', {f: {b: round(v[b]['qps']) for b in ('batch1', 'batch256')} for f, v in j['config3']['filters'].items()},
      'config5', round(j['config5']['qps']), 'via collection', round(j['config5']['collection_api']['qps']))
PY
