#!/bin/bash
# N-GPU bench as the driver launches it:  gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_scaling.sh N'
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 \
    > gpurun_out/scaling_bench$N.json 2> gpurun_out/scaling_bench$N.err; echo "bench rc=$?"
python - <<PY
import json
j = json.load(open('gpurun_out/scaling_bench$N.json'))
print('value', round(j['value']), 'ms/step', j['ms_per_step'], 'verified', j.get('verified'), 'exchange form', j['comm']['form_timed'], j['comm']['trial_ms_per_step'])
print('replicas', j['replicas']['ms_per_step'], 'e2e', round(j['e2e']['value']))
c = j.get('sharded_c4')
if c: print('config 4:', round(c['qps']), 'q/s', c['ms_per_step'], 'ms', 'fall-backs', c['exact_fallbacks_all_ranks'], 'verified', c['verified'], 'of sustained tensor peak', round(c['roofline']['frac'], 3))
PY
[ "$N" = 2 ] && python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -2
