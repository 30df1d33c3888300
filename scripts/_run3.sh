mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-c4 > gpurun_out/fused_bench2.json 2> gpurun_out/fused_bench2.err; echo "bench rc=$?"; tail -5 gpurun_out/fused_bench2.err
python - <<'PY'
import json
j = json.load(open('gpurun_out/fused_bench2.json'))
print('value', round(j['value']), 'ms/step', j['ms_per_step'], 'verified', j.get('verified'), 'form', j['comm']['form_timed'], j['comm']['trial_ms_per_step'])
print('replicas', j['replicas']['ms_per_step'], 'nccl serial', j['comm']['nccl_all_gather_ms_per_step'])
PY
