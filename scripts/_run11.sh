python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2i_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench1.json 2> gpurun_out/r2i_bench1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2i_bench1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2i_ref1.json 2> gpurun_out/r2i_ref1.err; echo "ref rc=$?"
python -c "
import json
j=json.load(open('gpurun_out/r2i_bench1.json')); r=json.load(open('gpurun_out/r2i_ref1.json'))
print('value',j['value'],'e2e',j['e2e']['value'],'ref',r['value'],r['cpu_baseline']['cores'],'ms',j['ms_per_step'],'kern',j['roofline']['kernel_ms_per_step'],'frac',j['roofline']['frac'],j['roofline']['tensor']['frac_sustained'])
print('b1',j['batch1']['qps'],j['batch1']['roofline']['frac'],'ingest',j['ingest']['rows_per_s'],j['ingest']['roofline']['frac'])
print('cpu',j['cpu_baseline']['value'],j['cpu_baseline']['cores'])
"
python __graft_entry__.py smoke 2>&1 | tail -2
