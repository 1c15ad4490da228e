mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2z_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench.log 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r2z_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'tte', d['time_to_eps']['ms'], d['time_to_eps']['e2e_ms'])
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms_per_sweep'), v['roofline']['frac'])
print(d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_refarm.log 2>&1; echo "ref arm rc=$?"; tail -1 gpurun_out/r2z_refarm.log | cut -c1-400
