mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2y2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2y2_pytest.log
timeout 900 python bench.py > gpurun_out/r2y2_bench.log 2> gpurun_out/r2y2_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r2y2_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'tte', d['time_to_eps']['ms'], d['time_to_eps']['e2e_ms'])
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms_per_sweep'), v['roofline']['frac'])
"
for d in 0 7340032 0 7340032; do
v=$(timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --eps 0 --quick --no-parity --e2e-sweeps 1 --dbg $d 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['ms_per_step'])")
echo "dbg=$d (7340032 = second gate off): $v"
done
