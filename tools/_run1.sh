mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2l_bench.log 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r2l_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms_per_sweep'), v['roofline']['frac'], v.get('launch'))
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lasso_fused -s 4 -c 1 -o gpurun_out/r2l_c4t -f python bench.py --steps 5 --warmup 3 --no-cpu --eps 0 --quick --no-parity --e2e-sweeps 1 --config c4shard --layout transposed > gpurun_out/r2l_ncu_c4t.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2l_c4t.ncu-rep --page raw --csv > gpurun_out/r2l_c4t_raw.csv 2>/dev/null
rm -f gpurun_out/r2l_c4t.ncu-rep
