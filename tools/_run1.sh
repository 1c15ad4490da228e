mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "matvec" > gpurun_out/r2bd_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2bd_pytest.log
for a in "" "--layout transposed" "--dtype double" "--dtype double --layout transposed"; do
n=$(echo $a | tr -d ' -')
timeout 200 python tools/matvec_bench.py $a > gpurun_out/r2bd_matvec_$n.log 2>&1; tail -1 gpurun_out/r2bd_matvec_$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['config'][:60], {k:(round(v['us_per_call'],2), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict) and 'us_per_call' in v})"
timeout 200 python tools/matvec_bench.py $a --dbg 262144 > gpurun_out/r2bd_matvec_ldg_$n.log 2>&1; tail -1 gpurun_out/r2bd_matvec_ldg_$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('LDG', d['config'][:60], {k:(round(v['us_per_call'],2), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict) and 'us_per_call' in v})"
done
