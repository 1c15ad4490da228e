mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2c_smoke.log
timeout 900 python bench.py > gpurun_out/r2c_bench.log 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r2c_bench.log | cut -c1-300
for a in "" "--layout transposed" "--dtype double"; do
n=$(echo $a | tr -d ' -')
timeout 200 python tools/matvec_bench.py $a > gpurun_out/r2c_matvec_$n.log 2>&1; tail -1 gpurun_out/r2c_matvec_$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['config'][:60], {k:(round(v['us_per_call'],2), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict) and 'us_per_call' in v})"
done
for k in rowdot_stream colwsum_fused; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 120 -c 1 -o gpurun_out/r2c_$k -f python tools/matvec_bench.py --reps 1 > gpurun_out/r2c_ncu_$k.log 2>&1
ncu -i gpurun_out/r2c_$k.ncu-rep --page raw --csv > gpurun_out/r2c_${k}_raw.csv 2>/dev/null
rm -f gpurun_out/r2c_$k.ncu-rep
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lasso_fused|gen_fill|scale_rows|rowdot|colwsum|reduce_partials|row_sumsq|neg_copy|objective_kernel|finish_diag" -c 2000 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --eps 0 --quick --no-parity --e2e-sweeps 1 --instance philox > gpurun_out/r2c_ncu_list.log 2>&1; echo "ncu list rc=$?"
