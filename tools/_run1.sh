mkdir -p gpurun_out
for d in 0 4096 0 4096; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --eps 0 --quick --no-parity --e2e-sweeps 1 --dbg $d > gpurun_out/r2j_c2_$d.log 2>&1; tail -1 gpurun_out/r2j_c2_$d.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dbg $d c2', round(d['value'],1), d['ms_per_step'])"
done
