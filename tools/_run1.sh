mkdir -p gpurun_out
for sb in 0 40000 20000 0; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --eps 0 --quick --no-parity --e2e-sweeps 1 --config c2 --layout transposed --slot-bytes $sb > gpurun_out/r2h_c2t_$sb.log 2>&1; tail -1 gpurun_out/r2h_c2t_$sb.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('slot<=$sb c2 transposed', round(d['value'],1), d['config']['launch'])"
done
