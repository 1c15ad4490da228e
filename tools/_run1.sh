mkdir -p gpurun_out
for sb in 0 32768 40960 57344 65536; do for inf in 0 3; do
v=$(timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --eps 0 --quick --no-parity --e2e-sweeps 1 --slot-bytes $sb --inflight $inf 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['config']['launch']['tile_rows'], d['config']['launch']['ring_slots'])")
echo "slot=$sb inflight=$inf: $v"
done; done
for d in 32 64 96; do
v=$(timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --eps 0 --quick --no-parity --e2e-sweeps 1 --dbg $d 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1))")
echo "gate dbg=$d: $v"
done
