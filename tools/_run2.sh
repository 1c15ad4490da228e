mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for d in 0 8192 16384 256; do
  timeout 600 $TR --master-port 29650 bench.py --gpus 2 --steps 20 --warmup 5 --eps 0 --quick --no-parity --dbg $d > gpurun_out/r2aa_bench2_$d.log 2>&1
  echo "2gpu dbg=$d: $(tail -1 gpurun_out/r2aa_bench2_$d.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['ms_per_step'])" 2>&1 | tail -1)"
done
