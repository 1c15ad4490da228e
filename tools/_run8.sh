mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
i=0
for v in "ipc 1 0" "symm 0 0" "symm 1 0" "symm 1 512" "symm 0 256"; do
  set -- $v; i=$((i+1))
  B200L_COMM=$1 B200L_MULTICAST=$2 timeout 600 $TR --master-port 2972$i bench.py --gpus 8 --steps 20 --warmup 5 --quick --eps 0 --no-parity --dbg $3 > gpurun_out/r02b_bench8_$i.log 2>&1
  echo "8gpu comm=$1 mc=$2 dbg=$3: $(tail -1 gpurun_out/r02b_bench8_$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['ms_per_step'], d['config']['workload'][d['config']['workload'].find('transport'):][:28])" 2>&1 | tail -1)"
done
