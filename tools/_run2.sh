mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "1gpu mode0: $(python bench.py --dbg 0 --steps 20 --warmup 5 --eps 0 --no-cpu --e2e-sweeps 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['ms_per_step'])")"
for d in 0 256; do
  timeout 600 $TR --master-port 2963$((d/256)) bench.py --gpus 2 --steps 20 --warmup 5 --eps 0 --dbg $d > gpurun_out/r2x_bench2_$d.log 2>&1
  echo "2gpu dbg=$d: $(tail -1 gpurun_out/r2x_bench2_$d.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['ms_per_step'])" 2>&1 | tail -1)"
done
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2x_mgpu_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_mgpu_pytest.log
tail -3 gpurun_out/r2x_mgpu_pytest.log | cut -c1-300
