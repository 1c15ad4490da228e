mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for c in 176 192 208 224 176; do
if [ $c = 176 ]; then unset B200L_LIB; else export B200L_LIB=$PWD/_ab_old/libb200lasso_c$c.so; fi
a=$(timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --eps 0 --quick --no-parity --e2e-sweeps 1 --dbg 4096 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1))")
b=$(timeout 600 $TR --master-port 29650 bench.py --gpus 2 --steps 20 --warmup 5 --eps 0 --quick --no-parity 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['ms_per_step'])")
echo "cregs=$c: 1gpu-mode1 $a | 2gpu $b"
done
unset B200L_LIB
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2k_mgpu_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2k_mgpu_pytest.log
