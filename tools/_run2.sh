# multi-GPU validation (gpurun --gpus N -- 'bash tools/_run2.sh'): the sharded solves against the oracle on
# 2 / 4 / 8 GPUs (as many as the box has) and the bench line on all GPUs of the box
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/val_mgpu_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/val_mgpu_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/val_bench$N.log 2> gpurun_out/val_bench$N.err; echo "bench rc=$?"
tail -1 gpurun_out/val_bench$N.log | cut -c1-400
