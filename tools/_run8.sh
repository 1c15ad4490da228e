mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 600 $TR --master-port 29721 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2n_bench8.log 2> gpurun_out/r2n_bench8.err; echo "bench8 rc=$?"
tail -1 gpurun_out/r2n_bench8.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'ms',d['ms_per_step'],'e2e', round(d['e2e']['value'],1))
print([ (p['shape'],p['ok'],p['rel_x']) for p in d['parity']])
for k,v in d.get('configs',{}).items(): print(k, {kk:v.get(kk) for kk in ('value','ms_per_sweep','ms_per_sweep_1gpu_shard','efficiency_vs_1gpu_shard','frac_of_aggregate_roofline','instance_sweeps_per_s','transport')})
"
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "8" > gpurun_out/r2n_mgpu8_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2n_mgpu8_pytest.log; grep "^case" gpurun_out/r2n_mgpu8_pytest.log | tail -3
