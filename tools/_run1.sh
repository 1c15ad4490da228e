# one-GPU validation on a B200 box (gpurun -- 'bash tools/_run1.sh'): GPU tests, smoke, the default bench line,
# the step-wise mat-vec roofline; logs under gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/val_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/val_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/val_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/val_smoke.log
timeout 900 python bench.py > gpurun_out/val_bench.log 2> gpurun_out/val_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/val_bench.log | cut -c1-400
for a in "" "--layout transposed" "--dtype double"; do
n=$(echo $a | tr -d ' -')
timeout 200 python tools/matvec_bench.py $a > gpurun_out/val_matvec_$n.log 2>&1; tail -1 gpurun_out/val_matvec_$n.log | cut -c1-600
done
