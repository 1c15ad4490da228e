mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
grep -E "FAILED|ERROR|passed|failed|rc=" gpurun_out/r2z_pytest.log | cut -c1-300
