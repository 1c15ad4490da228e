mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2m_mgpu_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_mgpu_pytest.log; grep "^case" gpurun_out/r2m_mgpu_pytest.log | tail -9
