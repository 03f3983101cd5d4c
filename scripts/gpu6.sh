mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --maxfail=6 --timeout=600 --durations=6 > gpurun_out/tc_tests.log 2>&1; echo "rc=$?" >> gpurun_out/tc_tests.log; tail -40 gpurun_out/tc_tests.log | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lifecycle.py -m gpu -q --maxfail=6 --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log | cut -c1-400
