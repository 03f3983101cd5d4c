mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -q -x --timeout=900 2>&1 | tail -4
timeout 600 python scripts/tc_tune.py 2>&1 | tee gpurun_out/tc_tune.log | tail -60
