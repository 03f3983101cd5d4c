mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -q -x --timeout=900 2>&1 | tail -8
echo "pytest rc=${PIPESTATUS[0]}"
timeout 300 python scripts/bench_configs.py c3 m10 2>&1 | tail -16
