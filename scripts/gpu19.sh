mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x --timeout=900 2>&1 | tail -6
timeout 300 python scripts/bench_configs.py c2 c3 2>&1 | grep -v GEMV | tail -12
