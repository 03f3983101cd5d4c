mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -q -x --timeout=900 2>&1 | tail -8
timeout 300 python scripts/bench_configs.py c2 c3 m10 2>&1 | grep -v GEMV | grep -E '"nq": (4|8|16|64|128),' 
