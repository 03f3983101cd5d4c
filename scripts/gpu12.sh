mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x --timeout=300 -k "pair" 2>&1 | tail -25
echo "pytest rc=${PIPESTATUS[0]}"
timeout 300 python scripts/bench_configs.py c3 2>&1 | tail -8
