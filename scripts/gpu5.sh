mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x --timeout=300 -k "raw_scores" > gpurun_out/tc_raw.log 2>&1; echo "rc=$?" >> gpurun_out/tc_raw.log; tail -30 gpurun_out/tc_raw.log
