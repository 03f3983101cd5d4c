mkdir -p gpurun_out
echo "== pytest tensorcore"; timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --maxfail=5 --timeout 300 --timeout-method=thread > gpurun_out/r2k_pytest_tc.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2k_pytest_tc.log
CASES="1000000:512:f32:64 1000000:512:bf16:128 1000000:512:bf16:4096 1000000:512:f32:16 10000:512:f32:1:12"
python scripts/multi_search.py $CASES > gpurun_out/r2k_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_launches.csv python scripts/multi_search.py $CASES > gpurun_out/r2k_ncu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2k_plain.log
