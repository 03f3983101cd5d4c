mkdir -p gpurun_out
echo "== pytest x3"; timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "raw_scores or x3 or guard" --timeout 300 > gpurun_out/r2f_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2f_pytest.log
echo "== probe x3"; timeout 300 python scripts/scan_tail_probe.py --rows "" > gpurun_out/r2f_probe_x3.jsonl 2> gpurun_out/r2f_probe.err; echo "probe rc=$?"; cat gpurun_out/r2f_probe_x3.jsonl; tail -3 gpurun_out/r2f_probe.err
