mkdir -p gpurun_out
echo "== pytest tensorcore"; timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x --timeout 300 --timeout-method=thread > gpurun_out/r2d_pytest_tc.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2d_pytest_tc.log
echo "== pytest normalize"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "normalize or layout" --timeout 200 > gpurun_out/r2d_pytest_norm.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2d_pytest_norm.log
echo "== probe x3"; timeout 300 python scripts/scan_tail_probe.py --rows "" > gpurun_out/r2d_probe_x3.jsonl 2> gpurun_out/r2d_probe.err; echo "probe rc=$?"; cat gpurun_out/r2d_probe_x3.jsonl; tail -3 gpurun_out/r2d_probe.err
echo "== normalize probe"; timeout 300 python scripts/normalize_probe.py > gpurun_out/r2d_normalize.jsonl 2>&1; cat gpurun_out/r2d_normalize.jsonl
