mkdir -p gpurun_out
echo "== pytest parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 --timeout 400 --timeout-method=thread -k "pool or single_query or hand or ties or small or nan or sweep" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2i_pytest.log
echo "== probe"; timeout 600 python scripts/scan_tail_probe.py --rows 1250000,10000000 > gpurun_out/r2i_probe.jsonl 2> gpurun_out/r2i_probe.err; echo "probe rc=$?"; cut -c1-900 gpurun_out/r2i_probe.jsonl | head -8; tail -3 gpurun_out/r2i_probe.err
