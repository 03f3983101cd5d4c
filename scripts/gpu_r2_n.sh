mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -q --maxfail=5 --timeout 300 --timeout-method=thread > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2n_pytest.log
echo "== tiny"; timeout 300 python scripts/scan_tail_probe.py --rows 10000,100000 --reps 500 2>&1 | grep "pool static" | cut -c1-700
echo "== mid batches"; timeout 600 python scripts/mid_batch_probe.py --nqs 33,64,128,256 > gpurun_out/r2n_mid.jsonl 2> gpurun_out/r2n_mid.err; cut -c1-200 gpurun_out/r2n_mid.jsonl; tail -3 gpurun_out/r2n_mid.err
echo "== c3"; timeout 300 python scripts/c3_repeat.py 3
