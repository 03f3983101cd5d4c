mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -8
echo "pytest rc=${PIPESTATUS[0]}"
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-3000
