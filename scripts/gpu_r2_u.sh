mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 --timeout-method=thread > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2u_pytest.log
echo "== bench"; timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err; echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/r2u_bench_n1.json | head -8; tail -5 gpurun_out/r2u_bench_n1.err
