mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 --timeout-method=thread > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2c_pytest.log
echo "== probe"; timeout 600 python scripts/scan_tail_probe.py --rows 1250000,10000000 > gpurun_out/r2c_probe_f32.jsonl 2> gpurun_out/r2c_probe.err; echo "probe rc=$?"; cat gpurun_out/r2c_probe_f32.jsonl; tail -3 gpurun_out/r2c_probe.err
echo "== bench"; timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/r2c_bench_n1.json; tail -5 gpurun_out/r2c_bench_n1.err
echo "== ncu launches (C1 + metric)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
