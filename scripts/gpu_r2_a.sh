# round 2, first GPU call: the whole GPU suite, smoke, the scan-tail probe, a bench line and the launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/r2a_box.txt 2>&1
nproc >> gpurun_out/r2a_box.txt; free -g >> gpurun_out/r2a_box.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 --timeout-method=thread > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2a_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2a_smoke.log
echo "== probe"; timeout 600 python scripts/scan_tail_probe.py > gpurun_out/r2a_probe_f32.jsonl 2> gpurun_out/r2a_probe_f32.err; echo "probe rc=$?"; cat gpurun_out/r2a_probe_f32.jsonl
timeout 300 python scripts/scan_tail_probe.py --storage bf16 --rows 12500000 --reps 100 > gpurun_out/r2a_probe_bf16.jsonl 2>> gpurun_out/r2a_probe_f32.err; cat gpurun_out/r2a_probe_bf16.jsonl
echo "== bench"; timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2a_bench_n1.json; tail -5 gpurun_out/r2a_bench_n1.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"
