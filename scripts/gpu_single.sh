# One B200: the GPU test suite, smoke(), the default bench line, the reference arm, then (one ncu call per gpurun call) the
# launch list of a short bench run.  Outputs under gpurun_out/single_*; copy what should be kept into profiles/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 --timeout-method=thread > gpurun_out/single_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/single_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/single_bench_n1.json 2> gpurun_out/single_bench_n1.err; echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/single_bench_n1.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/single_bench_ref.json 2> gpurun_out/single_bench_ref.err; echo "reference arm rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/single_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/single_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/single_ncu.log 2>&1; echo "ncu rc=$?"
