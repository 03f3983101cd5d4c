mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 --timeout-method=thread > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2r_pytest.log
echo "== bench"; timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err; echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/r2r_bench_n1.json; tail -5 gpurun_out/r2r_bench_n1.err
echo "== c3"; timeout 300 python scripts/c3_repeat.py 2
echo "== ncu scan_pool"; python scripts/run_search.py 10000000 512 f32 1 48 3 > gpurun_out/r2r_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_pool_kernel -s 1 -c 1 -o gpurun_out/r2r_prof_scan_pool python scripts/run_search.py 10000000 512 f32 1 48 3 > gpurun_out/r2r_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2r_ncu.log
