mkdir -p gpurun_out
python scripts/run_search.py 1000000 512 f32 16 48 3 > gpurun_out/r2e_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 2 -c 2 -o gpurun_out/r2e_prof_x3 python scripts/run_search.py 1000000 512 f32 16 48 3 > gpurun_out/r2e_ncu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2e_ncu.log; cat gpurun_out/r2e_plain.log | tail -2
