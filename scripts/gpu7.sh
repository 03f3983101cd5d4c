mkdir -p gpurun_out
C="python scripts/run_search.py 1000000 512 bf16 1024 48 2"
$C > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 2 -c 1 -o gpurun_out/prof_tc_select -f $C > gpurun_out/ncu_tc.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_tc.log
