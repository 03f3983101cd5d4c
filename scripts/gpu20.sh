mkdir -p gpurun_out
timeout 300 python scripts/c3_repeat.py 6 2>&1 | tail -8
C="python scripts/run_search.py 1000000 512 bf16 4096 48 2"
$C > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tc2_scan_kernel -s 3 -c 1 -o gpurun_out/prof_tc2_select_v2 -f $C > gpurun_out/ncu_tc2v2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_tc2v2.log
