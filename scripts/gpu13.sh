mkdir -p gpurun_out
C="python scripts/run_search.py 1000000 512 bf16 4096 48 2"
$C > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tc2_scan_kernel -s 3 -c 1 -o gpurun_out/prof_tc2_select -f $C > gpurun_out/ncu_tc2.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_tc2.log
$C > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv $C > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_c3.csv | awk -F'","' '{print $5, $NF}' | tail -14
