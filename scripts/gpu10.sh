mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -3
C="python scripts/run_search.py 1000000 512 bf16 512 48 2"
$C > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 3 -c 1 -o gpurun_out/prof_tc_select2 -f $C > gpurun_out/ncu_tc2.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_tc2.log
