mkdir -p gpurun_out
C="python scripts/run_search.py 1000000 512 f32 1 48 4"
$C > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --cache-control none --import-source on -k regex:finalize_kernel -s 2 -c 1 -o gpurun_out/prof_finalize -f $C > gpurun_out/ncu_fin.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_fin.log
