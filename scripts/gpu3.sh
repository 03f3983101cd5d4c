mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$B > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_direct -s 3 -c 2 -o gpurun_out/prof_scan_f32 -f $B > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full.log
