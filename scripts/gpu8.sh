mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=600 2>&1 | tail -2
for cfg in "1000000 512 f32 1 48 5"; do
  set -- $cfg
  C="python scripts/run_search.py $cfg"
  $C > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$1.csv $C > /dev/null 2>&1
  echo "rows=$1 rc=$?"
  grep -E "scan_|finalize" gpurun_out/launches_$1.csv | awk -F'","' '{print $5, $NF}' | tail -4
done
