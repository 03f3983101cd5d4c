mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x --timeout=400 2>&1 | tail -15
echo "pytest rc=${PIPESTATUS[0]}"
timeout 300 python scripts/bench_configs.py c2 c3 2>&1 | tail -16
C="python scripts/run_search.py 1000000 512 bf16 4096 48 2"
$C > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv $C > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_c3.csv | awk -F'","' '{print $5, $NF}' | tail -6
C="python scripts/run_search.py 1000000 512 f32 16 48 2"
$C > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv $C > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_c2.csv | awk -F'","' '{print $5, $NF}' | tail -6
