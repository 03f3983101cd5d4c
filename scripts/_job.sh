mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x --timeout=900 2>&1 | tail -3
timeout 300 python scripts/c3_repeat.py 3 2>&1 | tail -3
C="python scripts/run_search.py 1000000 512 bf16 4096 48 2"
$C > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv $C > /dev/null 2>&1
