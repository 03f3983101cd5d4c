mkdir -p gpurun_out
echo "== pytest tc"; timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --maxfail=5 --timeout 300 --timeout-method=thread > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2t_pytest.log
echo "== c3"; timeout 300 python scripts/c3_repeat.py 3
echo "== mid"; timeout 600 python scripts/mid_batch_probe.py --rows 1000000 --nqs 16,64 2>&1 | cut -c1-175
CASES="1000000:512:bf16:4096 1000000:512:f32:64"
python scripts/multi_search.py $CASES > gpurun_out/r2t_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_launches.csv python scripts/multi_search.py $CASES > gpurun_out/r2t_ncu.log 2>&1; echo "rc=$?"
