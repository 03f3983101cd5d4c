mkdir -p gpurun_out
CASES="1000000:512:bf16:4096 1000000:512:f32:64"
python scripts/multi_search.py $CASES > gpurun_out/r2s_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches.csv python scripts/multi_search.py $CASES > gpurun_out/r2s_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s_plain.log
python -c "import __graft_entry__ as g; g.smoke()"
