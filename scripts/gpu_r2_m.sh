mkdir -p gpurun_out
timeout 300 python scripts/scan_tail_probe.py --rows 10000,100000 --reps 500 2>&1 | grep -v '"rows": 1000000' | cut -c1-1000
