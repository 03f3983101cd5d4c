mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 --timeout=900 --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python scripts/tune_scan.py 10000000 512 f32 > gpurun_out/tune_f32.log 2>&1; tail -30 gpurun_out/tune_f32.log
timeout 600 python scripts/tune_scan.py 10000000 512 bf16 > gpurun_out/tune_bf16.log 2>&1; tail -30 gpurun_out/tune_bf16.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log; tail -2 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
