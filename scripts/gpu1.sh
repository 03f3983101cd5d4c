mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv; nproc; free -g | head -2; } > gpurun_out/box.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 --timeout=600 -x --deselect tests/test_gpu_parity.py::test_metric_size_10m_x_512_properties > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for v in 1 2; do timeout 600 python bench.py --steps 50 --warmup 5 --variant $v --no-cpu-baseline > gpurun_out/bench_v$v.log 2>&1; echo "rc=$?" >> gpurun_out/bench_v$v.log; done
tail -3 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_v1.log; tail -2 gpurun_out/bench_v2.log
