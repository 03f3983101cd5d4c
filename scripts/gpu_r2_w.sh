# 8 x B200, final: parity log, the metric at N = 8 (with C4 / C5 legs), N = 4, N = 2, exchange probe
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
rm -f gpurun_out/r2w_check_sharded_8gpu.log
EVS_CHECK_LIGHT=1 EVS_CHECK_LOG=gpurun_out/r2w_check_sharded_8gpu.log timeout 500 $TR --nproc-per-node 8 --master-port 29601 scripts/check_sharded.py > gpurun_out/r2w_check.out 2>&1; echo "check rc=$?"; grep -E "MISMATCH|PARITY|rror" gpurun_out/r2w_check.out | tail -6
timeout 600 $TR --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 --steps 300 --warmup 10 > gpurun_out/r2w_bench_n8.json 2> gpurun_out/r2w_bench_n8.err; echo "bench n8 rc=$?"; python scripts/show_bench.py gpurun_out/r2w_bench_n8.json; tail -3 gpurun_out/r2w_bench_n8.err
timeout 300 $TR --nproc-per-node 4 --master-port 29603 bench.py --gpus 4 --steps 300 --warmup 10 --no-configs > gpurun_out/r2w_bench_n4.json 2> gpurun_out/r2w_bench_n4.err; echo "bench n4 rc=$?"; python scripts/show_bench.py gpurun_out/r2w_bench_n4.json | head -2
timeout 300 $TR --nproc-per-node 2 --master-port 29604 bench.py --gpus 2 --steps 100 --warmup 10 --no-configs > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err; echo "bench n2 rc=$?"; python scripts/show_bench.py gpurun_out/r2w_bench_n2.json | head -2
timeout 300 $TR --nproc-per-node 8 --master-port 29611 scripts/exchange_probe.py > gpurun_out/r2w_exchange_probe.jsonl 2> gpurun_out/r2w_probe.err; echo "probe rc=$?"; grep "^{" gpurun_out/r2w_exchange_probe.jsonl | head -3
