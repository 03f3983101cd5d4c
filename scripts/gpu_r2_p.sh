# 2 x B200: the flag-in-data exchange: sharded parity (both exchanges, guard, loader), GPU tests of the exchange, bench N = 2 and the N = 8 shard size
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "exchange or partial or device_tensor or back_to_back" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2p_pytest.log
rm -f gpurun_out/r2p_check_sharded.log
EVS_CHECK_LOG=gpurun_out/r2p_check_sharded.log timeout 600 $TR --nproc-per-node 2 --master-port 29601 scripts/check_sharded.py > gpurun_out/r2p_check.out 2>&1; echo "check rc=$?"; grep -E "MISMATCH|PARITY|rror" gpurun_out/r2p_check.out | tail -8
timeout 300 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 --steps 200 --warmup 10 --rows 2500000 --no-configs > gpurun_out/r2p_bench_n2_2p5m.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/r2p_bench_n2_2p5m.json | head -4
timeout 300 $TR --nproc-per-node 2 --master-port 29604 scripts/exchange_probe.py --rows 2500000 > gpurun_out/r2p_probe.jsonl 2>&1; grep "^{" gpurun_out/r2p_probe.jsonl
