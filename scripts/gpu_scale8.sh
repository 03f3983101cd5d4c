# 8 x B200: sharded parity log, the metric at N = 8 (with the C4 / C5 legs), N = 4, N = 2, and the exchange probe.
# Outputs under gpurun_out/scale8_*; copy what should be kept into profiles/.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
rm -f gpurun_out/scale8_check_sharded.log
EVS_CHECK_LIGHT=1 EVS_CHECK_LOG=gpurun_out/scale8_check_sharded.log timeout 500 $TR --nproc-per-node 8 --master-port 29601 scripts/check_sharded.py > gpurun_out/scale8_check.out 2>&1
echo "check rc=$?"; grep -E "MISMATCH|PARITY|rror" gpurun_out/scale8_check.out | tail -6
for n in 8 4 2; do
  extra="--no-configs"; [ $n = 8 ] && extra=""
  timeout 600 $TR --nproc-per-node $n --master-port 2960$n bench.py --gpus $n --steps 300 --warmup 10 $extra > gpurun_out/scale8_bench_n$n.json 2> gpurun_out/scale8_bench_n$n.err
  echo "bench n$n rc=$?"; python scripts/show_bench.py gpurun_out/scale8_bench_n$n.json
done
timeout 300 $TR --nproc-per-node 8 --master-port 29611 scripts/exchange_probe.py > gpurun_out/scale8_exchange_probe.jsonl 2> gpurun_out/scale8_probe.err
grep "^{" gpurun_out/scale8_exchange_probe.jsonl
