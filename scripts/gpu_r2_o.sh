mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29611 scripts/exchange_probe.py > gpurun_out/r2o_exchange_probe.jsonl 2> gpurun_out/r2o_probe.err; echo "probe rc=$?"; cat gpurun_out/r2o_exchange_probe.jsonl; tail -3 gpurun_out/r2o_probe.err
for dyn in 0 1; do
timeout 300 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --steps 300 --warmup 10 --no-configs --no-parity --set scan_dynamic=$dyn --set scan_chunk_groups=4 > gpurun_out/r2o_bench_n8_dyn$dyn.json 2> gpurun_out/r2o_bench.err; echo "bench n8 dyn=$dyn rc=$?"; python scripts/show_bench.py gpurun_out/r2o_bench_n8_dyn$dyn.json | head -2
done
