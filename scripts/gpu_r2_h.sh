# 2 x B200: sharded parity (exchanges, guard, shard loader), the metric at N = 2, and the N = 8 shard size (1.25M rows per GPU) at N = 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
rm -f gpurun_out/r2h_check_sharded.log
EVS_CHECK_LOG=gpurun_out/r2h_check_sharded.log timeout 600 $TR --nproc-per-node 2 --master-port 29601 scripts/check_sharded.py > gpurun_out/r2h_check.out 2>&1; echo "check rc=$?"; grep -E "MISMATCH|PARITY|Error|error|load" gpurun_out/r2h_check.out | tail -12
timeout 400 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "bench n2 rc=$?"; python scripts/show_bench.py gpurun_out/r2h_bench_n2.json; tail -3 gpurun_out/r2h_bench_n2.err
for fuse in 1 0; do
timeout 300 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 --steps 200 --warmup 10 --rows 2500000 --no-configs --no-parity --set fuse_finalize=$fuse > gpurun_out/r2h_bench_n2_2p5m_fuse$fuse.json 2> gpurun_out/r2h_bench_n2_2p5m.err; echo "bench n2 2.5M fuse=$fuse rc=$?"; python scripts/show_bench.py gpurun_out/r2h_bench_n2_2p5m_fuse$fuse.json | head -3
done
timeout 300 python bench.py --steps 200 --warmup 10 --rows 1250000 --no-configs --no-parity --no-cpu-baseline > gpurun_out/r2h_bench_n1_1p25m.json 2>&1; python scripts/show_bench.py gpurun_out/r2h_bench_n1_1p25m.json | head -3
