mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -q -x --timeout=900 2>&1 | tail -15
for ex in nccl peer; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 200 --warmup 10 --exchange $ex > gpurun_out/scale_n2_$ex.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/scale_n2_$ex.log | cut -c1-600
done
