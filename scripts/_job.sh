mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_lifecycle.py -m gpu -q -x --timeout=900 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/scale_n2_peer.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/scale_n2_peer.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=2 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1), d['e2e']['api'][:60])"
