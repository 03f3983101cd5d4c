# 8 x B200: sharded parity (both exchanges), the metric at N = 8 / 4 / 2, BASELINE configs 4 and 5.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
EVS_CHECK_LIGHT=1 timeout 400 $TR --nproc-per-node 8 --master-port 29601 scripts/check_sharded.py 2>&1 | grep -E "MISMATCH|PARITY|Error|error" | tail -5
run() { # name nproc args...
  name=$1; np=$2; shift 2
  timeout 300 $TR --nproc-per-node $np --master-port 29602 bench.py --gpus $np --steps 100 --warmup 5 "$@" > gpurun_out/scale8_$name.log 2>&1
  echo "$name rc=$?"; tail -1 gpurun_out/scale8_$name.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('  ', d['config']['workload'], 'N=',d['n_gpus'], d['config'].get('exchange'), 'q/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'scan_ms', round(d['roofline']['scan_ms_per_search'],4), d['clocks']['reasons'])
except Exception as e: print('   parse failed', e)
"
}
run m10_n8_peer 8
run m10_n8_nccl 8 --exchange nccl
run m10_n4_peer 4
run m10_n2_peer 2
run c5_100m_bf16_nq1 8 --rows 100000000 --storage bf16
run c5_100m_bf16_nq16 8 --rows 100000000 --storage bf16 --nq 16
run c5_100m_bf16_nq256 8 --rows 100000000 --storage bf16 --nq 256 --steps 30
run c4_10m768_nq1 8 --rows 10000000 --dim 768
run c4_10m768_bf16_nq1024 8 --rows 10000000 --dim 768 --nq 1024 --storage bf16 --steps 20
