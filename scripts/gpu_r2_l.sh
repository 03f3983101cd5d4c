mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -q --maxfail=5 --timeout 300 --timeout-method=thread -k "overflow or guard or pool or hand or small or c1_golden or single_query" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2l_pytest.log
echo "== skip mid"
echo "== c1"; python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, evo_ssearch_b200 as evs
idx = evs.IndexFlatIP(512); idx.add_synthetic(10000, seed=0)
qi = evs.IndexFlatIP(512); qi.add_synthetic(64, seed=1); qh = qi.reconstruct_n(0, 64); qd = torch.from_numpy(qh).cuda()
for i in range(50): idx.search(qd[:1], 12)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(1000): idx.search(qd[i % 64:i % 64 + 1], 12)
e1.record(); torch.cuda.synchronize()
print("C1 device us/query (back-to-back, launch-bound):", round(e0.elapsed_time(e1), 3))
evs.set_option("scan_clock", 1)
idx.search(qd[:1], 12); c = idx.scan_clocks(); st = idx.last_cta_stamps
print("C1 kernel span us (first CTA start -> last CTA done):", (int(st[1]) - int(c[:, 0].min())) / 1e3)
PY
