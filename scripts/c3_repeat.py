"""C3 (1M x 512 bf16, nq = 4096) repeated: run-to-run spread and clocks (device-timed)."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evo_ssearch_b200 as evs
rows, d, nq, k = 1_000_000, 512, 4096, 48
idx = evs.IndexFlatIP(d, storage="bf16"); idx.reserve(rows); idx.add_synthetic(rows, seed=0)
qi = evs.IndexFlatIP(d); qi.add_synthetic(nq, seed=1)
xq = torch.from_numpy(qi.reconstruct_n(0, nq)).cuda()
def clk():
    try:
        return subprocess.check_output(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits"], text=True).strip()
    except Exception as e:
        return str(e)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    for _ in range(2):
        idx.search(xq, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        idx.search(xq, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    scan = idx.time_scan(xq, k, iters=4)
    print(json.dumps(dict(rep=rep, ms=round(ms, 4), qps=round(nq / ms * 1e3), scan_ms=round(scan, 4), clocks=clk())), flush=True)
    time.sleep(0.5)
