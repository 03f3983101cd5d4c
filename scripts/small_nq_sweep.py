"""1-4 queries: CUDA-core GEMV scan vs tensor-core scan with on-chip heaps (device-timed whole search)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evo_ssearch_b200 as evs

def timed(idx, xq, k, reps):
    for _ in range(3):
        idx.search(xq, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        idx.search(xq, k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

d = 512
qi = evs.IndexFlatIP(d); qi.add_synthetic(8, seed=1)
q = torch.from_numpy(qi.reconstruct_n(0, 8)).cuda()
for rows in (1_000_000, 10_000_000):
    for storage in ("f32", "bf16"):
        idx = evs.IndexFlatIP(d, storage=storage); idx.reserve(rows); idx.add_synthetic(rows, seed=0)
        for nq in (1, 2, 3, 4):
            xq = q[:nq].contiguous()
            res = {}
            for name, tcmin in (("gemv", 0), ("tc", 1)):
                evs.set_option("tc_min_nq", tcmin)
                res[name] = round(timed(idx, xq, 48, 30 if rows <= 1_000_000 else 8), 4)
            print(json.dumps(dict(rows=rows, storage=storage, nq=nq, **res)), flush=True)
        evs.set_option("tc_min_nq", 2)
        del idx
