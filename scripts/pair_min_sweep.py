"""Where should the CTA-pair kernel take over from the one-CTA tensor-core kernel?  (device-timed)"""
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
qi = evs.IndexFlatIP(d); qi.add_synthetic(256, seed=1)
q = torch.from_numpy(qi.reconstruct_n(0, 256)).cuda()
for rows in (1_000_000, 10_000_000):
    for storage in ("bf16", "f32"):
        idx = evs.IndexFlatIP(d, storage=storage); idx.reserve(rows); idx.add_synthetic(rows, seed=0)
        for nq in (33, 48, 64, 96, 128, 192, 256):
            xq = q[:nq].contiguous()
            res = {}
            for pm in (1000, 1):
                evs.set_option("tc_pair_min_nq", pm)
                res["pair" if pm == 1 else "one"] = round(timed(idx, xq, 48, 20 if rows <= 1_000_000 else 6), 4)
            print(json.dumps(dict(rows=rows, storage=storage, nq=nq, **res)), flush=True)
        evs.set_option("tc_pair_min_nq", 129)
        del idx
print("fallbacks", evs.get_option("tc_fallbacks"))
