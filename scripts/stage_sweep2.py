"""Ring depth of the one-CTA tensor-core scan in the HBM-bound regime (10M x 512, device-timed scan)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evo_ssearch_b200 as evs
rows, d = 10_000_000, 512
qi = evs.IndexFlatIP(d); qi.add_synthetic(128, seed=1)
q = torch.from_numpy(qi.reconstruct_n(0, 128)).cuda()
for storage in ("bf16", "f32"):
    idx = evs.IndexFlatIP(d, storage=storage); idx.reserve(rows); idx.add_synthetic(rows, seed=0)
    esz = 2 if storage == "bf16" else 4
    for nq in (16, 48, 128) if storage == "bf16" else (16, 48):
        for st in (6, 8, 10, 12, 14):
            evs.set_option("tc_stages", st)
            ms = idx.time_scan(q[:nq].contiguous(), 48, iters=8)
            print(json.dumps(dict(storage=storage, nq=nq, max_stages=st, scan_ms=round(ms, 4), GBps=round(rows*d*esz/ms/1e6, 1))), flush=True)
    evs.set_option("tc_stages", 8)
    del idx
