"""17-32 queries: on-chip-heap tensor-core scan (tc_heap_max_nq 32) against the threshold scan (tc_heap_max_nq 16),
device-timed whole search.   python scripts/heap_route_sweep.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402


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
qi = evs.IndexFlatIP(d)
qi.add_synthetic(64, seed=1)
q = torch.from_numpy(qi.reconstruct_n(0, 64)).cuda()
for rows in (1_000_000, 3_000_000, 10_000_000):
    for storage in ("f32", "bf16"):
        idx = evs.IndexFlatIP(d, storage=storage)
        idx.reserve(rows)
        idx.add_synthetic(rows, seed=0)
        for nq in (12, 16, 17, 20, 24, 28, 32):
            xq = q[:nq].contiguous()
            res = {}
            for name, hm in (("heap", 32), ("threshold", 8)):
                evs.set_option("tc_heap_max_nq", hm)
                res[name] = round(timed(idx, xq, 48, 30 if rows <= 1_000_000 else 8), 4)
            print(json.dumps(dict(rows=rows, storage=storage, nq=nq, **res)), flush=True)
        evs.set_option("tc_heap_max_nq", 32)
        del idx
        torch.cuda.empty_cache()
