"""Batches of 2..256 queries: device-timed whole-search time (CUDA events around back-to-back searches of a CUDA tensor)
against ONE pass over the rows at the measured HBM peak.  python scripts/mid_batch_probe.py [--rows 1000000,10000000]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", default="1000000,10000000")
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--nqs", default="16,32,33,64,128,256")
a = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
qi = evs.IndexFlatIP(a.dim)
qi.add_synthetic(256, seed=1)
q = torch.from_numpy(qi.reconstruct_n(0, 256)).cuda()
del qi
for rows in [int(r) for r in a.rows.split(",")]:
    for storage in ("f32", "bf16"):
        idx = evs.IndexFlatIP(a.dim, storage=storage)
        idx.reserve(rows)
        idx.add_synthetic(rows, seed=0)
        esz = 2 if storage == "bf16" else 4
        for nq in [int(x) for x in a.nqs.split(",")]:
            xq = q[:nq].contiguous()
            reps = 30 if rows <= 2_000_000 else 8
            for _ in range(3):
                idx.search(xq, 48)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = evs.kernel_launches()
            e0.record()
            for _ in range(reps):
                idx.search(xq, 48)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            r, u = idx.guard_stats()
            print(json.dumps({"rows": rows, "storage": storage, "nq": nq, "ms_per_search": round(ms, 4), "queries_per_s": round(nq / ms * 1e3),
                              "frac_one_pass_of_measured_hbm": round(rows * a.dim * esz / ms / 1e6 / peak, 3),
                              "launches_per_search": (evs.kernel_launches() - l0) / reps, "device_reruns": r, "uncertified": u}), flush=True)
        del idx
        torch.cuda.empty_cache()
