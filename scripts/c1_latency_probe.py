"""BASELINE config 1 (10 000 x 512, one query): device time per search (back-to-back searches of a CUDA tensor, CUDA events)
and end-to-end time (IndexFlatIP.search(numpy) -> numpy), small-shard kernel against the pool kernel, k = 12 and 48.
    python scripts/c1_latency_probe.py [--rows 10000,20000,30000]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", default="10000,20000,30000")
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--storage", default="f32")
ap.add_argument("--small-max", type=int, default=32768, help="routing limit of the small-shard kernel for its leg (sweeps beyond the default)")
ap.add_argument("--ks", default="12,48")
a = ap.parse_args()
qi = evs.IndexFlatIP(a.dim)
qi.add_synthetic(64, seed=1)
qh = qi.reconstruct_n(0, 64)
q = torch.from_numpy(qh).cuda()
for rows in [int(r) for r in a.rows.split(",")]:
    idx = evs.IndexFlatIP(a.dim, storage=a.storage)
    idx.add_synthetic(rows, seed=0)
    for k in [int(x) for x in a.ks.split(",")]:
        for name, small in (("pool kernel", 0), ("small-shard kernel", a.small_max)):
            evs.set_option("small_max_rows", small)
            D = torch.empty((1, k), dtype=torch.float32, device="cuda")
            I = torch.empty((1, k), dtype=torch.int64, device="cuda")
            for i in range(200):
                idx.search(q[i % 64:i % 64 + 1], k, D=D, I=I)
            torch.cuda.synchronize()
            dev = []
            for rnd in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(1000):
                    idx.search(q[i % 64:i % 64 + 1], k, D=D, I=I)
                e1.record()
                torch.cuda.synchronize()
                dev.append(e0.elapsed_time(e1))  # ms per 1000 = us per search
            e2e = []
            for rnd in range(7):
                t0 = time.perf_counter()
                for i in range(1000):
                    idx.search(qh[i % 64:i % 64 + 1], k)
                e2e.append((time.perf_counter() - t0) * 1e3)
            print(json.dumps({"rows": rows, "k": k, "kernel": name, "device_us_min": round(min(dev), 2), "device_us_median": round(float(np.median(dev)), 2),
                              "e2e_us_min": round(min(e2e), 2), "e2e_us_median": round(float(np.median(e2e)), 2)}), flush=True)
    evs.set_option("small_max_rows", 32768)
    del idx
