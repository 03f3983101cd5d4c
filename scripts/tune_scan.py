"""Scan-stage tuning sweep on one GPU: mean device time of the scan kernel alone (CUDA events on the
index's stream, evs_index_time_scan) for each option setting.  Usage: python scripts/tune_scan.py [rows] [dim] [storage]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 512
storage = sys.argv[3] if len(sys.argv) > 3 else "f32"
idx = evs.IndexFlatIP(dim, storage=storage)
idx.reserve(rows)
idx.add_synthetic(rows, seed=0)
qi = evs.IndexFlatIP(dim)
qi.add_synthetic(16, seed=1)
q = torch.from_numpy(qi.reconstruct_n(0, 16)).cuda()
esz = 2 if storage == "bf16" else 4
out = []


def run(tag, nq=1, **opts):
    for k in ("scan_variant", "tile_rows", "stages", "ctas_per_sm"):
        evs.set_option(k, opts.get(k, 0))
    ms = idx.time_scan(q[:nq].contiguous(), 48, iters=20)
    gbs = rows * dim * esz / (ms * 1e-3) / 1e9
    rec = dict(tag=tag, nq=nq, ms=round(ms, 4), GBps=round(gbs, 1), **opts)
    out.append(rec)
    print(json.dumps(rec), flush=True)


for c in (1, 2, 3, 4):
    run("direct", scan_variant=1, ctas_per_sm=c)
for tr, st in ((8, 4), (8, 8), (16, 4), (16, 6), (16, 3), (32, 3), (32, 2), (4, 8), (4, 16), (24, 4)):
    run("ring", scan_variant=2, tile_rows=tr, stages=st)
for nq in (2, 3, 4, 8, 16):
    run("direct", nq=nq, scan_variant=1)
    run("ring", nq=nq, scan_variant=2)
