"""How long does one single-query search take at the per-GPU shard sizes of the metric (10M x 512 over 1/2/4/8 GPUs), and where
does the scan's tail go?  For each option set: device-timed back-to-back searches (CUDA events) and, from the per-CTA
%globaltimer records (option scan_clock), the spread of the CTAs' scan-loop end times -- the part of the kernel during
which some SMs have already run out of rows.

    python scripts/scan_tail_probe.py [--storage f32|bf16] [--rows 1250000,2500000,...]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--storage", default="f32")
ap.add_argument("--rows", default="1250000,2500000,10000000")
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--reps", type=int, default=200)
a = ap.parse_args()
dev = torch.device("cuda", 0)
qi = evs.IndexFlatIP(a.dim)
qi.add_synthetic(64, seed=1)
q = torch.from_numpy(qi.reconstruct_n(0, 64)).to(dev)
del qi

SETS = [
    ("unfused static (round-1 path: scan + finalize)", dict(fuse_finalize=0, scan_dynamic=0, pool_select=0)),
    ("fused lists static", dict(fuse_finalize=1, scan_dynamic=0, pool_select=0)),
    ("fused lists dynamic c=6", dict(fuse_finalize=1, scan_dynamic=2, scan_chunk_groups=6, pool_select=0)),
    ("fused pool static", dict(fuse_finalize=1, scan_dynamic=0, pool_select=1)),
    ("fused pool, dynamic tail c=2", dict(fuse_finalize=1, scan_dynamic=1, scan_chunk_groups=2, pool_select=1)),
    ("fused pool, dynamic tail c=4", dict(fuse_finalize=1, scan_dynamic=1, scan_chunk_groups=4, pool_select=1)),
]
if os.environ.get("EVS_PROBE_ALL"):
    SETS += [("fused dynamic c=%d" % c, dict(fuse_finalize=1, scan_dynamic=1, scan_chunk_groups=c)) for c in (1, 2, 8)]
for rows in [int(r) for r in a.rows.split(",") if r]:
    idx = evs.IndexFlatIP(a.dim, storage=a.storage)
    idx.reserve(rows)
    idx.add_synthetic(rows, seed=0)
    esz = 2 if a.storage == "bf16" else 4
    for name, opts in SETS:
        for k_, v_ in opts.items():
            evs.set_option(k_, v_)
        for i in range(10):
            idx.search(q[i % 64:i % 64 + 1], 48)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.reps):
            idx.search(q[i % 64:i % 64 + 1], 48)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        rec = {"rows": rows, "storage": a.storage, "options": name, "ms_per_search": round(ms, 4),
               "GBps_whole_search": round(rows * a.dim * esz / ms / 1e6, 1)}
        if opts.get("fuse_finalize"):
            evs.set_option("scan_clock", 1)
            spans, spreads, tails, epi, fin = [], [], [], [], []
            for i in range(8):
                idx.search(q[i:i + 1], 48)
                c = idx.scan_clocks().astype(np.int64)
                st = idx.last_cta_stamps.astype(np.int64)
                t0 = c[:, 0].min()
                end = c[:, 1] - t0
                spans.append(end.max() / 1e3)
                spreads.append((end.max() - end.min()) / 1e3)
                tails.append((end.max() - np.median(end)) / 1e3)
                epi.append((st[0] - t0 - end.max()) / 1e3)  # last scan-loop end -> the last CTA has its ticket
                fin.append((st[1] - st[0]) / 1e3)           # the fused finalise
            evs.set_option("scan_clock", 0)
            fs = idx.finalize_stamps.astype(np.int64)
            rec["last_cta_us"] = {"loop_end_to_sorted": round((st[2] - st[4]) / 1e3, 1), "sorted_to_stored": round((st[3] - st[2]) / 1e3, 1),
                                  "stored_to_ticket": round((st[0] - st[3]) / 1e3, 1), "ticket_to_heads": round((fs[0] - st[0]) / 1e3, 1),
                                  "heads_to_survivors": round((fs[1] - fs[0]) / 1e3, 1), "survivors_to_ranked": round((fs[2] - fs[1]) / 1e3, 1),
                                  "ranked_to_rescored": round((fs[3] - fs[2]) / 1e3, 1), "rescored_to_written": round((fs[4] - fs[3]) / 1e3, 1)}
            rec.update(scan_loop_us=round(float(np.median(spans)), 1), cta_end_spread_us=round(float(np.median(spreads)), 1),
                       last_cta_after_median_us=round(float(np.median(tails)), 1),
                       start_skew_us=round(float((c[:, 0].max() - c[:, 0].min()) / 1e3), 1),
                       epilogue_us=round(float(np.median(epi)), 1), fused_finalize_us=round(float(np.median(fin)), 1))
        print(json.dumps(rec), flush=True)
    del idx
for k_, v_ in dict(fuse_finalize=1, scan_dynamic=1, scan_chunk_groups=4, pool_select=1).items():
    evs.set_option(k_, v_)

# small fp32 batches: 3xTF32 vs single tf32 vs the fp32 GEMV (1M x 512, the C2 shape)
idx = evs.IndexFlatIP(a.dim)
idx.add_synthetic(1_000_000, seed=0)
q16 = q[:16].contiguous()
for name, opts in (("3xTF32 + device guard", dict(x3=1, guard=1)), ("3xTF32, guard off", dict(x3=1, guard=0)),
                   ("single tf32 + device guard", dict(x3=0, guard=1)), ("single tf32, guard off", dict(x3=0, guard=0))):
    for k_, v_ in opts.items():
        evs.set_option(k_, v_)
    for nq in (2, 8, 16, 32):
        qq = q[:nq].contiguous()
        for i in range(5):
            idx.search(qq, 48)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(50):
            idx.search(qq, 48)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"rows": 1_000_000, "nq": nq, "options": name, "ms_per_search": round(e0.elapsed_time(e1) / 50, 4),
                          "scan_ms": round(idx.time_scan(qq, 48, iters=20), 4)}), flush=True)
evs.set_option("x3", 1)
evs.set_option("guard", 1)
