"""Device-timed search latency / throughput for the BASELINE.json configs on one GPU (CUDA events on torch's
current stream around K back-to-back IndexFlatIP.search calls with CUDA tensors).  Usage:
    python scripts/bench_configs.py [c1] [c2] [c3] [m10]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402

PEAK_GBS = 6531.9
try:
    PEAK_GBS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


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


def queries(d, nq):
    qi = evs.IndexFlatIP(d)
    qi.add_synthetic(nq, seed=1)
    return torch.from_numpy(qi.reconstruct_n(0, nq)).cuda()


def run(tag, rows, d, storage, nqs, k, reps=20, **opts):
    idx = evs.IndexFlatIP(d, storage=storage)
    idx.reserve(rows)
    idx.add_synthetic(rows, seed=0)
    esz = 2 if storage == "bf16" else 4
    for nq in nqs:
        for name, val in opts.items():
            evs.set_option(name, val)
        xq = queries(d, nq)
        ms = timed(idx, xq, k, reps if nq <= 256 else 3)
        scan_ms = idx.time_scan(xq, k, iters=5 if nq <= 256 else 2)
        rec = dict(config=tag, rows=rows, d=d, storage=storage, nq=nq, k=k, ms_per_search=round(ms, 4),
                   queries_per_s=round(nq / ms * 1e3, 1), scan_ms=round(scan_ms, 4),
                   db_GBps=round(rows * d * esz / (ms * 1e-3) / 1e9, 1),
                   frac_hbm_peak=round(rows * d * esz / (ms * 1e-3) / 1e9 / PEAK_GBS, 3),
                   TFLOPs=round(2.0 * nq * rows * d / (ms * 1e-3) / 1e12, 2), **opts)
        print(json.dumps(rec), flush=True)
    del idx


which = set(sys.argv[1:]) or {"c1", "c2", "c3", "m10"}
if "c1" in which:
    run("C1 10k x 512 f32", 10_000, 512, "f32", (1,), 12, reps=200)
if "c2" in which:
    run("C2 1M x 512 f32", 1_000_000, 512, "f32", (1, 2, 4, 8, 16, 64), 48)
    run("C2 1M x 512 f32 (GEMV only)", 1_000_000, 512, "f32", (4, 16), 48, tc_min_nq=0)
    evs.set_option("tc_min_nq", 2)
if "c3" in which:
    run("C3 1M x 512 bf16", 1_000_000, 512, "bf16", (1, 16, 128, 1024, 4096), 48)
if "m10" in which:
    run("metric 10M x 512 f32", 10_000_000, 512, "f32", (1, 4, 16, 64), 48)
    run("metric 10M x 512 bf16", 10_000_000, 512, "bf16", (1, 4, 16, 128), 48)
