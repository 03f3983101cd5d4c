"""Times the one-time kernels of north_star (1) at the metric's size: L2-normalise (in place) and the fp32 -> bf16 layout
kernel over 10M x 512 fp32, against the measured HBM peak.  CUDA events, best / median of `reps`.

    python scripts/normalize_probe.py [--rows 10000000] [--dim 512]
"""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402
from evo_ssearch_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
x = torch.randn(a.rows, a.dim, device="cuda")
out16 = torch.empty(a.rows, a.dim, dtype=torch.bfloat16, device="cuda")


def timed(fn):
    ts = []
    for i in range(a.reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


for name, fn, nbytes in (
        ("l2_normalize f32 in place", lambda: evs.normalize_L2(x), 2 * x.numel() * 4),
        ("f32_to_bf16 layout", lambda: _lib.check(_lib.lib().evs_f32_to_bf16_dev(0, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out16.data_ptr()), x.numel(), None)), x.numel() * 6)):
    med, best = timed(fn)
    print(json.dumps({"kernel": name, "rows": a.rows, "dim": a.dim, "ms_median": round(med, 4), "ms_best": round(best, 4),
                      "GBps_median": round(nbytes / med / 1e6, 1), "frac_of_measured_hbm_peak": round(nbytes / med / 1e6 / peak, 3)}), flush=True)
