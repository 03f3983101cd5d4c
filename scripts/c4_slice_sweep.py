"""The per-GPU shard of BASELINE config 4 at N = 8 (1.25M x 768 bf16, 1024 queries) on one GPU: whole-search time against the
CTA-pair kernel's slice length (option tc2_slice_tiles; 0 = the plan's own choice).   python scripts/c4_slice_sweep.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402

rows, d, nq, k = 1_250_000, 768, 1024, 48
qi = evs.IndexFlatIP(d)
qi.add_synthetic(nq, seed=1)
q = torch.from_numpy(qi.reconstruct_n(0, nq)).cuda()
idx = evs.IndexFlatIP(d, storage="bf16")
idx.reserve(rows)
idx.add_synthetic(rows, seed=0)
flop = 2.0 * nq * rows * d
ref = None
for ts in (0, 2, 4, 8, 12, 17, 24, 32, 48, 64):
    evs.set_option("tc2_slice_tiles", ts)
    for _ in range(3):
        D, I = idx.search(q, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        D, I = idx.search(q, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if ref is None:
        ref = (D.clone(), I.clone())
    same = bool(torch.equal(I, ref[1]) and torch.equal(D, ref[0]))
    print(json.dumps({"tc2_slice_tiles": ts, "ms_per_search": round(ms, 4), "tflops": round(flop / ms / 1e9, 1), "same_result": same,
                      "tc_fallbacks": evs.get_option("tc_fallbacks")}), flush=True)
evs.set_option("tc2_slice_tiles", 0)
