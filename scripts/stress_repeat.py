"""Stability check on one GPU: thousands of back-to-back searches (no host synchronisation in between) must return the same
bits every time -- exercises the pool reset by the last CTA, the dynamic tail counter, the guard counters and the programmatic
chaining of consecutive searches.  python scripts/stress_repeat.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402

qi = evs.IndexFlatIP(512)
qi.add_synthetic(64, seed=1)
q = torch.from_numpy(qi.reconstruct_n(0, 64)).cuda()
total_bad = 0
for rows, k, reps in ((1_250_000, 48, 60), (10_000, 12, 200)):
    idx = evs.IndexFlatIP(512)
    idx.add_synthetic(rows, seed=0)
    ref = [tuple(t.clone() for t in idx.search(q[i:i + 1], k)) for i in range(64)]
    torch.cuda.synchronize()
    bad = 0
    for rep in range(reps):
        outs = [idx.search(q[i:i + 1], k) for i in range(64)]  # 64 searches enqueued back to back
        torch.cuda.synchronize()
        bad += sum(0 if (torch.equal(D, ref[i][0]) and torch.equal(I, ref[i][1])) else 1 for i, (D, I) in enumerate(outs))
    print(f"{reps * 64} back-to-back single-query searches at {rows} rows, k={k}: {bad} mismatches", flush=True)
    total_bad += bad
    if rows <= 32_768:
        # the host path of a small shard: the query in the kernel's parameter block, the completion word polled; interleaved
        # with device-tensor searches of the same handle
        qh = q.cpu().numpy()
        bad = 0
        for rep in range(300):
            for i in range(64):
                D, I = idx.search(qh[i:i + 1], k)
                if not (np.array_equal(D, ref[i][0].cpu().numpy()) and np.array_equal(I, ref[i][1].cpu().numpy())):
                    bad += 1
                if i % 16 == 0:
                    idx.search(q[i:i + 1], k)
        print(f"{300 * 64} host (numpy) single-query searches at {rows} rows, k={k}: {bad} mismatches", flush=True)
        total_bad += bad
    if rows > 100_000:
        q16 = q[:16].contiguous()
        r16 = tuple(t.clone() for t in idx.search(q16, 48))
        bad = sum(0 if all(torch.equal(a, b) for a, b in zip(idx.search(q16, 48), r16)) else 1 for _ in range(300))
        print(f"300 16-query searches at {rows} rows: {bad} mismatches", flush=True)
        total_bad += bad
sys.exit(1 if total_bad else 0)
