"""One search call on a synthetic index (profiling target).  Usage: run_search.py rows d storage nq k [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402

rows, d, storage, nq, k = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
idx = evs.IndexFlatIP(d, storage=storage)
idx.reserve(rows)
idx.add_synthetic(rows, seed=0)
qi = evs.IndexFlatIP(d)
qi.add_synthetic(nq, seed=1)
xq = torch.from_numpy(qi.reconstruct_n(0, nq)).cuda()
for _ in range(reps):
    D, I = idx.search(xq, k)
torch.cuda.synchronize()
print("ok", D[0, :3].tolist(), I[0, :3].tolist())
