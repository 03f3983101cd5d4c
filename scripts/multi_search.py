"""Several searches in one process (profiling target for an ncu launch list).
Usage: multi_search.py rows:d:storage:nq[:k] [...]   -- each case: 2 warm-up searches, then 2 more."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402

cur = None
idx = None
for spec in sys.argv[1:]:
    f = spec.split(":")
    rows, d, storage, nq = int(f[0]), int(f[1]), f[2], int(f[3])
    k = int(f[4]) if len(f) > 4 else 48
    if cur != (rows, d, storage):
        del idx
        torch.cuda.empty_cache()
        idx = evs.IndexFlatIP(d, storage=storage)
        idx.reserve(rows)
        idx.add_synthetic(rows, seed=0)
        cur = (rows, d, storage)
    qi = evs.IndexFlatIP(d)
    qi.add_synthetic(nq, seed=1)
    xq = torch.from_numpy(qi.reconstruct_n(0, nq)).cuda()
    del qi
    for _ in range(4):
        D, I = idx.search(xq, k)
    torch.cuda.synchronize()
    print("ok", spec, I[0, :3].tolist(), flush=True)
