"""index.faiss loader / writer throughput: write_index of a synthetic index, then read_index with 1, 2, 4, 8, 16 threads per
chunk (option io_threads; the file sits in the page cache, as it does for an index the application has just saved or loaded
before), each checked against the source rows.   python scripts/loader_probe.py [--rows 4000000] [--dir /dev/shm]"""
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
ap.add_argument("--rows", type=int, default=4_000_000)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--dir", default="/tmp")
ap.add_argument("--write-threads", default="1,8")
ap.add_argument("--read-threads", default="1,2,4,8,16,0,0")
a = ap.parse_args()
path = os.path.join(a.dir, "evs_loader_probe.faiss")
src = evs.IndexFlatIP(a.dim)
src.reserve(a.rows)
src.add_synthetic(a.rows, seed=3)
gb = a.rows * a.dim * 4 / 1e9
probe_rows = [0, 1, a.rows // 3, a.rows - 1]
want = np.stack([src.reconstruct(i) for i in probe_rows])
xq = torch.from_numpy(src.reconstruct_n(7, 5)).cuda()
D0, I0 = src.search(xq, 12)
for thr in [int(x) for x in a.write_threads.split(",")]:
    evs.set_option("io_threads", thr)
    t0 = time.perf_counter()
    evs.write_index(src, path)
    dt = time.perf_counter() - t0
    print(json.dumps({"op": "write_index", "io_threads": thr, "gb": round(gb, 3), "s": round(dt, 3), "gb_per_s": round(gb / dt, 2)}), flush=True)
del src
torch.cuda.empty_cache()
for thr in [int(x) for x in a.read_threads.split(",")]:
    evs.set_option("io_threads", thr)
    t0 = time.perf_counter()
    idx = evs.read_index(path)
    dt = time.perf_counter() - t0
    got = np.stack([idx.reconstruct(i) for i in probe_rows])
    D, I = idx.search(xq, 12)
    ok = bool(np.array_equal(got, want) and torch.equal(I, I0) and torch.equal(D, D0) and idx.ntotal == a.rows)
    print(json.dumps({"op": "read_index", "io_threads": thr, "gb": round(gb, 3), "s": round(dt, 3), "gb_per_s": round(gb / dt, 2),
                      "equals_source": ok}), flush=True)
    del idx
    torch.cuda.empty_cache()
# one shard's slice (rows [n/2, n)), as the sharded loader reads it
evs.set_option("io_threads", 0)
t0 = time.perf_counter()
from evo_ssearch_b200.index import read_index_rows  # noqa: E402
part, _ = read_index_rows(path, a.rows // 2, a.rows)
dt = time.perf_counter() - t0
print(json.dumps({"op": "read_index_rows [n/2, n)", "io_threads": 0, "gb": round(gb / 2, 3), "s": round(dt, 3), "gb_per_s": round(gb / 2 / dt, 2),
                  "rows": part.ntotal}), flush=True)
os.remove(path)
