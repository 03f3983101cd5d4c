"""torchrun --nproc-per-node G scripts/exchange_probe.py [--rows 10000000]: where does a sharded single-query search spend
its time on each rank?  Per-CTA %globaltimer stamps (option scan_clock) of the one-launch search: scan loop, the last CTA's
finalise phases, and the time between "partial stored to the peers" and "merged" (= waiting for the slowest rank + merge).
GPU clocks of different devices are not synchronised: only differences taken on one device are printed."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--reps", type=int, default=40)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sh = evs.ShardedIndexFlatIP(a.dim, device=local, exchange="peer", exchange_max_nq=64, exchange_max_k=48)
sh.add_synthetic(a.rows, seed=0)
qi = evs.IndexFlatIP(a.dim, device=local)
qi.add_synthetic(64, seed=1)
q = torch.from_numpy(qi.reconstruct_n(0, 64)).to(dev)
for i in range(10):
    sh.search_tensor(q[i:i + 1], 48)
torch.cuda.synchronize()
dist.barrier()
# free-running timing (what bench.py measures)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(200):
    sh.search_tensor(q[i % 64:i % 64 + 1], 48)
e1.record()
torch.cuda.synchronize()
free_ms = e0.elapsed_time(e1) / 200
dist.barrier()
evs.set_option("scan_clock", 1)
rows = []
for rep in range(a.reps):
    # free-running burst; the stamps left behind are those of its LAST search (steady state, ranks in lock step through the flags)
    for i in range(20):
        sh.search_tensor(q[(rep + i) % 64:(rep + i) % 64 + 1], 48)
    torch.cuda.synchronize()
    c = sh.local.scan_clocks().astype(np.int64)
    st = sh.local.last_cta_stamps.astype(np.int64)
    fs = sh.local.finalize_stamps.astype(np.int64)
    t0 = c[:, 0].min()
    rows.append(dict(span=(st[1] - t0) / 1e3, loop_max=(c[:, 1].max() - t0) / 1e3, loop_med=(np.median(c[:, 1]) - t0) / 1e3,
                     ticket=(st[0] - t0) / 1e3, written=(fs[4] - t0) / 1e3, wait_merge=(st[1] - fs[4]) / 1e3,
                     start_skew=(c[:, 0].max() - t0) / 1e3))
    dist.barrier()
evs.set_option("scan_clock", 0)
med = {k: round(float(np.median([r[k] for r in rows])), 1) for k in rows[0]}
med["rank"] = rank
med["free_running_ms"] = round(free_ms, 4)
out = [None] * world
dist.all_gather_object(out, med)
if rank == 0:
    for m in out:
        print(json.dumps(m), flush=True)
dist.destroy_process_group()
