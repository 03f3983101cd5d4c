"""torchrun --nproc-per-node G scripts/check_sharded.py : row-sharded search over G GPUs must equal the CPU
oracle's canonical ranking bit for bit, for f32 and bf16 storage, through both exchanges: the NCCL
all-gather + merge kernel, and the peer-store exchange (finalise kernel writes every rank's slot over
NVLink, flag-waiting merge kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import evo_ssearch_b200 as evs  # noqa: E402
import oracle  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
CASES = ((1_000_003, 512, 48, (1, 16)), (200_001, 768, 12, (1, 5)), (5, 512, 12, (3,)))
if os.environ.get("EVS_CHECK_LIGHT"):  # many ranks share the host cores for the oracle: keep it small
    CASES = ((300_007, 512, 48, (1, 16)), (5, 512, 12, (3,)))
for n, d, k, nqs in CASES:
    xb = oracle.synth_fill(n, d, 0)
    if n > 10:
        xb[n - 1] = xb[0]  # exact tie across the first and last shard
    xq = oracle.synth_fill(max(nqs), d, 1)
    for storage, exchange in (("f32", "nccl"), ("bf16", "nccl"), ("f32", "peer"), ("bf16", "peer")):
        sh = evs.ShardedIndexFlatIP(d, device=local, storage=storage, exchange=exchange, exchange_max_nq=64)
        sh.add(xb)
        for nq in nqs:
            for rep in range(3 if exchange == "peer" else 1):  # both slot generations and their reuse
                D, I = sh.search(xq[:nq], k)
                Dr, Ir = oracle.canon_search(xq[:nq], xb, k)
                good = bool(np.array_equal(I, Ir) and np.array_equal(D, Dr))
                ok = ok and good
            if rank == 0:
                print(f"n={n} d={d} k={k} nq={nq} storage={storage} exchange={exchange} world={world}: "
                      f"{'OK' if good else 'MISMATCH'}", flush=True)
        if sh._px is not None:
            timed_out, searches = sh._px.status()
            ok = ok and not timed_out and searches > 0
        # generated-in-place shards equal host-fed shards
        sh2 = evs.ShardedIndexFlatIP(d, device=local, storage=storage)
        sh2.add_synthetic(n, seed=0)
        lo, hi = evs.shard_bounds(n, world, rank)
        if hi > lo:
            ref = oracle.synth_fill(n, d, 0)
            ok = ok and bool(np.array_equal(sh2.local.reconstruct_n(0, hi - lo), ref[lo:hi]))
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
if rank == 0:
    print("SHARDED_PARITY_OK" if int(t.item()) == 1 else "SHARDED_PARITY_FAILED", flush=True)
sys.exit(0 if int(t.item()) == 1 else 1)
