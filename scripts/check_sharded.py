"""torchrun --nproc-per-node G scripts/check_sharded.py : row-sharded search over G GPUs must equal the CPU
oracle's canonical ranking bit for bit, for f32 and bf16 storage, through both exchanges: the NCCL
all-gather + merge kernel, and the peer-store exchange (finalise kernel writes every rank's slot over
NVLink, flag-waiting merge kernel).  Also: the device-side certification guard through the sharded API (planted
near-ties), the shard loader (every rank reads only its block of index.faiss) and the sharded load_index.

Prints one line per case and SHARDED_PARITY_OK / SHARDED_PARITY_FAILED; EVS_CHECK_LOG=<file> also appends the lines there
(rank 0), so that the multi-GPU evidence can be kept under profiles/."""
import os
import pickle
import sys
import tempfile
import time

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
LOG = os.environ.get("EVS_CHECK_LOG")


def say(msg):
    if rank == 0:
        print(msg, flush=True)
        if LOG:
            with open(LOG, "a") as f:
                f.write(msg + "\n")


def bcast_obj(obj):
    box = [obj]
    dist.broadcast_object_list(box, src=0)
    return box[0]


say(f"# check_sharded: world={world} gpu={torch.cuda.get_device_name(local)}")
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
            say(f"n={n} d={d} k={k} nq={nq} storage={storage} exchange={exchange} world={world}: {'OK' if good else 'MISMATCH'}")
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

# ---- the certification guard through the sharded API: 80 planted rows 2e-7 apart straddle rank 48 of query 0 -------------
n, d, k = 400_000, 512, 48
xb = oracle.synth_fill(n, d, 61)
q = oracle.synth_fill(6, d, 62)
rng = np.random.default_rng(9)
u = rng.standard_normal((80, d))
u -= (u @ q[0].astype(np.float64))[:, None] * q[0]
u /= np.linalg.norm(u, axis=1, keepdims=True)
a = 0.9 + 2e-7 * np.arange(80)
rows = 1000 + np.arange(80)  # all in shard 0: that shard's own top-48 is what the scan cannot resolve
xb[rows] = (a[:, None] * q[0] + np.sqrt(1 - a[:, None] ** 2) * u).astype(np.float32)
Dr, Ir = oracle.canon_search(q, xb, k)
for exchange in ("nccl", "peer"):
    for x3 in (1, 0):
        evs.set_option("x3", x3)
        sh = evs.ShardedIndexFlatIP(d, device=local, storage="f32", exchange=exchange, exchange_max_nq=64)
        sh.add(xb)
        D, I = sh.search(q, k)
        Dt, It = sh.search_tensor(torch.from_numpy(q).cuda(), k)
        good = bool(np.array_equal(I, Ir) and np.array_equal(D, Dr) and np.array_equal(It.cpu().numpy(), Ir)
                    and np.array_equal(Dt.cpu().numpy(), Dr))
        reruns, uncert = sh.local.guard_stats()
        ok = ok and good and uncert == 0
        say(f"guard: planted near-ties, 6 fp32 queries, exchange={exchange} x3={x3} world={world}: {'OK' if good else 'MISMATCH'} "
            f"(rank 0: {reruns} device re-runs, {uncert} uncertified)")
evs.set_option("x3", 0)

# ---- a rank that fails a collective search: its peers must report it, never merge without that shard ---------------------
if world >= 2:
    n, d, k = 200_003, 512, 48
    xb = oracle.synth_fill(n, d, 33)
    xq = oracle.synth_fill(2, d, 34)
    Dr, Ir = oracle.canon_search(xq, xb, k)
    sh = evs.ShardedIndexFlatIP(d, device=local, storage="f32", exchange="peer", exchange_max_nq=64)
    sh.add(xb)
    D, I = sh.search(xq[:1], k)
    good = bool(np.array_equal(I, Ir[:1]))
    for api in ("host", "device"):
        if rank == world - 1:
            evs.set_option("exchange_fail_next", 1)  # this rank's next exchange search fails after taking its sequence number
        raised = False
        try:
            if api == "host":
                D, I = sh.search(xq[:1], k)  # host entry point: the failure of THIS search is reported by THIS call on every rank
            else:
                Dt, It = sh.search_tensor(torch.from_numpy(xq[:1]).cuda(), k)  # asynchronous: the failing rank raises now ...
                torch.cuda.synchronize()
                D, I = Dt.cpu().numpy(), It.cpu().numpy()
        except evs.EvsError:
            raised = True
        if api == "host":
            good = good and raised
        else:
            # ... its peers got padding (never a merge of stale entries) and raise on their next call
            good = good and (raised if rank == world - 1 else bool((I == -1).all()))
            raised2 = False
            try:
                sh.search_tensor(torch.from_numpy(xq[:1]).cuda(), k)
                torch.cuda.synchronize()
            except evs.EvsError:
                raised2 = True
            good = good and (raised2 if rank != world - 1 else True)
        # the ranks are still in step: the following searches are right again
        D, I = sh.search(xq, k)
        good = good and bool(np.array_equal(I, Ir) and np.array_equal(D, Dr))
    flags = [None] * world
    dist.all_gather_object(flags, good)
    ok = ok and all(flags)
    say(f"failure injection (rank {world - 1} fails one host and one device search): peers report it, later searches correct: "
        f"{'OK' if all(flags) else 'MISMATCH ' + str(flags)}")

# ---- shard loader: rank 0 writes a single-GPU index.faiss; every rank streams only its own block ---------------------------
n, d, k = 600_011, 512, 48
tmp = bcast_obj(tempfile.mkdtemp(prefix="evs_shard_") if rank == 0 else None)
folder = os.path.join(tmp, "photos")
xb = oracle.synth_fill(n, d, 5)
xq = oracle.synth_fill(4, d, 6)
if rank == 0:
    os.makedirs(os.path.join(folder, ".clip_index"))
    whole = evs.IndexFlatIP(d, device=local)
    whole.add(xb)
    evs.write_index(whole, os.path.join(folder, ".clip_index", "index.faiss"))
    with open(os.path.join(folder, ".clip_index", "paths.pkl"), "wb") as f:
        pickle.dump([f"img{i}.jpg" for i in range(n)], f)
    with open(os.path.join(folder, ".clip_index", "metadata.pkl"), "wb") as f:
        pickle.dump([{"path": f"img{i}.jpg", "mtime": float(i), "size": i} for i in range(n)], f)
    Dw, Iw = whole.search(xq, k)
    del whole
dist.barrier()
Dr, Ir = oracle.canon_search(xq, xb, k)
fname = os.path.join(folder, ".clip_index", "index.faiss")
for exchange in ("peer", "nccl"):
    t0 = time.perf_counter()
    sh = evs.ShardedIndexFlatIP.read_index(fname, device=local, exchange=exchange, exchange_max_nq=64)
    dt = time.perf_counter() - t0
    lo, hi = evs.shard_bounds(n, world, rank)
    good = sh.ntotal == n and sh.local.ntotal == hi - lo and sh.local.id_base == lo
    D, I = sh.search(xq, k)
    good = bool(good and np.array_equal(I, Ir) and np.array_equal(D, Dr))
    ok = ok and good
    say(f"shard loader: {n}x{d} index.faiss over {world} ranks, exchange={exchange}: {'OK' if good else 'MISMATCH'}; rank 0 read "
        f"{(hi - lo) * d * 4 / 1e6:.0f} MB in {dt * 1e3:.0f} ms = {(hi - lo) * d * 4 / dt / 1e9:.2f} GB/s (page cache -> pinned -> HBM)")
    # a sharded index writes back the same bytes
    if exchange == "peer":
        out = os.path.join(tmp, "rewritten.faiss")
        sh.write_index(out)
        dist.barrier()
        if rank == 0:
            same = open(out, "rb").read() == open(fname, "rb").read()
            ok = ok and same
            say(f"sharded write_index is byte-identical to the single-GPU file: {'OK' if same else 'MISMATCH'}")
# the application's entry point: load_index shards automatically under torch.distributed, and keeps the shards resident
os.environ["EVS_EXCHANGE"] = "peer"
index, paths, meta = evs.load_index(folder)
good = index is not None and hasattr(index, "local") and index.ntotal == n and len(paths) == n
index2, _, _ = evs.load_index(folder)
good = good and index2 is index


class _Enc:
    def get_text_embedding(self, text):
        return xq[int(text)]


res = evs.search_text(folder, "2", _Enc(), limit=12)
good = bool(good and res is not None and [r["path"] for r in res] == [f"img{i}.jpg" for i in Ir[2][:12]])
ok = ok and good
say(f"load_index(folder) -> ShardedIndexFlatIP resident on {world} GPUs, search_text equals the oracle: {'OK' if good else 'MISMATCH'}")
evs.evict_index()

t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
say("SHARDED_PARITY_OK" if int(t.item()) == 1 else "SHARDED_PARITY_FAILED")
sys.exit(0 if int(t.item()) == 1 else 1)
