"""CPU, world_size 2 over gloo: the host side of the row-sharded search (layout, one all-gather,
replicated merge) with the oracle standing in for the per-GPU engine and the merge kernel."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NEG64 = -np.finfo(np.float64).max


class OracleShard:
    """Duck-types IndexFlatIP for ShardedIndexFlatIP using the CPU oracle (tests only)."""

    torch_device = torch.device("cpu")

    def __init__(self, d):
        self.d, self.id_base, self.xb = d, 0, np.zeros((0, d), np.float32)

    def add(self, x):
        self.xb = np.concatenate([self.xb, np.asarray(x, np.float32)])

    def reserve(self, n):
        pass

    def add_synthetic(self, n, seed, normalize=True):
        import oracle
        self.add(oracle.synth_fill(n, self.d, seed, row_base=self.id_base + self.xb.shape[0], normalize=normalize))

    @property
    def ntotal(self):
        return self.xb.shape[0]

    def reconstruct_n(self, n0, ni):
        return self.xb[n0:n0 + ni]

    def search_partial(self, xq, k):
        import oracle
        q = xq.numpy()
        if self.xb.shape[0] == 0:
            S = np.full((q.shape[0], k), NEG64)
            I = np.full((q.shape[0], k), -1, np.int64)
        else:
            _, I, S = oracle.canon_search(q, self.xb, k, id_base=self.id_base, return_f64=True)
            S = np.where(I >= 0, S, NEG64)
        return torch.from_numpy(S.copy()), torch.from_numpy(I.copy())


def oracle_merge(scores, ids, k):
    """Same contract as evs_merge_partials_dev, on CPU tensors (possibly strided views)."""
    G, nq, _ = scores.shape
    D = np.full((nq, k), -np.finfo(np.float32).max, np.float32)
    I = np.full((nq, k), -1, np.int64)
    s, i = scores.numpy(), ids.numpy()
    for q in range(nq):
        ss, ii = s[:, q].reshape(-1), i[:, q].reshape(-1)
        keep = ii >= 0
        ss, ii = ss[keep], ii[keep]
        order = np.lexsort((ii, -ss))[:k]
        D[q, :order.size] = ss[order].astype(np.float32)
        I[q, :order.size] = ii[order]
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, n, d, nq, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import evo_ssearch_b200 as evs
        import oracle
        xb = oracle.synth_fill(n, d, 21)
        xb[n - 1] = xb[1]  # an exact tie across the shard boundary
        xq = oracle.synth_fill(nq, d, 22)
        sh = evs.ShardedIndexFlatIP(d, local_index=OracleShard(d), merge=oracle_merge)
        sh.add(xb)
        lo, hi = evs.shard_bounds(n, world, rank)
        assert sh.local.id_base == lo and sh.local.xb.shape[0] == hi - lo and sh.ntotal == n
        D, I = sh.search(xq, k)
        Dr, Ir = oracle.canon_search(xq, xb, k)
        ok = bool(np.array_equal(I, Ir) and np.array_equal(D, Dr))
        # synthetic fill: every rank generates its own block from global ids
        sh2 = evs.ShardedIndexFlatIP(d, local_index=OracleShard(d), merge=oracle_merge)
        sh2.add_synthetic(n, 21)
        ok = ok and bool(np.array_equal(sh2.local.xb, oracle.synth_fill(n, d, 21)[lo:hi]))
        # sharded write_index: header by rank 0, every rank its own block -> the single-GPU file, byte for byte
        from oracle import faiss_io
        path = os.path.join(os.environ["EVS_TEST_TMP"], f"sharded_{n}.faiss")
        sh.write_index(path)
        dist.barrier()
        ok = ok and open(path, "rb").read() == faiss_io.pack_index_flat(xb)
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n,k", [(1001, 48), (5, 12)])  # uneven shards; shards with fewer than k rows
def test_two_rank_sharded_search_equals_single(n, k, tmp_path):
    world = 2
    os.environ["EVS_TEST_TMP"] = str(tmp_path)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, 64, 3, k, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
