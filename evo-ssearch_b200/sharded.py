"""Row-sharded flat inner-product index: one process per GPU, one small exchange per query batch.

SURVEY.md section 8(e): top-k over a union is the top-k of the per-part top-k, so database rows are
split into contiguous blocks -- rank g owns rows ``[g*ceil(N/G), min(N, (g+1)*ceil(N/G)))`` with global
ids = local row + block start -- queries are replicated, every rank scans its block and finalises its
own k best in the canonical fp64 order, and ONE all-gather (NCCL over NVLink; ``nq*k*16`` bytes per
rank) followed by a replicated G-way merge kernel gives every rank the final ``(D, I)``.  Because each
shard's scores are produced by the same canonical arithmetic, the merged result is bit-identical to
the single-GPU result for any G.

``exchange="peer"`` replaces the NCCL call and the separate merge by the library's peer-store exchange
(``evs_exchange_*``): the finalise kernel writes the shard's partial into every rank's mapped buffer over
NVLink and a flag-waiting merge kernel produces ``(D, I)`` -- two kernels fewer on the single-query
latency path, no host-launched collective.  The default ``"nccl"`` path is the portable one.

``torch.distributed`` is plumbing only (process group, the all-gather).  The local engine and the merge
are injectable so that the host logic can be exercised on CPU with ``gloo`` in the tests; the defaults
are the CUDA ones and nothing here falls back to CPU on its own.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows ``[lo, hi)`` of an ``n``-row database owned by ``rank`` of ``world`` (contiguous blocks)."""
    per = -(-n // world) if n > 0 else 0
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


class ShardedIndexFlatIP:
    """``IndexFlatIP`` over the GPUs of one box; call the same methods with the same arguments on every rank."""

    def __init__(self, d: int, *, group=None, device: Optional[int] = None, storage: Optional[str] = None,
                 local_index=None, merge: Optional[Callable] = None, exchange: str = "nccl",
                 exchange_max_nq: int = 1024, exchange_max_k: int = 48):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.d = int(d)
        self.is_trained = True
        if local_index is None:
            from .index import IndexFlatIP, merge_partials
            local_index = IndexFlatIP(d, device=device, storage=storage)
            merge = merge_partials if merge is None else merge
        assert merge is not None, "an injected local_index needs an injected merge"
        self.local = local_index
        self._merge = merge
        self._ntotal = 0
        assert exchange in ("nccl", "peer")
        self.exchange = exchange
        self._px = None
        if exchange == "peer" and self.world > 1:
            self._px = self._make_peer_exchange(exchange_max_nq, exchange_max_k)

    def _make_peer_exchange(self, max_nq: int, max_k: int):
        """Create this rank's buffer and swap the 64-byte IPC handles with one small all-gather."""
        import torch
        from .index import PeerExchange
        px = PeerExchange(self.local.device, self.rank, self.world, max_nq, max_k)
        dev = torch.device("cuda", self.local.device)
        mine = torch.frombuffer(bytearray(px.handle()), dtype=torch.uint8).to(dev)
        allh = torch.empty((self.world, mine.numel()), dtype=torch.uint8, device=dev)
        self._dist.all_gather_into_tensor(allh, mine, group=self.group)
        px.connect(allh.cpu().numpy().tobytes())
        self._dist.barrier(group=self.group)  # every rank has mapped every buffer before the first search
        return px

    # -- bookkeeping --------------------------------------------------------------------------
    @property
    def ntotal(self) -> int:
        return self._ntotal

    def _set_layout(self, n_total: int) -> Tuple[int, int]:
        if self._ntotal != 0:
            raise RuntimeError("a sharded index is filled once (contiguous row blocks); reset() first")
        lo, hi = shard_bounds(n_total, self.world, self.rank)
        self.local.id_base = lo
        return lo, hi

    # -- add ----------------------------------------------------------------------------------
    def add(self, x) -> None:
        """``x`` is the FULL ``(N, d)`` array, identical on every rank; each rank keeps its block."""
        x = np.asarray(x)
        n, d = x.shape
        assert d == self.d
        lo, hi = self._set_layout(n)
        if hi > lo:
            self.local.add(np.ascontiguousarray(x[lo:hi], dtype="float32"))
        self._ntotal = n

    def add_local(self, x_local, n_total: int) -> None:
        """Each rank passes only its own block (rows ``shard_bounds(n_total, world, rank)``)."""
        lo, hi = self._set_layout(n_total)
        assert x_local.shape[0] == hi - lo and x_local.shape[1] == self.d
        if hi > lo:
            self.local.add(x_local)
        self._ntotal = n_total

    def add_synthetic(self, n_total: int, seed: int, normalize: bool = True) -> None:
        """Every rank generates its own block on its GPU (rows are a function of the global id)."""
        lo, hi = self._set_layout(n_total)
        if hi > lo:
            self.local.reserve(hi - lo)
            self.local.add_synthetic(hi - lo, seed, normalize)
        self._ntotal = n_total

    # -- persistence ----------------------------------------------------------------------------
    @classmethod
    def read_index(cls, fname: str, *, group=None, device: Optional[int] = None, storage: Optional[str] = None,
                   exchange: str = "nccl", exchange_max_nq: int = 1024, exchange_max_k: int = 48) -> "ShardedIndexFlatIP":
        """``faiss.read_index(path)`` (oldapp.py:117) for a row-sharded index: every rank reads ONLY its own block of
        ``index.faiss`` -- bytes ``[45 + 4*d*lo, 45 + 4*d*hi)`` -- through the library's double-buffered pinned path
        (``evs_index_read_rows``); no rank ever holds the whole database.  Collective: call on every rank."""
        from .index import index_file_info, read_index_rows
        d, n = index_file_info(fname)
        self = cls(d, group=group, device=device, storage=storage, exchange=exchange, exchange_max_nq=exchange_max_nq,
                   exchange_max_k=exchange_max_k)
        lo, hi = shard_bounds(n, self.world, self.rank)
        shard, n_file = read_index_rows(fname, lo, hi, device=self.local.device, storage=self.local.storage)
        assert n_file == n and shard.id_base == lo and shard.ntotal == hi - lo
        self.local = shard  # the handle created by __init__ was empty
        self._ntotal = n
        return self

    def write_index(self, fname: str) -> None:
        """``faiss.write_index`` for a sharded index: rank 0 writes the header, every rank its own block, in rank order
        (one writer at a time; the file is byte-identical to the single-GPU one).  Collective."""
        import struct
        for r in range(self.world):
            if r == self.rank:
                lo, hi = shard_bounds(self._ntotal, self.world, self.rank)
                with open(fname, "wb" if r == 0 else "r+b") as f:
                    if r == 0:
                        f.write(b"IxFI" + struct.pack("<iqqqBiQ", self.d, self._ntotal, 1 << 20, 1 << 20, 1, 0,
                                                      self._ntotal * self.d))
                    f.seek(45 + 4 * self.d * lo)
                    step = max(1, (64 << 20) // (4 * self.d))
                    for r0 in range(0, hi - lo, step):
                        f.write(self.local.reconstruct_n(r0, min(step, hi - lo - r0)).tobytes())
            if self.world > 1:
                self._dist.barrier(group=self.group)

    # -- search -------------------------------------------------------------------------------
    def search_tensor(self, xq, k: int):
        """``xq``: ``(nq, d)`` tensor on this rank's device, identical on every rank -> tensors ``(D, I)``."""
        import torch
        assert k > 0
        nq = xq.shape[0]
        if self.world == 1 and hasattr(self.local, "_search_torch"):
            return self.local.search(xq, k)  # one shard: the finalise kernel emits (D, I) directly
        if self._px is not None and nq <= self._px.max_nq and k <= self._px.max_k:
            return self.local.search_exchange(self._px, xq, k)  # partial -> peers' slots -> merge, no NCCL call
        S, I = self.local.search_partial(xq, k)  # float64 [nq,k], int64 [nq,k] with global ids
        if self.world == 1:
            return self._merge(S.unsqueeze(0), I.unsqueeze(0), k)
        mine = torch.stack((S.view(torch.int64), I))  # [2,nq,k] -- one payload, one collective
        flat = torch.empty((self.world * 2, nq, k), dtype=torch.int64, device=mine.device)
        self._dist.all_gather_into_tensor(flat, mine, group=self.group)  # concatenation along dim 0
        allp = flat.view(self.world, 2, nq, k)
        # strided views of the gathered buffer: part stride 2*nq*k, each [nq,k] block dense
        return self._merge(allp.view(torch.float64)[:, 0], allp[:, 1], k)

    def search(self, x, k: int):
        """numpy in, numpy out, like ``IndexFlatIP.search``; every rank returns the same ``(D, I)``."""
        import torch
        x = np.ascontiguousarray(np.asarray(x), dtype="float32")
        n, d = x.shape
        assert d == self.d
        assert k > 0
        if self.world == 1 and hasattr(self.local, "_search_torch"):
            return self.local.search(x, k)  # one shard: the C ABI host entry point (evs_index_search)
        dev = getattr(self.local, "torch_device", None)
        if dev is None:
            dev = torch.device("cuda", self.local.device)
        if self._px is not None and n <= self._px.max_nq and k <= self._px.max_k:
            return self.local.search_exchange_host(self._px, x, k)  # all in the library: no torch on this path
        if dev.type != "cuda":  # injected CPU engine (tests)
            D, I = self.search_tensor(torch.from_numpy(x).to(dev), k)
            return D.cpu().numpy(), I.cpu().numpy()
        # pinned staging both ways, one synchronisation per call: H2D of the queries, search, D2H of (D, I)
        st = self._staging(n, d, k)
        st["q"][:n].copy_(torch.from_numpy(x))
        xq = st["q"][:n].to(dev, non_blocking=True)
        D, I = self.search_tensor(xq, k)
        st["D"][:n].copy_(D, non_blocking=True)
        st["I"][:n].copy_(I, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return st["D"][:n].numpy().copy(), st["I"][:n].numpy().copy()

    def _staging(self, n: int, d: int, k: int):
        import torch
        st = getattr(self, "_stage", None)
        if st is None or st["q"].shape[0] < n or st["D"].shape[1] != k:
            cap = max(n, 16)
            st = {"q": torch.empty((cap, d), dtype=torch.float32).pin_memory(),
                  "D": torch.empty((cap, k), dtype=torch.float32).pin_memory(),
                  "I": torch.empty((cap, k), dtype=torch.int64).pin_memory()}
            self._stage = st
        return st
