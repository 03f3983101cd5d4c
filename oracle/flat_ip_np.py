"""Independent numpy / pure-Python restatements used to pin ``flat_ip_oracle.c``.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see ``flat_ip_oracle.c``): faiss-cpu, which owns the reference's arithmetic
(oldapp.py:87-88, :2005, :2112; requirements.txt:6), is not available in this image.
"""
from __future__ import annotations

import numpy as np

FLT_MAX = float(np.finfo(np.float32).max)


def canon_scores_np(q: np.ndarray, xb: np.ndarray, chunk: int = 65536) -> np.ndarray:
    """CANON-32 fp64 scores of all rows of ``xb`` against ``q`` (definition in flat_ip_oracle.c)."""
    q = np.ascontiguousarray(q, np.float32).reshape(-1)
    xb = np.ascontiguousarray(xb, np.float32)
    n, d = xb.shape
    dp = (d + 31) // 32 * 32
    q64 = np.zeros(dp, np.float64)
    q64[:d] = q
    out = np.empty(n, np.float64)
    for lo in range(0, n, chunk):
        x = xb[lo:lo + chunk].astype(np.float64)
        if dp != d:
            x = np.concatenate([x, np.zeros((x.shape[0], dp - d))], axis=1)
        prod = (x * q64).reshape(x.shape[0], dp // 32, 32)  # exact products; padded ones add +0.0
        p = np.zeros((x.shape[0], 32), np.float64)
        for j in range(dp // 32):
            p = p + prod[:, j, :]
        lanes = np.arange(32)
        for off in (16, 8, 4, 2, 1):
            p = p + p[:, lanes ^ off]
        out[lo:lo + chunk] = p[:, 0]
    return out


def canon_search_np(xq: np.ndarray, xb: np.ndarray, k: int, id_base: int = 0):
    """(score desc, id asc) ranking of CANON-32 scores; pads with (-FLT_MAX, -1)."""
    xq = np.ascontiguousarray(xq, np.float32)
    nq = xq.shape[0]
    n = xb.shape[0]
    D = np.full((nq, k), -FLT_MAX, np.float32)
    I = np.full((nq, k), -1, np.int64)
    D64 = np.full((nq, k), -FLT_MAX, np.float64)
    for i in range(nq):
        s = canon_scores_np(xq[i], xb)
        order = np.lexsort((np.arange(n), -s))[:k]  # primary: -score ascending; ties: id ascending
        m = order.shape[0]
        D[i, :m] = s[order].astype(np.float32)
        D64[i, :m] = s[order]
        I[i, :m] = order + id_base
    return D, I, D64


def faiss_heap_search_py(xq: np.ndarray, xb: np.ndarray, k: int):
    """Pure-Python model of faiss's k-entry min-heap result handler over sequential fp32 dots.

    Only the heap *contents* matter for the result, so the heap is modelled as a plain list whose
    minimum under (score, id) order is the root: strict-greater admission, evict the root,
    descending (score, id) output, (-FLT_MAX, -1) padding.  Small inputs only.
    """
    xq = np.ascontiguousarray(xq, np.float32)
    xb = np.ascontiguousarray(xb, np.float32)
    nq, d = xq.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    neutral = np.float32(-FLT_MAX)
    for i in range(nq):
        heap = [(neutral, -1)] * k
        for j in range(xb.shape[0]):
            s = np.float32(0.0)
            for t in range(d):  # strict sequential fp32 accumulation, no FMA
                s = np.float32(s + np.float32(xq[i, t] * xb[j, t]))
            root = min(range(k), key=lambda r: (heap[r][0], heap[r][1]))
            if heap[root][0] < s:
                heap[root] = (s, j)
        real = sorted([h for h in heap if h[1] != -1], key=lambda h: (h[0], h[1]), reverse=True)
        real += [(neutral, -1)] * (k - len(real))
        D[i] = [h[0] for h in real]
        I[i] = [h[1] for h in real]
    return D, I


def bruteforce_f64(xq: np.ndarray, xb: np.ndarray, k: int):
    """Plain fp64 matmul + stable sort: a tolerance reference, not an order-exact one."""
    s = xq.astype(np.float64) @ xb.astype(np.float64).T
    n = xb.shape[0]
    I = np.full((xq.shape[0], k), -1, np.int64)
    D = np.full((xq.shape[0], k), -FLT_MAX, np.float64)
    for i in range(xq.shape[0]):
        order = np.lexsort((np.arange(n), -s[i]))[:k]
        I[i, :order.shape[0]] = order
        D[i, :order.shape[0]] = s[i, order]
    return D, I


def l2_normalize_np(x: np.ndarray) -> np.ndarray:
    """Same definition as ``orc_l2_normalize_f32``: fp64 CANON-32 sum of squares, fp32 division."""
    x = np.array(x, np.float32, copy=True)
    if x.ndim == 1:
        x = x[None]
    out = np.empty_like(x)
    with np.errstate(all="ignore"):
        for r in range(x.shape[0]):
            nrm = np.float32(np.sqrt(canon_scores_np(x[r], x[r:r + 1])[0]))
            out[r] = x[r] / nrm
    return out


def _splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_raw_np(n: int, d: int, seed: int, row_base: int = 0) -> np.ndarray:
    """Un-normalised synthetic integers (Irwin-Hall n=4 of 16-bit uniforms), as float32."""
    with np.errstate(over="ignore"):
        rows = (np.arange(n, dtype=np.uint64) + np.uint64(row_base))
        s0 = _splitmix64(np.array([np.uint64(seed) ^ np.uint64(0xD1B54A32D192ED03)], np.uint64))[0]
        hr = _splitmix64(s0 + rows * np.uint64(0x2545F4914F6CDD1D))
        h = _splitmix64(hr[:, None] + np.arange(d, dtype=np.uint64)[None, :])
    m = np.uint64(0xFFFF)
    s = ((h & m).astype(np.int64) + ((h >> np.uint64(16)) & m).astype(np.int64)
         + ((h >> np.uint64(32)) & m).astype(np.int64) + ((h >> np.uint64(48)) & m).astype(np.int64) - 131070)
    return s.astype(np.float32)
