"""CPU oracle for the evo-ssearch flat inner-product hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker / the timed CPU baseline.  The product
(``evo-ssearch_b200/``) never imports it.

PARITY UNPINNED: the reference's arithmetic for this path lives in the third-party ``faiss-cpu``
wheel (``/root/reference/requirements.txt:6``, ``>=1.7.4``, unpinned), which is absent from this image
and cannot be installed; the reference has no tests or golden vectors.  See ``flat_ip_oracle.c``.

Python surface (all numpy in / numpy out):

* :func:`faiss_seq_search`  -- restatement of faiss ``IndexFlatIP.search`` (oldapp.py:2005, :2112)
* :func:`canon_search`      -- the canonical fp64 ranking the CUDA path must match bit-exactly
* :func:`canon_scores`, :func:`dot_canon32`, :func:`l2_normalize`, :func:`synth_fill`
* :mod:`oracle.flat_ip_np`  -- independent numpy restatements used to pin the C code
* :mod:`oracle.faiss_io`    -- ``index.faiss`` byte layout (oldapp.py:98, :117)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

FLT_MAX = float(np.finfo(np.float32).max)


def build(force: bool = False) -> str:
    """Compile ``liboracle.so`` with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "flat_ip_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        i64p = ctypes.POINTER(ctypes.c_int64)
        i64 = ctypes.c_int64
        L.orc_faiss_seq_search.argtypes = [f32p, f32p, i64, i64, i64, i64, f32p, i64p, ctypes.c_int, ctypes.c_int]
        L.orc_allcores_search.argtypes = [f32p, f32p, i64, i64, i64, i64, f32p, i64p, ctypes.c_int]
        L.orc_canon_search.argtypes = [f32p, f32p, i64, i64, i64, i64, f32p, i64p, f64p, i64, ctypes.c_int]
        L.orc_canon_scores.argtypes = [f32p, f32p, i64, i64, f64p, ctypes.c_int]
        L.orc_l2_normalize_f32.argtypes = [f32p, i64, i64, ctypes.c_int]
        L.orc_synth_fill.argtypes = [f32p, i64, i64, ctypes.c_uint64, i64, ctypes.c_int, ctypes.c_int]
        L.orc_synth_value.argtypes = [ctypes.c_uint64, i64, i64]
        L.orc_synth_value.restype = ctypes.c_float
        L.orc_dot_canon32.argtypes = [f32p, f32p, i64]
        L.orc_dot_canon32.restype = ctypes.c_double
        L.orc_dot_f32_sequential.argtypes = [f32p, f32p, i64]
        L.orc_dot_f32_sequential.restype = ctypes.c_float
        L.orc_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _check_2d(xq, xb):
    xq = _f32(xq)
    xb = _f32(xb)
    if xq.ndim != 2 or xb.ndim != 2:
        raise ValueError("xq and xb must be 2-D")
    if xq.shape[1] != xb.shape[1]:
        raise AssertionError("dimension mismatch")  # faiss wrapper: assert d == self.d
    return xq, xb


def faiss_seq_search(xq, xb, k: int, simd: bool = False, nthreads: int = 0):
    """faiss ``IndexFlatIP.search`` restated: (D float32[nq,k] descending, I int64[nq,k])."""
    xq, xb = _check_2d(xq, xb)
    assert k > 0  # faiss python wrapper asserts k > 0
    nq, d = xq.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    rc = lib().orc_faiss_seq_search(_p(xq, ctypes.c_float), _p(xb, ctypes.c_float), d, nq, xb.shape[0], k,
                                    _p(D, ctypes.c_float), _p(I, ctypes.c_int64), int(simd), nthreads)
    if rc:
        raise RuntimeError(f"orc_faiss_seq_search failed: {rc}")
    return D, I


def allcores_search(xq, xb, k: int, nthreads: int = 0):
    """Same scan with database rows split over all host threads (CPU baseline, not faiss's threading)."""
    xq, xb = _check_2d(xq, xb)
    nq, d = xq.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    rc = lib().orc_allcores_search(_p(xq, ctypes.c_float), _p(xb, ctypes.c_float), d, nq, xb.shape[0], k,
                                   _p(D, ctypes.c_float), _p(I, ctypes.c_int64), nthreads)
    if rc:
        raise RuntimeError(f"orc_allcores_search failed: {rc}")
    return D, I


def canon_search(xq, xb, k: int, id_base: int = 0, nthreads: int = 0, return_f64: bool = False):
    """Canonical ranking: CANON-32 fp64 scores, (score desc, id asc), fp32 scores out."""
    xq, xb = _check_2d(xq, xb)
    assert k > 0
    nq, d = xq.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    D64 = np.empty((nq, k), np.float64)
    rc = lib().orc_canon_search(_p(xq, ctypes.c_float), _p(xb, ctypes.c_float), d, nq, xb.shape[0], k,
                                _p(D, ctypes.c_float), _p(I, ctypes.c_int64), _p(D64, ctypes.c_double),
                                id_base, nthreads)
    if rc:
        raise RuntimeError(f"orc_canon_search failed: {rc}")
    return (D, I, D64) if return_f64 else (D, I)


def canon_scores(q, xb, nthreads: int = 0) -> np.ndarray:
    """CANON-32 fp64 score of every database row against one query."""
    q = _f32(q).reshape(-1)
    xb = _f32(xb)
    assert xb.shape[1] == q.shape[0]
    out = np.empty(xb.shape[0], np.float64)
    rc = lib().orc_canon_scores(_p(q, ctypes.c_float), _p(xb, ctypes.c_float), q.shape[0], xb.shape[0],
                                _p(out, ctypes.c_double), nthreads)
    if rc:
        raise RuntimeError(f"orc_canon_scores failed: {rc}")
    return out


def dot_canon32(x, q) -> float:
    x = _f32(x).reshape(-1)
    q = _f32(q).reshape(-1)
    return float(lib().orc_dot_canon32(_p(x, ctypes.c_float), _p(q, ctypes.c_float), x.shape[0]))


def l2_normalize(x, nthreads: int = 0) -> np.ndarray:
    """Row-wise ``x / ||x||`` (oldapp.py:35/43/51), fp64 CANON-32 sum of squares, no epsilon."""
    x = np.array(x, dtype=np.float32, order="C", copy=True)
    if x.ndim == 1:
        x = x[None, :]
    with np.errstate(all="ignore"):
        rc = lib().orc_l2_normalize_f32(_p(x, ctypes.c_float), x.shape[0], x.shape[1], nthreads)
    if rc:
        raise RuntimeError(f"orc_l2_normalize_f32 failed: {rc}")
    return x


def synth_fill(n: int, d: int, seed: int, row_base: int = 0, normalize: bool = True, nthreads: int = 0):
    """Counter-based synthetic embeddings; bit-identical to the CUDA generator (evs_synth_fill)."""
    out = np.empty((n, d), np.float32)
    rc = lib().orc_synth_fill(_p(out, ctypes.c_float), n, d, seed, row_base, int(normalize), nthreads)
    if rc:
        raise RuntimeError(f"orc_synth_fill failed: {rc}")
    return out
