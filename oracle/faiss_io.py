"""``index.faiss`` byte layout for ``IndexFlatIP`` -- pure-Python restatement.  TEST INFRASTRUCTURE ONLY.

Restates what ``faiss.write_index`` / ``faiss.read_index`` (oldapp.py:98, :117) put on disk for a flat
index (faiss impl/index_write.cpp ``write_index_header`` + ``IxFI`` branch),
from knowledge of the upstream source.  PARITY UNPINNED: no faiss-written file exists in this image
to check the 45-byte header against.

  off  size  field
    0     4  fourcc  b"IxFI" (inner product)          reader also accepts b"IxF2" (L2) and b"IxFl"
    4     4  d            int32
    8     8  ntotal       int64
   16     8  dummy        int64 = 1 << 20
   24     8  dummy        int64 = 1 << 20
   32     1  is_trained   uint8 = 1
   33     4  metric_type  int32 (0 = inner product, 1 = L2)
   37     8  count        uint64 = ntotal * d  (number of float32 values)
   45  4*N*d payload      float32, row-major, little-endian
"""
from __future__ import annotations

import struct

import numpy as np

HEADER = struct.Struct("<4siqqqBiQ")
assert HEADER.size == 45

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


def pack_index_flat(xb: np.ndarray, fourcc: bytes = b"IxFI", metric: int = METRIC_INNER_PRODUCT) -> bytes:
    xb = np.ascontiguousarray(xb, dtype="<f4")
    n, d = xb.shape
    return HEADER.pack(fourcc, d, n, 1 << 20, 1 << 20, 1, metric, n * d) + xb.tobytes()


def write_index_flat(path: str, xb: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(pack_index_flat(xb))


def parse_index_flat(buf: bytes):
    """Returns (d, ntotal, metric, xb float32[N,d]); raises ValueError the way faiss raises."""
    if len(buf) < HEADER.size:
        raise ValueError("truncated header")
    fourcc, d, n, _d1, _d2, trained, metric, count = HEADER.unpack_from(buf, 0)
    if fourcc not in (b"IxFI", b"IxF2", b"IxFl"):
        raise ValueError(f"unsupported index type {fourcc!r}")
    if d <= 0 or n < 0:
        raise ValueError("bad header")
    if count >= (1 << 40):
        raise ValueError("vector too large")
    if count != n * d:
        raise ValueError("codes size mismatch")
    if len(buf) < HEADER.size + 4 * count:
        raise ValueError("truncated payload")
    xb = np.frombuffer(buf, dtype="<f4", count=count, offset=HEADER.size).reshape(n, d)
    return d, n, metric, xb


def read_index_flat(path: str):
    with open(path, "rb") as f:
        return parse_index_flat(f.read())
