"""Generates the committed fixtures in this directory.  Run from the repo root:  python tests/golden/make_golden.py

The reference (oldapp.py) cannot be imported here (flask, clip and faiss are absent) and ships no
golden vectors, so these are SELF-GENERATED known-answer vectors:
  * hand_cases.json  -- tiny integer-valued cases whose dot products are exact in fp32 under ANY
    summation order, so the expected output follows from the published faiss heap semantics alone
    (strict-greater admission, (score,id) heap order, descending output, (-FLT_MAX,-1) padding)
    and from the canonical (score desc, id asc) order.  Expected values were derived by hand and
    are asserted below against both oracle implementations before being written.
  * c1_10k_512.json  -- BASELINE config 1 (10k x 512, nq=1, k=12) on the counter-based synthetic
    generator: ids, fp32 scores and fp64 CANON-32 scores from the C oracle, cross-checked against
    the numpy restatement; plus a checksum of the generated inputs.
  * index_flat_3x4.faiss -- byte-exact IndexFlatIP file for N=3, d=4 built from SURVEY.md 5.1.
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
from oracle import faiss_io, flat_ip_np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NEG = -float(np.finfo(np.float32).max)


def hand_cases():
    cases = []
    # 1. exact ties at the k boundary.  scores: ids 0,1,2 -> 2 ; id 3 -> 5 ; id 4 -> 2 ; id 5 -> 1
    xb = [[2, 0, 0, 0], [0, 2, 0, 0], [0, 0, 2, 0], [5, 0, 0, 0], [1, 1, 0, 0], [1, 0, 0, 0]]
    q = [[1, 1, 1, 0]]
    cases.append(dict(
        name="ties_at_boundary_k3", xb=xb, xq=q, k=3,
        # canonical: best score, then lowest ids among the tied 2's
        canon_I=[[3, 0, 1]], canon_D=[[5, 2, 2]],
        # faiss heap: heap {0,1,2}; id3 (5) evicts the root = smallest (score,id) = id 0; id4 (2) is
        # not strictly greater than the root (2) -> rejected; output descending (score,id): 3,2,1
        faiss_I=[[3, 2, 1]], faiss_D=[[5, 2, 2]]))
    # 2. duplicate rows (all scores equal): canonical keeps the lowest ids, ascending
    xb = [[1, 2, 3, 4]] * 5
    cases.append(dict(
        name="duplicates_k2", xb=xb, xq=[[1, 0, 0, 1]], k=2,
        canon_I=[[0, 1]], canon_D=[[5, 5]],
        faiss_I=[[1, 0]], faiss_D=[[5, 5]]))
    # 3. k > N: padding with (-FLT_MAX, -1)
    xb = [[1, 0], [0, 1], [3, 3]]
    cases.append(dict(
        name="k_greater_than_n", xb=xb, xq=[[1, 2]], k=5,
        canon_I=[[2, 1, 0, -1, -1]], canon_D=[[9, 2, 1, NEG, NEG]],
        faiss_I=[[2, 1, 0, -1, -1]], faiss_D=[[9, 2, 1, NEG, NEG]]))
    # 4. k = 1, N = 1
    cases.append(dict(
        name="k1_n1", xb=[[2, -1, 0.5]], xq=[[4, 2, 2]], k=1,
        canon_I=[[0]], canon_D=[[7]], faiss_I=[[0]], faiss_D=[[7]]))
    # 5. negative and zero scores, zero-vector row, two queries
    xb = [[0, 0, 0], [-1, 0, 0], [1, 0, 0], [0, -2, 0]]
    cases.append(dict(
        name="negatives_zero_row_two_queries", xb=xb, xq=[[1, 1, 0], [-1, 0, 0]], k=3,
        canon_I=[[2, 0, 1], [1, 0, 3]], canon_D=[[1, 0, -1], [1, 0, 0]],
        faiss_I=[[2, 0, 1], [1, 3, 0]], faiss_D=[[1, 0, -1], [1, 0, 0]]))
    # 6. d not a multiple of 32 and larger than 32 (CANON-32 tail lanes)
    d = 45
    xb = np.zeros((4, d)); xb[0, 44] = 3; xb[1, 31] = 2; xb[2, 32] = 4; xb[3, 0] = 1
    q = np.ones((1, d))
    cases.append(dict(
        name="d45_tail_lanes", xb=xb.tolist(), xq=q.tolist(), k=4,
        canon_I=[[2, 0, 1, 3]], canon_D=[[4, 3, 2, 1]], faiss_I=[[2, 0, 1, 3]], faiss_D=[[4, 3, 2, 1]]))
    return cases


def main():
    cases = hand_cases()
    for c in cases:
        xb = np.array(c["xb"], np.float32)
        xq = np.array(c["xq"], np.float32)
        k = c["k"]
        D, I = oracle.canon_search(xq, xb, k)
        assert I.tolist() == c["canon_I"], (c["name"], "canon", I)
        assert np.array_equal(D, np.array(c["canon_D"], np.float32)), (c["name"], D)
        D2, I2, _ = flat_ip_np.canon_search_np(xq, xb, k)
        assert np.array_equal(I2, I) and np.array_equal(D2, D), c["name"]
        Df, If = oracle.faiss_seq_search(xq, xb, k)
        assert If.tolist() == c["faiss_I"], (c["name"], "faiss", If)
        assert np.array_equal(Df, np.array(c["faiss_D"], np.float32)), (c["name"], Df)
        Dp, Ip = flat_ip_np.faiss_heap_search_py(xq, xb, k)
        assert np.array_equal(Ip, If) and np.array_equal(Dp, Df), c["name"]
    with open(os.path.join(HERE, "hand_cases.json"), "w") as f:
        json.dump(cases, f, indent=1)

    # BASELINE config 1
    n, d, k = 10_000, 512, 12
    xb = oracle.synth_fill(n, d, seed=0)
    xq = oracle.synth_fill(1, d, seed=1)
    assert np.array_equal(xb[:64], flat_ip_np.l2_normalize_np(flat_ip_np.synth_raw_np(64, d, 0)))
    D, I, D64 = oracle.canon_search(xq, xb, k, return_f64=True)
    Dn, In, D64n = flat_ip_np.canon_search_np(xq, xb, k)
    assert np.array_equal(I, In) and np.array_equal(D64, D64n) and np.array_equal(D, Dn)
    Df, If = oracle.faiss_seq_search(xq, xb, k)
    c1 = dict(n=n, d=d, k=k, seed_xb=0, seed_xq=1,
              xb_sha256=hashlib.sha256(xb.tobytes()).hexdigest(),
              xq_sha256=hashlib.sha256(xq.tobytes()).hexdigest(),
              canon_I=I.tolist(), canon_D_f32_hex=[float(v).hex() for v in D[0]],
              canon_D_f64_hex=[float(v).hex() for v in D64[0]],
              faiss_I=If.tolist(), faiss_D_f32_hex=[float(v).hex() for v in Df[0]])
    with open(os.path.join(HERE, "c1_10k_512.json"), "w") as f:
        json.dump(c1, f, indent=1)

    # index.faiss fixture: 45-byte header + 48-byte payload, written out field by field
    xb = np.array([[1.0, 2.0, 3.0, 4.0], [-0.5, 0.25, 0.0, 8.0], [1e-3, -1e3, 0.1, 0.2]], np.float32)
    blob = bytearray()
    blob += b"IxFI"
    blob += (4).to_bytes(4, "little", signed=True)
    blob += (3).to_bytes(8, "little", signed=True)
    blob += (1 << 20).to_bytes(8, "little", signed=True)
    blob += (1 << 20).to_bytes(8, "little", signed=True)
    blob += b"\x01"
    blob += (0).to_bytes(4, "little", signed=True)
    blob += (12).to_bytes(8, "little", signed=False)
    blob += xb.astype("<f4").tobytes()
    assert len(blob) == 45 + 48 and bytes(blob) == faiss_io.pack_index_flat(xb)
    with open(os.path.join(HERE, "index_flat_3x4.faiss"), "wb") as f:
        f.write(bytes(blob))
    print("golden fixtures written")


if __name__ == "__main__":
    main()
