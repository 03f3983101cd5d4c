"""CPU: the oracle against the committed golden vectors and against its independent restatements."""
import hashlib
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
from oracle import faiss_io, flat_ip_np

NEG = -float(np.finfo(np.float32).max)


def _hand_cases(golden_dir):
    with open(os.path.join(golden_dir, "hand_cases.json")) as f:
        return json.load(f)


def test_hand_cases_c_and_numpy(golden_dir):
    for c in _hand_cases(golden_dir):
        xb = np.array(c["xb"], np.float32)
        xq = np.array(c["xq"], np.float32)
        k = c["k"]
        D, I = oracle.canon_search(xq, xb, k)
        assert I.tolist() == c["canon_I"], c["name"]
        assert np.array_equal(D, np.array(c["canon_D"], np.float32)), c["name"]
        D2, I2, _ = flat_ip_np.canon_search_np(xq, xb, k)
        assert np.array_equal(I2, I) and np.array_equal(D2, D), c["name"]
        Df, If = oracle.faiss_seq_search(xq, xb, k)
        assert If.tolist() == c["faiss_I"], c["name"]
        assert np.array_equal(Df, np.array(c["faiss_D"], np.float32)), c["name"]
        Dp, Ip = flat_ip_np.faiss_heap_search_py(xq, xb, k)
        assert np.array_equal(Ip, If) and np.array_equal(Dp, Df), c["name"]
        # the all-cores scan keeps the same set of (score) values
        Da, _ = oracle.allcores_search(xq, xb, k, nthreads=3)
        assert np.array_equal(Da, Df), c["name"]


def test_c1_golden(golden_dir):
    with open(os.path.join(golden_dir, "c1_10k_512.json")) as f:
        g = json.load(f)
    xb = oracle.synth_fill(g["n"], g["d"], g["seed_xb"])
    xq = oracle.synth_fill(1, g["d"], g["seed_xq"])
    assert hashlib.sha256(xb.tobytes()).hexdigest() == g["xb_sha256"]
    assert hashlib.sha256(xq.tobytes()).hexdigest() == g["xq_sha256"]
    D, I, D64 = oracle.canon_search(xq, xb, g["k"], return_f64=True)
    assert I.tolist() == g["canon_I"]
    assert [float(v).hex() for v in D[0]] == g["canon_D_f32_hex"]
    assert [float(v).hex() for v in D64[0]] == g["canon_D_f64_hex"]
    Df, If = oracle.faiss_seq_search(xq, xb, g["k"])
    assert If.tolist() == g["faiss_I"]
    assert [float(v).hex() for v in Df[0]] == g["faiss_D_f32_hex"]
    # faiss restatement and canonical ranking agree on this case; scores within 1e-5 relative
    assert If.tolist() == I.tolist()
    assert np.allclose(Df, D, rtol=1e-5, atol=0)
    # unit norm
    assert np.allclose(np.linalg.norm(xb[:100].astype(np.float64), axis=1), 1.0, atol=1e-6)


def test_canon_c_matches_numpy_random():
    rng = np.random.default_rng(3)
    for d in (1, 7, 32, 45, 512, 768):
        xb = rng.standard_normal((257, d)).astype(np.float32)
        xq = rng.standard_normal((3, d)).astype(np.float32)
        D, I, D64 = oracle.canon_search(xq, xb, 9, return_f64=True, nthreads=4)
        Dn, In, D64n = flat_ip_np.canon_search_np(xq, xb, 9)
        assert np.array_equal(I, In) and np.array_equal(D64, D64n) and np.array_equal(D, Dn)
        s = oracle.canon_scores(xq[0], xb)
        assert np.array_equal(s, flat_ip_np.canon_scores_np(xq[0], xb))
        assert np.allclose(s, xb.astype(np.float64) @ xq[0].astype(np.float64), rtol=0, atol=1e-12 * d)


def test_canon_independent_of_thread_count_and_id_base():
    rng = np.random.default_rng(4)
    xb = rng.standard_normal((1000, 64)).astype(np.float32)
    xb[500:520] = xb[10]  # duplicates -> exact ties
    xq = rng.standard_normal((2, 64)).astype(np.float32)
    ref = oracle.canon_search(xq, xb, 48, nthreads=1)
    for nt in (2, 3, 8):
        got = oracle.canon_search(xq, xb, 48, nthreads=nt)
        assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])
    D, I = oracle.canon_search(xq, xb, 48, id_base=1000)
    assert np.array_equal(I, ref[1] + 1000)


def test_faiss_restatement_vs_bruteforce_tolerance():
    xb = oracle.synth_fill(5000, 512, 11)
    xq = oracle.synth_fill(4, 512, 12)
    Df, If = oracle.faiss_seq_search(xq, xb, 48)
    Db, Ib = flat_ip_np.bruteforce_f64(xq, xb, 48)
    assert np.allclose(Df, Db, rtol=1e-5, atol=1e-7)
    for i in range(4):
        assert set(If[i]) == set(Ib[i]) or len(set(If[i]) ^ set(Ib[i])) <= 2  # near-ties at the boundary only
    Ds, Is = oracle.faiss_seq_search(xq, xb, 48, simd=True)
    assert np.allclose(Ds, Df, rtol=1e-5, atol=1e-7)


def test_normalize_and_synth_bitexact_restatements():
    raw = oracle.synth_fill(40, 100, 5, row_base=17, normalize=False)
    assert np.array_equal(raw, flat_ip_np.synth_raw_np(40, 100, 5, row_base=17))
    assert np.array_equal(oracle.l2_normalize(raw), flat_ip_np.l2_normalize_np(raw))
    # row_base shifts the stream: rows are a function of the global id
    a = oracle.synth_fill(10, 64, 9, row_base=0)
    b = oracle.synth_fill(6, 64, 9, row_base=4)
    assert np.array_equal(a[4:], b)
    z = oracle.l2_normalize(np.zeros((1, 8), np.float32))
    assert np.isnan(z).all()  # no epsilon, as oldapp.py:35


def test_index_faiss_fixture(golden_dir, tmp_path):
    blob = open(os.path.join(golden_dir, "index_flat_3x4.faiss"), "rb").read()
    assert len(blob) == 45 + 3 * 4 * 4
    d, n, metric, xb = faiss_io.parse_index_flat(blob)
    assert (d, n, metric) == (4, 3, 0)
    assert xb[1].tolist() == [-0.5, 0.25, 0.0, 8.0]
    assert faiss_io.pack_index_flat(xb) == blob
    p = tmp_path / "index.faiss"
    faiss_io.write_index_flat(str(p), xb)
    assert p.read_bytes() == blob
    for bad in (blob[:30], b"IxHN" + blob[4:], blob[:-1], blob[:37] + (13).to_bytes(8, "little") + blob[45:]):
        with pytest.raises(ValueError):
            faiss_io.parse_index_flat(bad)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 60), st.integers(1, 40), st.integers(1, 70), st.integers(0, 2 ** 31))
def test_topk_properties(n, d, k, seed):
    rng = np.random.default_rng(seed)
    xb = rng.integers(-3, 4, size=(n, d)).astype(np.float32)  # small ints: many exact ties, exact arithmetic
    xq = rng.integers(-3, 4, size=(2, d)).astype(np.float32)
    for D, I in (oracle.canon_search(xq, xb, k), oracle.faiss_seq_search(xq, xb, k)):
        m = min(n, k)
        assert (I[:, m:] == -1).all() and (D[:, m:] == np.float32(NEG)).all()
        assert (np.diff(D[:, :m].astype(np.float64), axis=1) <= 0).all()  # descending
        scores = xq.astype(np.float64) @ xb.astype(np.float64).T
        for i in range(2):
            ids = I[i, :m]
            assert len(set(ids.tolist())) == m
            assert np.array_equal(scores[i, ids].astype(np.float32), D[i, :m])
            rest = np.delete(scores[i], ids)
            if rest.size:
                assert rest.max() <= D[i, m - 1]  # nothing left out beats the k-th
    # a permutation of the rows permutes ids only (canonical scores are per-row)
    perm = rng.permutation(n)
    D1, I1 = oracle.canon_search(xq, xb, k)
    D2, I2 = oracle.canon_search(xq, xb[perm], k)
    assert np.array_equal(D1, D2)


def test_oracle_vs_independent_blas_topk():
    """An implementation the oracle shares no code with: torch's CPU sgemm (what faiss itself calls for nq >= 20,
    `exhaustive_inner_product_blas`) followed by torch.topk.  Same ids except where two scores are closer than the
    fp32 accumulation error, scores within 1e-5 relative -- the tolerance the north_star states for faiss-cpu."""
    import torch
    n, d, k, nq = 30_000, 512, 48, 24
    xb = oracle.synth_fill(n, d, 91)
    xq = oracle.synth_fill(nq, d, 92)
    D, I = oracle.canon_search(xq, xb, k)
    Df, If = oracle.faiss_seq_search(xq, xb, k)
    S = torch.from_numpy(xq) @ torch.from_numpy(xb).T  # fp32 sgemm, blocked summation order
    Dt, It = torch.topk(S, k, dim=1)
    Dt, It = Dt.numpy(), It.numpy()
    eps = d * 2.0 ** -24
    for got_D, got_I in ((D, I), (Df, If)):
        assert np.allclose(got_D, Dt, rtol=1e-5, atol=1e-7)
        for q in range(nq):
            for r in np.nonzero(got_I[q] != It[q])[0]:
                a, b = got_I[q, r], It[q, r]
                assert abs(oracle.dot_canon32(xb[a], xq[q]) - oracle.dot_canon32(xb[b], xq[q])) <= eps
    assert (I == It).mean() > 0.99
