"""GPU: the tcgen05/TMEM tensor-core scan (query batches) -- raw scores against torch, and the full
search against the oracle's canonical ranking (must be bit-identical to every other scan path)."""
import numpy as np
import pytest

import evo_ssearch_b200 as evs
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset_options():
    yield
    evs.set_option("tc_min_nq", 2)
    evs.set_option("tc_pair_min_nq", 129)
    evs.set_option("tc_heap_max_nq", 32)
    evs.set_option("tc_heap_pure_max_nq", 0)
    evs.set_option("tc2_slice_tiles", 0)
    evs.set_option("scan_variant", 0)
    evs.set_option("x3", 0)
    evs.set_option("x3_max_nq", 16)
    evs.set_option("guard", 1)
    evs.set_option("tf32_guard_eps_e6", 0)
    evs.set_option("tc_inline_pre", 1)


@pytest.mark.parametrize("storage,d", [("bf16", 512), ("f32", 512), ("bf16", 768), ("f32", 768), ("bf16", 1024), ("bf16", 64)])
def test_tc_raw_scores_match_torch(storage, d):
    import torch
    n = 70_001  # tail tile, more tiles than SMs
    idx = evs.IndexFlatIP(d, storage=storage)
    idx.add_synthetic(n, seed=3)
    xb = torch.from_numpy(idx.reconstruct_n(0, n)).cuda()
    nmax = idx.tc_max_queries()
    assert nmax >= 16
    evs.set_option("x3", 1)  # fp32 rows: batches up to tc_x3_max_queries() through the 3xTF32 split scan
    for nq in sorted({1, 16, 17, nmax}):
        xq = torch.from_numpy(oracle.synth_fill(nq, d, 4)).cuda()
        got = idx.tc_scores(xq)
        torch.cuda.synchronize()
        if storage == "bf16":
            ref = (xb.bfloat16().double() @ xq.bfloat16().double().T).float()  # exact products of the rounded inputs
            tol = 2e-6  # only fp32 accumulation order differs
        elif nq <= idx.tc_x3_max_queries():
            ref = (xb.double() @ xq.double().T).float()
            tol = 2e-6  # 3xTF32 split scan: fp32-class error (the dropped terms are ~2^-20 relative, plus fp32 accumulation)
        else:
            ref = (xb.double() @ xq.double().T).float()
            tol = 5e-4  # single tf32 truncates each operand to 10 mantissa bits: ~4e-5 rms on unit vectors, ~4 sigma max
        err = (got - ref).abs().max().item()
        assert err <= tol, (storage, d, nq, err)
    if storage == "f32" and idx.tc_x3_max_queries() > 0:
        # the same batch with the split switched off is a plain tf32 scan: three orders of magnitude further from fp64
        evs.set_option("x3", 0)
        xq = torch.from_numpy(oracle.synth_fill(16, d, 4)).cuda()
        err1 = (idx.tc_scores(xq) - (xb.double() @ xq.double().T).float()).abs().max().item()
        evs.set_option("x3", 1)
        err3 = (idx.tc_scores(xq) - (xb.double() @ xq.double().T).float()).abs().max().item()
        assert err3 <= 2e-6 and err1 > 20 * err3, (d, err1, err3)


@pytest.mark.parametrize("pair", [0, 1])
def test_tf32_scan_truncates_its_operands(pair):
    """The certification bound of the single-tf32 scans (evs_api.cu: tf32_trunc_coef) rests on ONE hardware fact: kind::tf32
    ignores the low 13 mantissa bits of both operands (truncation towards zero, not rounding).  Pin it: the raw scores equal
    the fp64 product of the TRUNCATED inputs to fp32-accumulation accuracy -- a rounding tensor core would be ~1e-4 away --
    and every score lies inside the rigorous interval [true - 2^-9 P+, true + 2^-9 P-] the bound is derived from."""
    import torch
    d, n, nq = 512, 40_003, 48
    idx = evs.IndexFlatIP(d)
    idx.add_synthetic(n, seed=13)
    xb = idx.reconstruct_n(0, n)
    xq = oracle.synth_fill(nq, d, 14)
    evs.set_option("x3", 0)
    evs.set_option("tc_pair_min_nq", 1 if pair else 129)
    got = idx.tc_scores(torch.from_numpy(xq).cuda()).cpu().numpy().astype(np.float64)

    def trunc(a):
        return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)
    emu = trunc(xb) @ trunc(xq).T
    assert np.abs(got - emu).max() <= 3e-6, np.abs(got - emu).max()
    rounded = (xb.view(np.uint32) + np.uint32(0x1000) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64) @ \
              (xq.view(np.uint32) + np.uint32(0x1000) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64).T
    assert np.abs(got - rounded).max() > 1e-4  # the test can tell the two apart
    x64, q64 = xb.astype(np.float64), xq.astype(np.float64)
    true = x64 @ q64.T
    pos = np.maximum(x64, 0) @ np.maximum(q64, 0).T + np.maximum(-x64, 0) @ np.maximum(-q64, 0).T  # sum of the positive products
    neg = pos - true                                                                                  # |sum of the negative ones|
    acc = (d / 8 + 16) * 2.0 ** -22
    assert (got >= true - 2.0 ** -9 * pos - acc).all() and (got <= true + 2.0 ** -9 * neg + acc).all()
    # ... and inside the simplified form the finalise kernel uses: under-estimate <= 2^-10 (|q||x| + score)
    B = np.linalg.norm(x64, axis=1)[:, None] * np.linalg.norm(q64, axis=1)[None, :]
    assert (true - got <= 2.0 ** -10 * (1 + 2.0 ** -9) * (B + np.abs(got)) + acc * B).all()


@pytest.mark.parametrize("storage,d", [("bf16", 512), ("f32", 512), ("bf16", 768), ("bf16", 64)])
def test_tc_pair_raw_scores_match_torch(storage, d):
    """The CTA-pair kernel (cta_group::2, M = 256 rows per MMA, query block split over two CTAs): every score of
    every (row, query), including the second CTA's half of the queries, a tail pair-tile whose second CTA is
    empty, several query blocks and several slices per pair."""
    import torch
    n = 70_001 + 128  # last pair-tile: rows only in the first CTA
    idx = evs.IndexFlatIP(d, storage=storage)
    idx.add_synthetic(n, seed=3)
    xb = torch.from_numpy(idx.reconstruct_n(0, n)).cuda()
    evs.set_option("tc_pair_min_nq", 1)
    for nq, slice_tiles in ((1, 0), (33, 0), (256, 3), (600, 0)):
        evs.set_option("tc2_slice_tiles", slice_tiles)
        xq = torch.from_numpy(oracle.synth_fill(nq, d, 4)).cuda()
        got = idx.tc_scores(xq)
        torch.cuda.synchronize()
        if storage == "bf16":
            ref = (xb.bfloat16().double() @ xq.bfloat16().double().T).float()
            tol = 2e-6
        else:
            ref = (xb.double() @ xq.double().T).float()
            tol = 5e-4
        err = (got - ref).abs().max().item()
        assert err <= tol, (storage, d, nq, err)


@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_tc_pair_search_equals_oracle(storage):
    d, n = 512, 200_003
    xb = oracle.synth_fill(n, d, 11)
    xb[n - 1] = xb[5]
    idx = evs.IndexFlatIP(d, storage=storage)
    idx.add(xb)
    xq_all = oracle.synth_fill(700, d, 12)
    evs.set_option("tc_min_nq", 1)
    evs.set_option("tc_pair_min_nq", 1)
    for nq, k in ((1, 48), (40, 12), (257, 48), (700, 100)):
        xq = xq_all[:nq]
        D, I = idx.search(xq, k)
        sample = np.unique(np.linspace(0, nq - 1, 24).astype(int))
        Dr, Ir = oracle.canon_search(xq[sample], xb, k)
        assert np.array_equal(I[sample], Ir), (storage, nq, k)
        assert np.array_equal(D[sample], Dr), (storage, nq, k)
    m = idx.last_margins(700)
    assert (m > (2e-4 if storage == "f32" else 5e-4)).all(), m.min()
    # identical to the one-CTA tensor-core kernel on the whole batch
    evs.set_option("tc_pair_min_nq", 0)
    D1, I1 = idx.search(xq_all, 100)
    assert np.array_equal(I1, I) and np.array_equal(D1, D)
    # more queries than one launch set takes (4096): a second, shorter chunk through its own plan
    evs.set_option("tc_pair_min_nq", 129)
    big = np.concatenate([xq_all] * 6)[:4100]
    Db, Ib = idx.search(big, 12)
    assert np.array_equal(Ib[:700], I[:, :12]) and np.array_equal(Ib[700:1400], I[:, :12]) and np.array_equal(Ib[4096:], I[596:600, :12])
    assert np.array_equal(Db[4096:], D[596:600, :12])


@pytest.mark.parametrize("storage", ["f32", "bf16"])
@pytest.mark.parametrize("d", [512, 768])
def test_tc_search_equals_oracle(storage, d):
    n = 200_003
    xb = oracle.synth_fill(n, d, 11)
    xb[n - 1] = xb[5]
    idx = evs.IndexFlatIP(d, storage=storage)
    idx.add(xb)
    nmax = idx.tc_max_queries()
    xq_all = oracle.synth_fill(2 * nmax + 3, d, 12)
    evs.set_option("tc_min_nq", 1)  # force the tensor-core scan even for one query
    for nq, k in ((1, 48), (5, 12), (16, 48), (17, 1), (nmax, 48), (nmax + 1, 48), (2 * nmax + 3, 100)):
        xq = xq_all[:nq]
        D, I = idx.search(xq, k)
        Dr, Ir = oracle.canon_search(xq, xb, k)
        assert np.array_equal(I, Ir), (storage, d, nq, k, np.argwhere(I != Ir)[:4])
        assert np.array_equal(D, Dr), (storage, d, nq, k)
    m = idx.last_margins(2 * nmax + 3)
    assert (m > (2e-4 if storage == "f32" else 5e-4)).all(), m.min()
    # identical to the GEMV path
    evs.set_option("tc_min_nq", 0)
    D2, I2 = idx.search(xq_all[:20], 48)
    evs.set_option("tc_min_nq", 1)
    D3, I3 = idx.search(xq_all[:20], 48)
    assert np.array_equal(I2, I3) and np.array_equal(D2, D3)


@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_tc_heap_mode_equals_threshold_mode_and_oracle(storage):
    """Small batches keep a running top-k' per (CTA, query) in shared memory (MODE_HEAP: one scan launch, no
    pre-pass / gather / host sync).  Same answer as the threshold scheme and as the oracle, for sorted-ascending
    data (every tile beats the last: the worst case for the on-chip compaction), ties and a ragged tail."""
    d, n = 512, 150_017
    xb = oracle.synth_fill(n, d, 41)
    q = oracle.synth_fill(32, d, 42)
    order = np.argsort(xb @ q[0])  # rows in ascending score order for query 0: admissions never stop
    xb = np.ascontiguousarray(xb[order])
    xb[n - 1] = xb[n - 2]  # exact tie at the very top
    idx = evs.IndexFlatIP(d, storage=storage)
    idx.add(xb)
    evs.set_option("tc_min_nq", 1)
    for nq, k in ((1, 48), (4, 1), (16, 48), (17, 12), (32, 48)):
        Dr, Ir = oracle.canon_search(q[:nq], xb, k)
        for pure in (4, 32):  # with the pre-pass thresholds for nq > 4 (the default), and without any pre-pass
            evs.set_option("tc_heap_max_nq", 32)
            evs.set_option("tc_heap_pure_max_nq", pure)
            fb0 = evs.get_option("tc_fallbacks")
            D, I = idx.search(q[:nq], k)
            assert evs.get_option("tc_fallbacks") == fb0  # the on-chip heaps have no overflow case, whatever the data
            assert np.array_equal(I, Ir) and np.array_equal(D, Dr), (storage, nq, k, pure)
        evs.set_option("tc_heap_max_nq", 0)  # the threshold scheme (here it needs its GEMV re-run for query 0)
        D0, I0 = idx.search(q[:nq], k)
        assert np.array_equal(I0, Ir) and np.array_equal(D0, Dr), (storage, nq, k)
    evs.set_option("tc_heap_max_nq", 32)
    assert (idx.last_margins(32) >= 0).all()


@pytest.mark.parametrize("d", [64, 128, 384, 768, 1024])
def test_tc_paths_across_dims_and_batch_sizes(d):
    """Every tensor-core route (on-chip heaps with and without pre-pass, thresholds + gather, CTA pairs; k' = 64 and
    128) on other embedding widths, ragged sizes and both storages: always the oracle's answer, never a re-run."""
    rng = np.random.default_rng(d)
    n = int(rng.integers(66_000, 90_000))
    xb = oracle.synth_fill(n, d, 50 + d)
    xq = oracle.synth_fill(200, d, 51 + d)
    evs.set_option("tc_min_nq", 1)
    for storage in ("f32", "bf16"):
        idx = evs.IndexFlatIP(d, storage=storage)
        idx.add(xb)
        if idx.tc_max_queries() == 0:
            continue
        fb0 = evs.get_option("tc_fallbacks")
        for nq, k in ((3, 48), (7, 5), (29, 48), (32, 100), (45, 48), (130, 12), (200, 48)):
            D, I = idx.search(xq[:nq], k)
            sample = np.unique(np.linspace(0, nq - 1, 6).astype(int))
            Dr, Ir = oracle.canon_search(xq[sample], xb, k)
            assert np.array_equal(I[sample], Ir) and np.array_equal(D[sample], Dr), (d, storage, nq, k)
        assert evs.get_option("tc_fallbacks") == fb0, (d, storage)


def test_inline_prepass_equals_separate_launches_and_concurrent_handles_do_not_deadlock():
    """Batches of 2..128 queries take their thresholds from a sample tile per CTA INSIDE the scan launch (two grid barriers)
    instead of a pre-pass launch + tau0 launch: same bits either way.  Kernels that synchronise their own grid need every CTA
    resident, so two handles searched concurrently from two threads (each on its own stream) must be ordered by the library:
    the test would hang (and hit the barrier's 2 s watchdog) otherwise."""
    import threading
    d, n, k = 512, 150_011, 48
    xb = oracle.synth_fill(n, d, 91)
    xq = oracle.synth_fill(64, d, 92)
    Dr, Ir = oracle.canon_search(xq, xb, k)
    for storage in ("f32", "bf16"):
        idx = evs.IndexFlatIP(d, storage=storage)
        idx.add(xb)
        for nq in (2, 16, 33, 64):
            l0 = evs.kernel_launches()
            D1, I1 = idx.search(xq[:nq], k)
            n_inline = evs.kernel_launches() - l0
            evs.set_option("tc_inline_pre", 0)
            l0 = evs.kernel_launches()
            D0, I0 = idx.search(xq[:nq], k)
            n_separate = evs.kernel_launches() - l0
            evs.set_option("tc_inline_pre", 1)
            assert n_inline == n_separate - 2, (storage, nq, n_inline, n_separate)  # pre-pass and tau0 launches are gone
            assert np.array_equal(I1, I0) and np.array_equal(D1, D0), (storage, nq)
            assert np.array_equal(I1, Ir[:nq]) and np.array_equal(D1, Dr[:nq]), (storage, nq)
    handles = [evs.IndexFlatIP(d), evs.IndexFlatIP(d, storage="bf16"), evs.IndexFlatIP(d)]
    for h in handles:
        h.add(xb)
    errs = []

    def work(t):
        nq = (16, 40, 64)[t]
        for _ in range(40):
            D, I = handles[t].search(xq[:nq], k)  # host API: each handle on its own stream
            if not (np.array_equal(I, Ir[:nq]) and np.array_equal(D, Dr[:nq])):
                errs.append(t)

    th = [threading.Thread(target=work, args=(t,)) for t in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs


def _planted(spacing, d=512, n=120_000, nq=6):
    """`n` unit rows, 80 of which score 0.9 + spacing * i against query 0: they straddle rank 48."""
    xb = oracle.synth_fill(n, d, 61)
    q = oracle.synth_fill(max(nq, 6), d, 62)[:nq]
    rng = np.random.default_rng(9)
    u = rng.standard_normal((80, d))
    u -= (u @ q[0].astype(np.float64))[:, None] * q[0]
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    a = 0.9 + spacing * np.arange(80)
    xb[40_000:40_080] = (a[:, None] * q[0] + np.sqrt(1 - a[:, None] ** 2) * u).astype(np.float32)
    return xb, q


@pytest.mark.parametrize("x3", [1, 0])
def test_device_guard_reruns_near_tie_queries_exactly_on_every_entry_point(x3):
    """fp32 storage, small batches.  The finalise kernel certifies every result against the error bound of the scan that
    produced the candidates (3xTF32: a few 1e-5 relative to |q| max|x| as a bound, ~1e-6 in fact; single tf32 with
    x3 = 0: the rigorous truncation bound, ~1.2e-3 on unit vectors); a query whose margin is inside the bound -- 80 planted rows closer together than the
    scan can resolve straddle rank 48 -- is queued, re-run ON THE DEVICE with the fp32 GEMV scan (k' = 128) and finalised
    again.  Same answer, bit for bit the oracle's, through the host API, the CUDA-tensor API, the shard-partial API and the
    exchange API; no host synchronisation is involved; ordinary queries of the batch are not re-run."""
    import torch
    k = 48
    xb, q = _planted(2e-7 if x3 else 2e-6)
    evs.set_option("x3", x3)
    idx = evs.IndexFlatIP(512)  # fp32 storage
    idx.add(xb)
    assert abs(idx.max_row_norm - 1.0) < 1e-3
    Dr, Ir = oracle.canon_search(q, xb, k)
    qt = torch.from_numpy(q).cuda()

    def check(D, I, what):
        assert np.array_equal(I, Ir) and np.array_equal(D, Dr), what
    r0, u0 = idx.guard_stats()
    check(*idx.search(q, k), "host")
    r1, u1 = idx.guard_stats()
    assert 1 <= r1 - r0 <= 2 and u1 == u0, (r1 - r0, u1 - u0, idx.last_margins(6))  # query 0, not the ordinary ones
    D, I = idx.search(qt, k)
    check(D.cpu().numpy(), I.cpu().numpy(), "cuda tensor")
    S, I = idx.search_partial(qt, k)
    check(S.cpu().numpy().astype(np.float32), I.cpu().numpy(), "partial")
    px = evs.PeerExchange(0, 0, 1, max_nq=16, max_k=48)
    D, I = idx.search_exchange(px, qt, k)
    check(D.cpu().numpy(), I.cpu().numpy(), "exchange (device)")
    check(*idx.search_exchange_host(px, q, k), "exchange (host)")
    r2, u2 = idx.guard_stats()
    assert 4 <= r2 - r1 <= 8 and u2 == u0
    assert evs.get_option("exact_reruns") == evs.get_option("exact_reruns")  # the host re-ran nothing: all on the device
    # guard off: the margin still says which query cannot be trusted, and nothing is re-run
    evs.set_option("guard", 0)
    D2, I2 = idx.search(q, k)
    m = idx.last_margins(6)
    bound = 1.5e-4 if not x3 else 3e-5
    assert m[0] < bound and (m[1:] > bound).all(), m
    assert np.array_equal(I2[1:], Ir[1:]) and np.array_equal(D2[1:], Dr[1:])
    assert idx.guard_stats() == (r2, u2 + 1)  # nothing re-run; the one uncertified result is counted


def test_guard_for_batches_beyond_the_heap_range():
    """fp32 storage, single-tf32 threshold scans.  40 queries (one-CTA kernel) and 200 queries (CTA-pair kernel) repair
    themselves ON THE DEVICE: the finalise queues the uncertified query, a predicated fp32 GEMV re-run and a second finalise
    overwrite its result -- no host synchronisation, so the CUDA-tensor API stays asynchronous.  300 queries (beyond the
    device queue) read the flags on the host and re-run there.  The oracle's bits every time."""
    import torch
    k = 48
    xb, q = _planted(2e-6, nq=300)
    idx = evs.IndexFlatIP(512)
    idx.add(xb)
    Dr, Ir = oracle.canon_search(q, xb, k)
    for nq in (40, 200):
        e0, (r0, u0) = evs.get_option("exact_reruns"), idx.guard_stats()
        D, I = idx.search(q[:nq], k)
        assert np.array_equal(I, Ir[:nq]) and np.array_equal(D, Dr[:nq]), nq
        Dt, It = idx.search(torch.from_numpy(q[:nq]).cuda(), k)
        assert np.array_equal(It.cpu().numpy(), Ir[:nq]) and np.array_equal(Dt.cpu().numpy(), Dr[:nq]), nq
        r1, u1 = idx.guard_stats()
        assert 2 <= r1 - r0 <= 8 and u1 == u0, (nq, r1 - r0, u1 - u0)  # query 0 in both searches (and hardly anything else)
        assert evs.get_option("exact_reruns") == e0  # the host re-ran nothing
    e0 = evs.get_option("exact_reruns")
    D, I = idx.search(q, k)
    assert np.array_equal(I, Ir) and np.array_equal(D, Dr)
    assert 1 <= evs.get_option("exact_reruns") - e0 <= 6


@pytest.mark.parametrize("x3", [0, 1])
def test_certification_scales_with_the_norms(x3):
    """IndexFlatIP.add accepts any scale: the bound is relative to |q| * max|x|, so un-normalised data neither triggers
    spurious re-runs nor escapes the guard."""
    k = 48
    evs.set_option("x3", x3)
    xb, q = _planted(2e-7)
    idx = evs.IndexFlatIP(512)
    idx.add(xb * np.float32(8.0))
    assert abs(idx.max_row_norm - 8.0) < 1e-2
    q8 = q * np.float32(4.0)
    Dr, Ir = oracle.canon_search(q8, xb * np.float32(8.0), k)
    r0, u0 = idx.guard_stats()
    D, I = idx.search(q8, k)
    assert np.array_equal(I, Ir) and np.array_equal(D, Dr)
    r1, u1 = idx.guard_stats()
    assert 1 <= r1 - r0 <= 2 and u1 == u0  # the planted query only: power-of-two scaling leaves every margin/bound ratio as it was


@pytest.mark.parametrize("nq", [2, 16, 17, 32])
def test_x3_blocks_equal_oracle(nq):
    """fp32 storage, 2..32 queries = one or two 3xTF32 blocks of 16: results are the oracle's, the margins are far above
    the scan's error bound on ordinary data (nothing is re-run), and no buffer can overflow."""
    d, n, k = 512, 150_001, 48
    evs.set_option("x3", 1)
    evs.set_option("x3_max_nq", 32)
    idx = evs.IndexFlatIP(d)
    idx.add_synthetic(n, seed=5)
    xb = idx.reconstruct_n(0, n)
    q = oracle.synth_fill(nq, d, 6)
    r0 = idx.guard_stats()
    fb0 = evs.get_option("tc_fallbacks")
    D, I = idx.search(q, k)
    Dr, Ir = oracle.canon_search(q, xb, k)
    assert np.array_equal(I, Ir) and np.array_equal(D, Dr)
    assert (idx.last_margins(nq) > 1e-4).all()
    assert idx.guard_stats() == r0 and evs.get_option("tc_fallbacks") == fb0


def test_tc_overflow_falls_back_exactly():
    """Adversarial data: thousands of identical rows all beat the pre-pass threshold -> candidate
    buffers overflow -> those queries are re-run through the GEMV scan; the answer stays exact."""
    d, n = 512, 150_000
    xb = oracle.synth_fill(n, d, 21)
    q = oracle.synth_fill(8, d, 22)
    xb[1000:61000] = q[0]  # 60000 duplicates of query 0 itself
    evs.set_option("tc_min_nq", 1)
    evs.set_option("tc_heap_max_nq", 0)  # the threshold scheme (the on-chip heaps of small batches cannot overflow)
    Dr, Ir = oracle.canon_search(q, xb, 48)
    for storage in ("f32", "bf16"):  # bf16 storage: no certification, but the overflow repair is the same
        idx = evs.IndexFlatIP(d, storage=storage)
        idx.add(xb)
        evs.set_option("tc_heap_max_nq", 0)
        fb0, (r0, u0) = evs.get_option("tc_fallbacks"), idx.guard_stats()
        D, I = idx.search(q, 48)
        assert np.array_equal(I, Ir) and np.array_equal(D, Dr), storage
        assert I[0].tolist() == list(range(1000, 1048))
        r1, u1 = idx.guard_stats()
        # the device-side repair fired (8 queries: no host synchronisation).  Query 0 stays formally uncertified: more than
        # k' = 128 rows tie exactly, no margin can separate them (the tie is broken by id, as the oracle does)
        assert r1 > r0 and u1 - u0 <= 1
        assert evs.get_option("tc_fallbacks") == fb0  # ... and the host re-ran nothing
        # a batch beyond the device queue (kRepairCap = 256): flags read on the host, GEMV re-run from there
        big = np.concatenate([q] * 40)[:300]
        Db, Ib = idx.search(big, 48)
        assert np.array_equal(Ib[:8], Ir) and np.array_equal(Db[:8], Dr) and np.array_equal(Ib[296:], Ir[:4]), storage
        assert evs.get_option("tc_fallbacks") > fb0
        evs.set_option("tc_heap_max_nq", 32)
        fb1, g1 = evs.get_option("tc_fallbacks"), idx.guard_stats()
        D, I = idx.search(q, 48)
        assert np.array_equal(I, Ir) and np.array_equal(D, Dr) and evs.get_option("tc_fallbacks") == fb1


def test_tc_clustered_rows_use_the_spill_list_not_the_fallback():
    """Realistic clustering (a burst of near-duplicate images stored in consecutive rows): the few CTAs that own
    those rows overflow their (CTA, query) buffers; the extra keys go to the query's spill list and the search
    stays on the tensor-core path (no GEMV re-run), exact as ever."""
    d, n = 512, 200_000
    xb = oracle.synth_fill(n, d, 31)
    q = oracle.synth_fill(300, d, 32)
    rng = np.random.default_rng(3)
    burst = q[0][None, :] + 0.02 * rng.standard_normal((700, d)).astype(np.float32)
    xb[5000:5700] = burst / np.linalg.norm(burst, axis=1, keepdims=True)
    for storage in ("f32", "bf16"):
        idx = evs.IndexFlatIP(d, storage=storage)
        idx.add(xb)
        evs.set_option("tc_min_nq", 1)
        fb0 = evs.get_option("tc_fallbacks")
        for nq, heap in ((8, 32), (8, 0), (300, 32)):  # on-chip heaps, one-CTA threshold kernel, CTA-pair kernel
            evs.set_option("tc_heap_max_nq", heap)
            D, I = idx.search(q[:nq], 48)
            sample = [0, 1, nq - 1]
            Dr, Ir = oracle.canon_search(q[sample], xb, 48)
            assert np.array_equal(I[sample], Ir) and np.array_equal(D[sample], Dr), (storage, nq)
            assert set(I[0].tolist()) <= set(range(5000, 5700))
        assert evs.get_option("tc_fallbacks") == fb0, storage


def test_config3_1m_x_512_bf16_nq4096_recall():
    """BASELINE config 3: 1M x 512 bf16, 4096 queries, k = 48: recall@k against the fp32 ground truth
    (bar: >= 0.999; the canonical re-rank makes it exact) checked on a sample of the batch."""
    n, d, k, nq = 1_000_000, 512, 48, 4096
    idx = evs.IndexFlatIP(d, storage="bf16")
    idx.add_synthetic(n, seed=0)
    xq = oracle.synth_fill(nq, d, 1)
    fb0 = evs.get_option("tc_fallbacks")
    D, I = idx.search(xq, k)
    assert evs.get_option("tc_fallbacks") == fb0  # no query needed the GEMV re-run
    xb = oracle.synth_fill(n, d, 0)
    sample = np.arange(0, nq, 64)
    Dr, Ir = oracle.canon_search(xq[sample], xb, k)
    hits = sum(len(set(I[s].tolist()) & set(Ir[j].tolist())) for j, s in enumerate(sample))
    recall = hits / (len(sample) * k)
    assert recall >= 0.999, recall
    assert np.array_equal(I[sample], Ir) and np.array_equal(D[sample], Dr)  # in fact exact
    assert (np.diff(D.astype(np.float64), axis=1) <= 0).all()
