"""Every kernel path once, at small sizes, checked against the oracle: the target for compute-sanitizer runs
(`compute-sanitizer --tool memcheck python scripts/sanitize_paths.py`; one tool per GPU job)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import evo_ssearch_b200 as evs
import oracle

d, n = 512, 66_003
xb = oracle.synth_fill(n, d, 7)
xq = oracle.synth_fill(300, d, 8)
ok = True
for storage in ("f32", "bf16"):
    idx = evs.IndexFlatIP(d, storage=storage)
    idx.add(xb)
    for nq, k in ((1, 48), (3, 12), (4, 48), (16, 48), (64, 48), (40, 100), (300, 48)):
        D, I = idx.search(xq[:nq], k)
        sample = sorted({0, nq // 2, nq - 1})
        Dr, Ir = oracle.canon_search(xq[sample], xb, k)
        good = bool(np.array_equal(I[sample], Ir) and np.array_equal(D[sample], Dr))
        ok &= good
        print(storage, nq, k, "OK" if good else "MISMATCH", flush=True)
    px = evs.PeerExchange(0, 0, 1, max_nq=64, max_k=48)
    Dh, Ih = idx.search_exchange_host(px, xq[:16], 48)
    Dr, Ir = idx.search(xq[:16], 48)
    ok &= bool(np.array_equal(Ih, Ir) and np.array_equal(Dh, Dr))
    sub = evs.IndexFlatIP(d, storage=storage)
    sub.add_rows_from(idx, np.arange(0, n, 7))
    ok &= bool(np.array_equal(sub.reconstruct_n(0, 5), xb[0:35:7]))
t = torch.from_numpy(oracle.synth_fill(1000, d, 9, normalize=False)).cuda()
evs.normalize_L2(t)
ok &= bool(np.array_equal(t.cpu().numpy(), oracle.l2_normalize(oracle.synth_fill(1000, d, 9, normalize=False))))
print("fallbacks", evs.get_option("tc_fallbacks"))
print("SANITIZE_PATHS_OK" if ok else "SANITIZE_PATHS_FAILED")
sys.exit(0 if ok else 1)
