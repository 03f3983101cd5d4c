"""Tuning sweep for the tensor-core scans: pre-pass sample size, ring depth, slice size (device-timed)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evo_ssearch_b200 as evs

def timed(idx, xq, k, reps):
    for _ in range(2):
        idx.search(xq, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        idx.search(xq, k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

rows, d = 1_000_000, 512
qi = evs.IndexFlatIP(d); qi.add_synthetic(4096, seed=1)
q = torch.from_numpy(qi.reconstruct_n(0, 4096)).cuda()
for storage in ("bf16", "f32"):
    idx = evs.IndexFlatIP(d, storage=storage); idx.reserve(rows); idx.add_synthetic(rows, seed=0)
    for nq in ((16, 128, 4096) if storage == "bf16" else (16, 1024)):
        xq = q[:nq].contiguous()
        for sample in (16384, 32768, 65536, 131072):
            evs.set_option("tc_sample_rows", sample)
            ms = timed(idx, xq, 48, 30 if nq <= 128 else 4)
            print(json.dumps(dict(storage=storage, nq=nq, sample=sample, ms=round(ms, 4), qps=round(nq / ms * 1e3))), flush=True)
        evs.set_option("tc_sample_rows", 65536)
        if nq >= 1024:
            for st in (0, 8, 16, 64):
                evs.set_option("tc2_slice_tiles", st)
                ms = timed(idx, xq, 48, 4)
                print(json.dumps(dict(storage=storage, nq=nq, slice_tiles=st, ms=round(ms, 4), qps=round(nq / ms * 1e3))), flush=True)
            evs.set_option("tc2_slice_tiles", 0)
            for stg in (4, 5, 6):
                evs.set_option("tc_stages", stg)
                ms = timed(idx, xq, 48, 4)
                print(json.dumps(dict(storage=storage, nq=nq, stages=stg, ms=round(ms, 4), qps=round(nq / ms * 1e3))), flush=True)
            evs.set_option("tc_stages", 8)
    del idx
print("fallbacks", evs.get_option("tc_fallbacks"))
