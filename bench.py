#!/usr/bin/env python
"""bench.py -- flat inner-product top-k queries/sec at 10M x 512, k = 48 (BASELINE.json `metric`).

    python bench.py [--gpus N --steps K --warmup W]            the CUDA path (one process per GPU)
    python bench.py --impl reference [...]                     the CPU reference arm (oracle port)

A "step" is ONE search call: a batch of `--nq` queries (default 1 -- what the app issues,
oldapp.py:2005, and the case the north_star's roofline target names) scored against all rows, top-48
selected, results delivered.  The database is synthetic (counter-based unit-norm rows, seed 0) and is
row-sharded over the N ranks (total work fixed -> "strong" scaling); each step uses a different query.
The database (20.48 GB fp32) is far larger than the 126 MB L2, so no L2 flush is needed between steps.

One JSON line is printed by rank 0:
  value     whole-job queries/s with queries already resident in HBM, device-timed with CUDA events
            between barrier + synchronize, max over ranks
  e2e       the same through the public host API (numpy in, numpy out): pinned H2D of the query and D2H
            of (D, I) inside the timed region
  roofline  the scan kernel: algorithmic bytes per launch / its mean duration (CUDA events recorded by
            libevs around every scan launch inside the timed region) against the measured HBM peak
  cpu_baseline  the CPU oracle (a labelled port of faiss-cpu's flat-IP scan; faiss itself is not
            installable here) timed on this box's host cores on a bounded row sample
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "flat-IP top-k queries/sec @10M x 512 k=48"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="evs", choices=["evs", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--nq", type=int, default=1)
    ap.add_argument("--k", type=int, default=48)
    ap.add_argument("--storage", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--variant", type=int, default=0, help="scan kernel: 0 auto, 1 direct loads, 2 bulk-async ring")
    ap.add_argument("--cpu-rows", type=int, default=2_000_000, help="row sample for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how shard partials meet -- peer-store exchange kernels over NVLink, or NCCL all-gather + merge")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the short BASELINE config 2 / 3 legs (1M x 512: fp32 nq 1 and 16; bf16 nq 4096) reported under 'configs'")
    ap.add_argument("--tc-min-nq", type=int, default=None, help="override the library's tensor-core threshold (experiments)")
    ap.add_argument("--extra", action="store_true", help="also time query batches 1/4/16/64 (reported under 'extra')")
    return ap.parse_args()


def workload_name(a) -> str:
    return f"{a.rows}x{a.dim} {a.storage} flat-IP, nq={a.nq}, k={a.k}"


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_peak_tflops():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["bf16_tflops"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


def config_legs(evs, torch, dev, k):
    """BASELINE configs 2 and 3 on this GPU, device-timed (CUDA events on torch's stream around back-to-back
    IndexFlatIP.search calls with device-resident queries).  Each leg carries its own roofline: HBM for the
    streaming cases, dense bf16 tensor throughput for the 4096-query batch."""
    hbm_peak, _ = measured_peak_gbs()
    tf_burst, tf_sust, tf_src = measured_peak_tflops()
    out = []

    def queries(d, nq):
        qi = evs.IndexFlatIP(d, device=dev.index)
        qi.add_synthetic(nq, seed=1)
        return torch.from_numpy(qi.reconstruct_n(0, nq)).to(dev)

    def timed(idx, xq, reps):
        for _ in range(3):
            idx.search(xq, k)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            idx.search(xq, k)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    for tag, rows, d, storage, nqs in (("C2", 1_000_000, 512, "f32", (1, 16)), ("C3", 1_000_000, 512, "bf16", (4096,))):
        idx = evs.IndexFlatIP(d, device=dev.index, storage=storage)
        idx.reserve(rows)
        idx.add_synthetic(rows, seed=0)
        esz = 2 if storage == "bf16" else 4
        for nq in nqs:
            xq = queries(d, nq)
            ms = timed(idx, xq, 50 if nq <= 64 else 5)
            scan_ms = idx.time_scan(xq, k, iters=20 if nq <= 64 else 3)
            rec = {"config": f"{tag}: {rows}x{d} {storage}, nq={nq}, k={k}", "ms_per_search": ms, "queries_per_s": nq / ms * 1e3,
                   "scan_ms": scan_ms}
            if nq <= 64:
                alg = rows * d * esz + nq * d * 4 + nq * k * 12
                ach = alg / (scan_ms * 1e-3) / 1e9
                rec["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                                   "frac_whole_search": alg / (ms * 1e-3) / 1e9 / hbm_peak}
            else:
                flops = 2.0 * nq * rows * d
                ach = flops / (scan_ms * 1e-3) / 1e12
                rec["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tf_burst, "unit": "TFLOP/s", "frac": ach / tf_burst,
                                   "frac_of_sustained": ach / tf_sust, "peak_source": tf_src,
                                   "frac_whole_search": flops / (ms * 1e-3) / 1e12 / tf_burst,
                                   "kernel": "evs::tc2_scan_kernel (tcgen05.mma cta_group::2) incl. threshold pre-pass, tau0, gather"}
            out.append(rec)
        del idx
    return out


def ncu_traffic_per_launch(a, world):
    """dram bytes per scan launch from the committed ncu capture of this same workload, else None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        for rec in json.load(open(p)):
            if (rec["rows_per_gpu"], rec["dim"], rec["storage"], rec["nq"]) == (-(-a.rows // world), a.dim, a.storage, a.nq):
                return rec["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        return None
    return None


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle = test infrastructure; used here only as the timed baseline)
# ------------------------------------------------------------------------------------------------
def host_threads() -> int:
    """Threads the CPU legs use: every core this process may run on.  (torchrun exports OMP_NUM_THREADS=1 to its
    workers, which would silently turn the all-cores baseline into a single-thread one.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(a, seconds_budget: float = 12.0) -> dict:
    import oracle
    cores = host_threads()
    rows = min(a.rows, a.cpu_rows)
    xb = oracle.synth_fill(rows, a.dim, seed=0)
    xq = oracle.synth_fill(max(a.nq, 8), a.dim, seed=1)
    scale = rows / a.rows  # a flat scan is linear in rows: q/s at the full size = q/s on the sample * rows/full

    def run(fn, min_reps=3):
        fn(0)
        t0 = time.perf_counter()
        reps = 0
        while reps < min_reps or time.perf_counter() - t0 < seconds_budget / 2:
            fn(reps)
            reps += 1
            if reps >= 2000:
                break
        return reps * a.nq / (time.perf_counter() - t0)

    def q(i):
        return np.roll(xq, i, axis=0)[:a.nq]

    all_qps = run(lambda i: oracle.allcores_search(q(i), xb, a.k, nthreads=cores))
    seq_qps = run(lambda i: oracle.faiss_seq_search(q(i), xb, a.k, simd=True, nthreads=cores))
    return {
        "value": all_qps * scale, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{rows} of {a.rows} rows x {a.dim}, nq={a.nq}, k={a.k}; q/s scaled by {scale:g} (flat scan is linear in rows); "
                  f"all-cores row-split scan of the oracle port (faiss-cpu not installable)",
        "faiss_threading_value": seq_qps * scale,
        "faiss_threading_note": "same port with faiss's own threading (parallel over queries only: one thread scans "
                                "the database for a single query)",
    }


def run_reference(a) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    cores = host_threads()
    rows = min(a.rows, a.cpu_rows)
    xb = oracle.synth_fill(rows, a.dim, seed=0, nthreads=cores)
    xq = oracle.synth_fill(64, a.dim, seed=1)
    scale = rows / a.rows

    def step(i):
        qs = np.roll(xq, -i, axis=0)[:a.nq]
        return oracle.allcores_search(qs, xb, a.k, nthreads=cores)

    for i in range(a.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(a.steps):
        step(a.warmup + i)
    dt = time.perf_counter() - t0
    qps = a.steps * a.nq / dt * scale
    # faiss's own threading for the same call, for the record (short)
    t1 = time.perf_counter()
    reps = 0
    while reps < 3 or time.perf_counter() - t1 < 3.0:
        oracle.faiss_seq_search(np.roll(xq, -reps, axis=0)[:a.nq], xb, a.k, simd=True, nthreads=cores)
        reps += 1
    seq_qps = reps * a.nq / (time.perf_counter() - t1) * scale
    sample = (f"each step scans {rows} of {a.rows} rows x {a.dim} (nq={a.nq}, k={a.k}) with all {cores} host threads "
              f"(rows split across threads); q/s scaled by {scale:g} to the full row count")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3 / scale, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "rows": a.rows, "dim": a.dim, "nq": a.nq, "k": a.k,
                   "engine": "oracle port of faiss-cpu IndexFlatIP (faiss not installable in this image)"},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "faiss_threading_value": seq_qps},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# the CUDA arm
# ------------------------------------------------------------------------------------------------
def run_evs(a) -> int:
    import torch
    import torch.distributed as dist
    import evo_ssearch_b200 as evs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        print(f"warning: --gpus {a.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    if evs.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the evs arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    evs.set_option("scan_variant", a.variant)
    if a.tc_min_nq is not None:
        evs.set_option("tc_min_nq", a.tc_min_nq)
    index = evs.ShardedIndexFlatIP(a.dim, device=local_rank, storage=a.storage, exchange=a.exchange,
                                   exchange_max_nq=max(64, a.nq), exchange_max_k=a.k)
    t_build = time.perf_counter()
    index.add_synthetic(a.rows, seed=0)
    barrier()
    t_build = time.perf_counter() - t_build
    lo, hi = evs.shard_bounds(a.rows, world, rank)

    # query pool: counter-based unit-norm queries (seed 1), identical on every rank
    pool = 64
    qidx = evs.IndexFlatIP(a.dim, device=local_rank)
    qidx.add_synthetic(pool * a.nq, seed=1)
    q_host = qidx.reconstruct_n(0, pool * a.nq).reshape(pool, a.nq, a.dim)
    del qidx
    q_dev = torch.from_numpy(q_host).to(dev)

    def timed_device(nq_pool, steps, warmup, profile=False):
        for i in range(warmup):
            index.search_tensor(nq_pool[i % pool], a.k)
        if profile:
            index.local.scan_profile()
            evs.set_option("profile_scans", 1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = evs.kernel_launches()
        e0.record()
        for i in range(steps):
            D, I = index.search_tensor(nq_pool[(warmup + i) % pool], a.k)
        e1.record()
        barrier()
        launches = evs.kernel_launches() - l0
        evs.set_option("profile_scans", 0)
        ms = max_over_ranks(e0.elapsed_time(e1))
        return ms, launches, (D, I)

    # ---- value: device-resident queries, device-timed ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches, (D_last, I_last) = timed_device(q_dev, a.steps, a.warmup, profile=True)
    n_prof, scan_ms_sum = index.local.scan_profile()
    clocks = sampler.stop() if rank == 0 else None
    value = a.steps * a.nq / (ms_total * 1e-3)

    # ---- e2e: public host API, numpy in / numpy out ----
    def host_step(i):
        return index.search(q_host[i % pool], a.k)

    for i in range(a.warmup):
        host_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        Dh, Ih = host_step(a.warmup + i)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = {"value": a.steps * a.nq / e2e_s, "unit": UNIT, "ms_per_step": e2e_s / a.steps * 1e3,
           "h2d_bytes_per_step": a.nq * a.dim * 4, "d2h_bytes_per_step": a.nq * a.k * 12,
           "api": "IndexFlatIP.search(numpy) -> numpy via evs_index_search" if world == 1
                  else ("ShardedIndexFlatIP.search(numpy) -> numpy (H2D, scan, NCCL all-gather, merge, D2H)" if a.exchange == "nccl"
                        else "ShardedIndexFlatIP.search(numpy) -> numpy via evs_index_search_exchange (H2D, scan, finalise + peer stores, merge, D2H)")}
    # host API and device API must agree on the last step's query
    same = bool(np.array_equal(Ih, I_last.cpu().numpy()) and np.array_equal(Dh, D_last.cpu().numpy()))

    # ---- roofline of the dominant kernel (the scan) ----
    rows_local = hi - lo
    esz = 2 if a.storage == "bf16" else 4
    alg_bytes = rows_local * a.dim * esz + a.nq * a.dim * 4 + a.nq * a.k * 12  # SURVEY.md 8(d), per GPU per launch
    peak, peak_src = measured_peak_gbs()
    scan_ms = scan_ms_sum / max(n_prof, 1)
    tcmin = evs.get_option("tc_min_nq")
    # database passes per search: one with the tensor-core scan, else <= 4 queries per GEMV pass at d = 512
    passes = 1 if (tcmin > 0 and a.nq >= tcmin and rows_local >= 65536) else -(-a.nq // 4)
    achieved = alg_bytes * passes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    achieved_alg = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "evs::scan_*_kernel (score + fused top-k')", "achieved": achieved_alg,
                "peak": peak, "unit": "GB/s", "frac": achieved_alg / peak, "traffic": ncu_traffic_per_launch(a, world),
                "peak_source": peak_src, "frac_of_nominal_8TBs": achieved_alg / 8000.0,
                "algorithmic_bytes_per_search_per_gpu": alg_bytes, "scan_ms_per_search": scan_ms,
                "scan_launches_per_search": passes, "searches_timed": n_prof,
                "hbm_read_rate_incl_repasses": achieved,
                "scan_share_of_step": scan_ms / (ms_total / a.steps) if ms_total > 0 else None}

    extra = None
    if a.extra:
        extra = {}
        for nq in (1, 4, 16, 64):
            qd = torch.from_numpy(np.ascontiguousarray(np.resize(q_host.reshape(-1, a.dim), (pool, nq, a.dim)))).to(dev)
            ms, _, _ = timed_device(qd, max(10, a.steps // 4), 3)
            extra[f"nq{nq}"] = {"queries_per_s": max(10, a.steps // 4) * nq / (ms * 1e-3)}

    configs = None
    if world == 1 and not a.no_configs:
        del index
        torch.cuda.empty_cache()
        configs = config_legs(evs, torch, dev, a.k)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline(a)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": a.storage, "data": "synthetic",
            "config": {"workload": workload_name(a), "rows": a.rows, "dim": a.dim, "nq": a.nq, "k": a.k,
                       "storage": a.storage, "sharding": f"rows/{world}", "rows_per_gpu": rows_local,
                       "exchange": None if world == 1 else a.exchange,
                       "scan_variant": evs.get_option("scan_variant"), "tc_min_nq": evs.get_option("tc_min_nq"),
                       "l2": "inputs larger than L2 (database >> 126 MB); a different query every step",
                       "build_s": round(t_build, 3)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "host_equals_device_result": same,
        }
        if extra:
            line["extra"] = extra
        if configs:
            line["configs"] = configs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main() -> int:
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)
    return run_evs(a)


if __name__ == "__main__":
    sys.exit(main())
