// evs_kernels.cu -- kernel definitions and launch wrappers of libevs (sm_100a).
//
//   scan launchers       score + fused top-k' selection               (evs_scan.cuh)
//   finalize_kernel      merge the per-CTA lists of one query, re-score the k' candidates in the
//                        canonical fp64 order, rank, emit (D,I) or a shard partial
//   merge_partials_kernel  G-way merge of shard partials after the NCCL all-gather
//   l2_normalize_kernel  oldapp.py:35/43/51, 128-bit coalesced, in place
//   f32_to_bf16_kernel   the one-time database layout kernel
//   synth_fill_kernel    counter-based synthetic embeddings (bit-identical to oracle/orc_synth_fill)
#include <float.h>

#include <atomic>

#include "evs_internal.h"
#include "evs_scan.cuh"

namespace evs {

std::atomic<long long> g_kernel_launches{0};

// =============================================================================================
// finalize
// =============================================================================================
struct FinalizeParams {
    const u64* lists;  // [nq][L][kp]
    int L;
    int kp;
    const void* xb;    // rows used for the canonical re-score (fp32 master, or bf16 when there is none)
    int xb_is_bf16;
    const float* xq;   // [nq][d]
    int d;
    int k;
    long long id_base;
    // mode 0: final results
    float* D;          // [nq][k]
    long long* I;      // [nq][k]
    // mode 1: shard partial
    double* P_scores;  // [nq][k]
    long long* P_ids;  // [nq][k]
    float* margins;    // [nq] (may be null)
    // mode 2: shard partial written straight into every rank's gather buffer over NVLink (peer stores)
    Exchange x;
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ bool cand_better(double sa, long long ia, double sb, long long ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}
// The same order on integers (DSETP is quarter-rate on B200, and short-circuit logic branches):
// map the score to an order-preserving u64 (-0.0 folded onto +0.0 so that equal doubles stay equal).
__device__ __forceinline__ u64 score_rank_key(double s) { return f64_to_ordered(s == 0.0 ? 0.0 : s); }
__device__ __forceinline__ int better_i(u64 oa, long long ia, u64 ob, long long ib) {
    return (int)(oa > ob) | ((int)(oa == ob) & (int)(ia < ib));
}
// exact fp32 -> fp64 widening with integer ops for normal numbers (F2F.F64.F32 issues at 1/8 rate)
__device__ __forceinline__ double widen_f32(float f) {
    const uint32_t u = __float_as_uint(f);
    const uint32_t e = (u >> 23) & 0xFFu;
    if (e == 0u || e == 255u) return (double)f;  // zero, subnormal, inf, nan: the slow exact path
    const uint32_t hi = (u & 0x80000000u) | ((e + 896u) << 20) | ((u & 0x007FFFFFu) >> 3);
    const uint32_t lo = u << 29;
    return __hiloint2double((int)hi, (int)lo);
}

__device__ __forceinline__ double canon32_dot_bf16(const __nv_bfloat16* __restrict__ x, const double* __restrict__ q,
                                                   int d, int lane) {
    double acc = 0.0;
    const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
    for (int i = lane; i < d; i += 32) acc = fma((double)__uint_as_float(((uint32_t)xs[i]) << 16), q[i], acc);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, off);
    return acc;
}

// CANON-32 of TWO rows at once with the row loads batched (same accumulation order per row; the two
// rows' load latencies and fp64 chains overlap).  A null row pointer yields -DBL_MAX.
__device__ __forceinline__ void canon32_dot_pair(const float* __restrict__ xa, const float* __restrict__ xb,
                                                 const double* __restrict__ qs, int d, int lane, double& ra, double& rb) {
    double acca = 0.0, accb = 0.0;
    for (int base = 0; base < d; base += 512) {
        float va[16], vb[16];
#pragma unroll
        for (int u = 0; u < 16; u++) {
            int i = base + lane + 32 * u;
            va[u] = (xa && i < d) ? __ldg(xa + i) : 0.f;
            vb[u] = (xb && i < d) ? __ldg(xb + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 16; u++) {
            int i = base + lane + 32 * u;
            if (i < d) {
                const double qv = qs[i];
                acca = fma(widen_f32(va[u]), qv, acca);
                accb = fma(widen_f32(vb[u]), qv, accb);
            }
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        acca = acca + __shfl_xor_sync(0xffffffffu, acca, off);
        accb = accb + __shfl_xor_sync(0xffffffffu, accb, off);
    }
    ra = xa ? acca : -DBL_MAX;
    rb = xb ? accb : -DBL_MAX;
}

// grid = nq, block = 1024.
// The L per-CTA lists are sorted, so the global top-kp is found without merging them all:
//   T0 = kp-th largest list HEAD is a lower bound of the kp-th best key (kp heads are >= it), only the
//   <= kp lists whose head is >= T0 can hold survivors, and only their prefix >= T0 does.  The
//   survivors (about kp + a few for unordered data, kp*kp at most) are sorted in shared memory.
// dynamic smem: surv[finalize_surv_slots(L, kp)] u64 | heads[L] u64 | sc[kp] f64 | id[kp] i64 | ok[kp] u64 | qs[d] f64
// (survivor capacity scap = min(L, kp) * kp: with one list per query -- the tensor-core scans -- the kernel needs
// 7 KB instead of 38 KB of shared memory and eight 256-thread CTAs fit an SM)
__host__ __device__ inline int finalize_surv_cap(int L, int kp) { return (L < kp ? L : kp) * kp; }
__host__ __device__ inline int finalize_surv_slots(int L, int kp) {
    const int scap = finalize_surv_cap(L, kp);
    int pow2 = kp;
    while (pow2 < scap) pow2 <<= 1;  // the sort path pads the survivors to a power of two
    return pow2 > scap + kp ? pow2 : scap + kp;  // the counting path puts kp result slots behind the survivors
}

__global__ void __launch_bounds__(1024) finalize_kernel(FinalizeParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int t = threadIdx.x, nt = blockDim.x;
    const int warp = t >> 5, lane = t & 31, nwarps = nt >> 5;
    const int kp = p.kp, L = p.L;
    const int scap = finalize_surv_cap(L, kp);
    u64* surv = reinterpret_cast<u64*>(smem_raw);           // [finalize_surv_slots(L, kp)]
    u64* heads = surv + finalize_surv_slots(L, kp);         // [L]
    double* sc = reinterpret_cast<double*>(heads + L);      // [kp]
    long long* id = reinterpret_cast<long long*>(sc + kp);  // [kp]
    u64* ok = reinterpret_cast<u64*>(id + kp);              // [kp] integer rank keys of the scores
    double* qs = reinterpret_cast<double*>(ok + kp);        // [d] the query widened once
    __shared__ int s_nsurv, s_nvalid;
    __shared__ unsigned s_maxerr;  // ordered-uint of the largest (canonical - scan) score difference among the candidates
    __shared__ u64 s_T0;
    const int qi = blockIdx.x;
    const u64* lists = p.lists + (size_t)qi * L * kp;
    const float* q = p.xq + (size_t)qi * p.d;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the merge kernel of the exchange path may get resident
    // launched with programmatic stream serialisation: this prologue (the query does not come from the scan) overlaps the
    // tail of the scan kernel; everything below the wait sees the scan's lists
    for (int i = t; i < p.d; i += nt) qs[i] = (double)q[i];
    if (t == 0) {
        s_nsurv = 0;
        s_nvalid = 0;
        s_maxerr = 0u;
        s_T0 = 0ull;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int l = t; l < L; l += nt) heads[l] = lists[(size_t)l * kp];
    __syncthreads();
    // 1. T0 = kp-th largest head (non-empty keys are unique, so exactly one head has rank kp-1).
    //    `nper` adjacent lanes share one head and split the comparison range.
    //    Only every `hs`-th head is ranked (about 2.3*kp of them): still a valid bound (kp keys are >= it),
    //    a quarter of the comparisons, a few more survivors.  If the survivors then overflow their
    //    buffer the bound is recomputed from every head (then at most kp lists qualify: <= kp*kp keys).
    int hs = 1;
    while ((L / (hs * 2)) * 10 >= kp * 23) hs <<= 1;
    for (int attempt = 0; attempt < 2; attempt++) {
    if (attempt == 1) {
        if (s_nsurv <= scap || hs == 1) break;  // uniform: read after the barrier below
        __syncthreads();
        if (t == 0) {
            s_nsurv = 0;
            s_T0 = 0ull;
        }
        hs = 1;
        __syncthreads();
    }
    const int Ls = (L + hs - 1) / hs;  // sampled heads: lists 0, hs, 2hs, ...
    if (Ls >= kp) {
        int nper = 1;
        while (nper < 32 && nper * 2 * Ls <= nt) nper <<= 1;
        const int part = t & (nper - 1);
        for (int l0 = 0; l0 < Ls; l0 += nt / nper) {
            const int l = l0 + t / nper;
            const u64 h = l < Ls ? heads[l * hs] : 0ull;
            int r = 0;
            if (h != 0ull)
                for (int j = part; j < Ls; j += nper) r += heads[j * hs] > h ? 1 : 0;
            for (int off = 1; off < nper; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
            if (h != 0ull && part == 0 && r == kp - 1) s_T0 = h;
        }
    }
    __syncthreads();
    const u64 T0 = s_T0;  // 0: fewer than kp non-empty lists -> every key survives (at most kp*kp)
    // 2. survivors: one warp per qualifying list, prefix >= T0
    for (int l = warp; l < L; l += nwarps) {
        const u64 h = heads[l];
        if (h == 0ull || h < T0) continue;  // warp-uniform
        const u64* src = lists + (size_t)l * kp;
        u64 keys[4];  // kp <= 128: all loads of the list issued before the first use
#pragma unroll
        for (int u = 0; u < 4; u++) keys[u] = (32 * u + lane < kp) ? src[32 * u + lane] : 0ull;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool keep = keys[u] != 0ull && keys[u] >= T0;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (m == 0u) break;
            int pos = 0;
            if (lane == 0) pos = atomicAdd(&s_nsurv, __popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            const int dst = pos + __popc(m & ((1u << lane) - 1u));
            if (keep && dst < scap) surv[dst] = keys[u];
        }
    }
    __syncthreads();
    }  // attempt
    // 3. the kp best survivors in descending order -> A[0..kp)
    const int nsurv = s_nsurv;
    const u64* A;
    if (nsurv <= nt) {
        // usual case (a few more than kp survivors): rank by counting, one barrier instead of a sort
        u64* top = surv + nsurv;  // kp slots behind the survivors
        for (int i = t; i < kp; i += nt) top[i] = 0ull;
        __syncthreads();
        if (t < nsurv) {
            const u64 key = surv[t];
            int r = 0;
            for (int j = 0; j < nsurv; j++) r += surv[j] > key ? 1 : 0;
            if (r < kp) top[r] = key;
        }
        __syncthreads();
        A = top;
    } else {
        int pow2 = kp;
        while (pow2 < nsurv) pow2 <<= 1;
        for (int i = nsurv + t; i < pow2; i += nt) surv[i] = 0ull;
        for (int size = 2; size <= pow2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                for (int e = t; e < (pow2 >> 1); e += nt) {
                    int i = ((e / stride) * (stride << 1)) + (e % stride);
                    cmpx_desc(surv, i, i + stride, (i & size) == 0);
                }
            }
        }
        __syncthreads();
        A = surv;  // A[0..kp) = the kp best scan keys
    }

    // 4. canonical re-score of the kp candidates (a warp takes two candidates at a time)
    for (int c = 2 * warp; c < kp; c += 2 * nwarps) {
        const u64 ka = A[c], kb = (c + 1 < kp) ? A[c + 1] : 0ull;
        const long long rowa = ka ? (long long)key_row(ka) : -1, rowb = kb ? (long long)key_row(kb) : -1;
        double sa = -DBL_MAX, sb = -DBL_MAX;
        if (p.xb_is_bf16) {
            const __nv_bfloat16* xb16 = reinterpret_cast<const __nv_bfloat16*>(p.xb);
            if (ka) sa = canon32_dot_bf16(xb16 + (size_t)rowa * p.d, qs, p.d, lane);
            if (kb) sb = canon32_dot_bf16(xb16 + (size_t)rowb * p.d, qs, p.d, lane);
        } else {
            const float* xb32 = reinterpret_cast<const float*>(p.xb);
            canon32_dot_pair(ka ? xb32 + (size_t)rowa * p.d : nullptr, kb ? xb32 + (size_t)rowb * p.d : nullptr, qs, p.d, lane,
                             sa, sb);
        }
        if (lane == 0) {
            sc[c] = sa;
            id[c] = rowa;
            ok[c] = rowa >= 0 ? score_rank_key(sa) : 0ull;
            if (c + 1 < kp) {
                sc[c + 1] = sb;
                id[c + 1] = rowb;
                ok[c + 1] = rowb >= 0 ? score_rank_key(sb) : 0ull;
            }
        }
    }
    __syncthreads();
    // how far the scan under-estimated its own candidates at most (tf32 truncates towards zero: a bias of ~1e-3 relative;
    // bf16 and fp32 scans scatter around zero): rows that were NOT retained are assumed to be under-estimated no worse
    if (t < kp && id[t] >= 0 && p.margins) atomicMax(&s_maxerr, score_to_ordered((float)(sc[t] - (double)key_score(A[t]))));
    __syncthreads();

    // 5. rank by counting under (score desc, id asc); ids are unique so ranks are a permutation
    if (t < kp && id[t] >= 0) {
        atomicAdd(&s_nvalid, 1);
        const double st = sc[t];
        const long long it = id[t];
        const u64 ot = ok[t];
        int rank = 0;
        // empty slots have id -1 and key 0: a real candidate never loses to them (its key is > 0 or, for
        // -DBL_MAX-like scores, ties are broken by id < -1 being impossible) -> mask by id >= 0 arithmetically
        for (int j = 0; j < kp; j++) rank += better_i(ok[j], id[j], ot, it) & (int)(id[j] >= 0);
        if (rank < p.k) {
            if (p.D) {
                p.D[(size_t)qi * p.k + rank] = (float)st;
                p.I[(size_t)qi * p.k + rank] = it + p.id_base;
            } else if (p.x.world > 0) {
                const size_t e = (size_t)(p.x.q_off + qi) * p.k + rank;
                for (int g = 0; g < p.x.world; g++) {  // the same 16 bytes to every rank's slot for this shard
                    unsigned char* slot = p.x.peer[g] + ((size_t)p.x.parity * p.x.world + p.x.rank) * p.x.slot_bytes;
                    reinterpret_cast<double*>(slot)[e] = st;
                    reinterpret_cast<long long*>(slot + (size_t)p.x.nq_total * p.k * 8)[e] = it + p.id_base;
                }
            } else {
                p.P_scores[(size_t)qi * p.k + rank] = st;
                p.P_ids[(size_t)qi * p.k + rank] = it + p.id_base;
            }
            if (rank == p.k - 1 && p.margins) {
                // all kp slots taken -> rows outside the list scored <= the worst retained scan score, i.e. their
                // canonical score is at most that plus the scan's under-estimate (taken as the largest one observed)
                const float worst = key_score(A[kp - 1]);
                const float under = fmaxf(ordered_to_score(s_maxerr), 0.f);
                p.margins[qi] = (A[kp - 1] != 0ull) ? (float)(st - (double)worst) - under : INFINITY;
            }
        }
    }
    __syncthreads();
    // 6. padding (-FLT_MAX,-1) / (-DBL_MAX,-1) for the slots no candidate ranked into
    const int nvalid = s_nvalid;
    for (int r = nvalid + t; r < p.k; r += nt) {
        if (p.D) {
            p.D[(size_t)qi * p.k + r] = -FLT_MAX;
            p.I[(size_t)qi * p.k + r] = -1;
        } else if (p.x.world > 0) {
            const size_t e = (size_t)(p.x.q_off + qi) * p.k + r;
            for (int g = 0; g < p.x.world; g++) {
                unsigned char* slot = p.x.peer[g] + ((size_t)p.x.parity * p.x.world + p.x.rank) * p.x.slot_bytes;
                reinterpret_cast<double*>(slot)[e] = -DBL_MAX;
                reinterpret_cast<long long*>(slot + (size_t)p.x.nq_total * p.k * 8)[e] = -1;
            }
        } else {
            p.P_scores[(size_t)qi * p.k + r] = -DBL_MAX;
            p.P_ids[(size_t)qi * p.k + r] = -1;
        }
    }
    if (t == 0 && p.margins && nvalid < p.k) p.margins[qi] = INFINITY;
    if (p.x.world > 0) {
        // publish: when the last query's CTA has written its part, raise this shard's flag on every rank
        __syncthreads();
        if (t == 0) {
            __threadfence_system();
            const unsigned prev = atomicAdd(p.x.done, 1u);
            if (prev == (unsigned)p.x.nq_total - 1u) {  // counts across the launches of one search
                *p.x.done = 0u;
                __threadfence_system();
                for (int g = 0; g < p.x.world; g++) {
                    unsigned long long* flags = reinterpret_cast<unsigned long long*>(p.x.peer[g] + 2 * (size_t)p.x.world * p.x.slot_bytes);
                    st_release_sys_u64(flags + (size_t)p.x.parity * p.x.world + p.x.rank, p.x.seq);
                }
            }
        }
    }
}

// =============================================================================================
// publish: copy this shard's partial (scores[nq*k], ids[nq*k]) into its slot on EVERY rank over NVLink
// (peer stores), then raise the arrival flag on every rank.  Used when the partial was produced by a
// path that cannot write the slots itself (tensor-core scan with its host-side overflow repair, empty
// shard).  grid = any, block = 256.
// =============================================================================================
__global__ void __launch_bounds__(256) publish_partials_kernel(Exchange x, long long count, const double* __restrict__ scores,
                                                              const long long* __restrict__ ids) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
        const double sv = scores[e];
        const long long iv = ids[e];
        for (int g = 0; g < x.world; g++) {
            unsigned char* slot = x.peer[g] + ((size_t)x.parity * x.world + x.rank) * x.slot_bytes;
            reinterpret_cast<double*>(slot)[e] = sv;
            reinterpret_cast<long long*>(slot + (size_t)count * 8)[e] = iv;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned prev = atomicAdd(x.done, 1u);
        if (prev == gridDim.x - 1) {
            *x.done = 0u;
            __threadfence_system();
            for (int g = 0; g < x.world; g++) {
                unsigned long long* flags = reinterpret_cast<unsigned long long*>(x.peer[g] + 2 * (size_t)x.world * x.slot_bytes);
                st_release_sys_u64(flags + (size_t)x.parity * x.world + x.rank, x.seq);
            }
        }
    }
}

// =============================================================================================
// merge after the peer-store exchange: wait until every shard's flag for this search has arrived in the
// LOCAL gather buffer, then rank the world*k partials of each query.  grid = nq, block = 256.
// Every rank runs this kernel on its own GPU; the flags are written by the other GPUs' finalize kernels.
// =============================================================================================
__global__ void __launch_bounds__(256) merge_exchange_kernel(Exchange x, long long nq, int k, float* __restrict__ D,
                                                            long long* __restrict__ I, int* __restrict__ timed_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = x.world * k;
    double* sc = reinterpret_cast<double*>(smem_raw);
    long long* id = reinterpret_cast<long long*>(sc + m);
    u64* ok = reinterpret_cast<u64*>(id + m);
    __shared__ int s_nvalid;
    const long long qi = blockIdx.x;
    unsigned char* local = x.peer[x.rank];
    asm volatile("griddepcontrol.wait;" ::: "memory");  // launched programmatically behind this rank's finalise / publish kernel
    if (threadIdx.x < x.world) {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(local + 2 * (size_t)x.world * x.slot_bytes) +
                                         (size_t)x.parity * x.world + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(flag) < x.seq) {
            if (clock64() - t0 > 20000000000ll) {  // ~10 s: a rank never arrived; report instead of hanging
                *timed_out = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    if (threadIdx.x == 0) s_nvalid = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < m; e += blockDim.x) {
        const int part = e / k, r = e % k;
        const unsigned char* slot = local + ((size_t)x.parity * x.world + part) * x.slot_bytes;
        const double sv = __ldcv(reinterpret_cast<const double*>(slot) + (size_t)qi * k + r);
        const long long iv = __ldcv(reinterpret_cast<const long long*>(slot + (size_t)nq * k * 8) + (size_t)qi * k + r);
        sc[e] = sv;
        id[e] = iv;
        ok[e] = iv >= 0 ? score_rank_key(sv) : 0ull;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < m; e += blockDim.x) {
        if (id[e] < 0) continue;
        atomicAdd(&s_nvalid, 1);
        const double st = sc[e];
        const long long it = id[e];
        const u64 ot = ok[e];
        int rank = 0;
        for (int j = 0; j < m; j++) rank += better_i(ok[j], id[j], ot, it) & (int)(id[j] >= 0);
        if (rank < k) {
            D[(size_t)qi * k + rank] = (float)st;
            I[(size_t)qi * k + rank] = it;
        }
    }
    __syncthreads();
    for (int r = s_nvalid + threadIdx.x; r < k; r += blockDim.x) {
        D[(size_t)qi * k + r] = -FLT_MAX;
        I[(size_t)qi * k + r] = -1;
    }
}

// =============================================================================================
// merge of shard partials: scores/ids laid out [part][nq][k]; grid = nq, block = 256
// =============================================================================================
__global__ void __launch_bounds__(256) merge_partials_kernel(int nparts, long long nq, int k,
                                                            const double* __restrict__ scores,
                                                            const long long* __restrict__ ids, long long part_stride,
                                                            float* __restrict__ D, long long* __restrict__ I) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int m = nparts * k;
    double* sc = reinterpret_cast<double*>(smem_raw);
    long long* id = reinterpret_cast<long long*>(sc + m);
    u64* ok = reinterpret_cast<u64*>(id + m);  // integer rank keys (see score_rank_key)
    __shared__ int s_nvalid;
    const long long qi = blockIdx.x;
    if (threadIdx.x == 0) s_nvalid = 0;
    for (int e = threadIdx.x; e < m; e += blockDim.x) {
        int part = e / k, r = e % k;
        size_t src = (size_t)part * part_stride + (size_t)qi * k + r;
        const double sv = scores[src];
        const long long iv = ids[src];
        sc[e] = sv;
        id[e] = iv;
        ok[e] = iv >= 0 ? score_rank_key(sv) : 0ull;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < m; e += blockDim.x) {
        if (id[e] < 0) continue;
        atomicAdd(&s_nvalid, 1);
        const double st = sc[e];
        const long long it = id[e];
        const u64 ot = ok[e];
        int rank = 0;
        for (int j = 0; j < m; j++) rank += better_i(ok[j], id[j], ot, it) & (int)(id[j] >= 0);
        if (rank < k) {
            D[(size_t)qi * k + rank] = (float)st;
            I[(size_t)qi * k + rank] = it;
        }
    }
    __syncthreads();
    for (int r = s_nvalid + threadIdx.x; r < k; r += blockDim.x) {
        D[(size_t)qi * k + r] = -FLT_MAX;
        I[(size_t)qi * k + r] = -1;
    }
}

// =============================================================================================
// L2 normalise, in place.  One warp per row; the row is staged in shared memory as fp32 with
// 128-bit coalesced global loads, the sum of squares is taken in the CANON-32 order in fp64,
// norm = T(float(sqrt(sum))), out = T(float(x) / float(norm)); written back with 128-bit stores.
// =============================================================================================
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T, bool VEC16>
__global__ void __launch_bounds__(256) l2_normalize_kernel(T* __restrict__ x, long long n, int d) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    float* stage = reinterpret_cast<float*>(smem_raw) + (size_t)warp * d;
    constexpr int EPV = 16 / sizeof(T);  // elements per 16-byte vector
    for (long long row = (long long)blockIdx.x * nwarps + warp; row < n; row += (long long)gridDim.x * nwarps) {
        T* xr = x + (size_t)row * d;
        if (VEC16) {
            const uint4* src = reinterpret_cast<const uint4*>(xr);
            for (int v = lane; v < d / EPV; v += 32) {
                uint4 raw = src[v];
                const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
                for (int c = 0; c < EPV; c++) stage[v * EPV + c] = to_f32<T>(e[c]);
            }
        } else {
            for (int i = lane; i < d; i += 32) stage[i] = to_f32<T>(xr[i]);
        }
        __syncwarp();
        double acc = 0.0;
        for (int i = lane; i < d; i += 32) {
            double v = (double)stage[i];
            acc = fma(v, v, acc);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, off);
        const float nrm = to_f32<T>(from_f32<T>((float)sqrt(acc)));
        if (VEC16) {
            uint4* dst = reinterpret_cast<uint4*>(xr);
            for (int v = lane; v < d / EPV; v += 32) {
                uint4 raw;
                T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
                for (int c = 0; c < EPV; c++) e[c] = from_f32<T>(__fdiv_rn(stage[v * EPV + c], nrm));
                dst[v] = raw;
            }
        } else {
            for (int i = lane; i < d; i += 32) xr[i] = from_f32<T>(__fdiv_rn(stage[i], nrm));
        }
        __syncwarp();
    }
}

// =============================================================================================
// fp32 -> bf16 (round to nearest even), 2 x 128-bit loads and 1 x 128-bit store per thread
// =============================================================================================
__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                         long long count) {
    const long long nvec = count / 8;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        float4 a = ldg_stream_f4(reinterpret_cast<const float4*>(src) + 2 * v);
        float4 b = ldg_stream_f4(reinterpret_cast<const float4*>(src) + 2 * v + 1);
        __nv_bfloat162 o0 = __floats2bfloat162_rn(a.x, a.y), o1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 o2 = __floats2bfloat162_rn(b.x, b.y), o3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 out;
        out.x = *reinterpret_cast<uint32_t*>(&o0);
        out.y = *reinterpret_cast<uint32_t*>(&o1);
        out.z = *reinterpret_cast<uint32_t*>(&o2);
        out.w = *reinterpret_cast<uint32_t*>(&o3);
        reinterpret_cast<uint4*>(dst)[v] = out;
    }
    // tail (count not a multiple of 8)
    for (long long i = nvec * 8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        dst[i] = __float2bfloat16_rn(src[i]);
}

// fp16 / bf16 -> fp32 (exact), used by add_dev for half-precision encoder output (oldapp.py:86 astype)
template <typename T>
__global__ void __launch_bounds__(256) to_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, long long count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[i] = to_f32<T>(src[i]);
}

// =============================================================================================
// row gather: dst[i][:] = src[ids[i]][:] (incremental re-index keeps the rows of unchanged files on the device).
// One warp per row, 128-bit copies when d % 4 == 0.
// =============================================================================================
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ ids,
                                                         float* __restrict__ dst, long long n, int d) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        const float* s = src + (size_t)ids[r] * d;
        float* o = dst + (size_t)r * d;
        if ((d & 3) == 0) {
            for (int i = lane; i < (d >> 2); i += 32) reinterpret_cast<float4*>(o)[i] = reinterpret_cast<const float4*>(s)[i];
        } else {
            for (int i = lane; i < d; i += 32) o[i] = s[i];
        }
    }
}

// =============================================================================================
// synthetic rows: value(seed, global row, col) -- same integer function as oracle/orc_synth_fill
// =============================================================================================
__device__ __forceinline__ u64 splitmix64(u64 z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float synth_value(u64 hr, int col) {
    u64 h = splitmix64(hr + (u64)col);
    int s = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) + (int)((h >> 48) & 0xFFFF) - 131070;
    return (float)s;
}
// one thread per 4 consecutive columns (d % 4 == 0) or per element otherwise
__global__ void __launch_bounds__(256) synth_fill_kernel(float* __restrict__ out, long long n, int d, u64 seed,
                                                        long long row_base) {
    const u64 s0 = splitmix64(seed ^ 0xD1B54A32D192ED03ull);
    const long long stride = (long long)gridDim.x * blockDim.x;
    if ((d & 3) == 0) {
        const int vpr = d >> 2;
        const long long total = n * vpr;
        for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
            long long r = v / vpr;
            int c = (int)(v % vpr) * 4;
            u64 hr = splitmix64(s0 + (u64)(r + row_base) * 0x2545F4914F6CDD1Dull);
            float4 o = make_float4(synth_value(hr, c), synth_value(hr, c + 1), synth_value(hr, c + 2),
                                   synth_value(hr, c + 3));
            reinterpret_cast<float4*>(out)[v] = o;
        }
    } else {
        const long long total = n * d;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            long long r = i / d;
            int c = (int)(i % d);
            u64 hr = splitmix64(s0 + (u64)(r + row_base) * 0x2545F4914F6CDD1Dull);
            out[i] = synth_value(hr, c);
        }
    }
}

// =============================================================================================
// launch wrappers
// =============================================================================================
#define EVS_LAUNCH_CHECK()                                  \
    do {                                                    \
        g_kernel_launches.fetch_add(1);                     \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return e__;                 \
    } while (0)

static int clamp_grid(long long want, int cap) {
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

cudaError_t launch_l2_normalize(void* x, long long n, int d, int dtype, int sm_count, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int threads = 256, nwarps = threads / 32;
    size_t smem = (size_t)nwarps * d * sizeof(float);
    int grid = clamp_grid((n + nwarps - 1) / nwarps, sm_count * 8);
    const size_t esz = dtype == EVS_F32 ? 4 : 2;
    const bool vec = ((size_t)d * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
#define EVS_NORM(T, V)                                                                                   \
    do {                                                                                                 \
        if (smem > 48 * 1024)                                                                            \
            cudaFuncSetAttribute(l2_normalize_kernel<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        l2_normalize_kernel<T, V><<<grid, threads, smem, st>>>(reinterpret_cast<T*>(x), n, d);          \
    } while (0)
    if (dtype == EVS_F32) { if (vec) EVS_NORM(float, true); else EVS_NORM(float, false); }
    else if (dtype == EVS_F16) { if (vec) EVS_NORM(__half, true); else EVS_NORM(__half, false); }
    else if (dtype == EVS_BF16) { if (vec) EVS_NORM(__nv_bfloat16, true); else EVS_NORM(__nv_bfloat16, false); }
    else return cudaErrorInvalidValue;
#undef EVS_NORM
    EVS_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_f32_to_bf16(const float* src, void* dst, long long count, int sm_count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    int grid = clamp_grid((count / 8 + 255) / 256, sm_count * 16);
    f32_to_bf16_kernel<<<grid, 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), count);
    EVS_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_to_f32(const void* src, int dtype, float* dst, long long count, int sm_count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    int grid = clamp_grid((count + 255) / 256, sm_count * 16);
    if (dtype == EVS_F16) to_f32_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(src), dst, count);
    else if (dtype == EVS_BF16)
        to_f32_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, count);
    else return cudaErrorInvalidValue;
    EVS_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_gather_rows(const float* src, const long long* ids_dev, float* dst, long long n, int d, int sm_count,
                               cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    int grid = clamp_grid((n * 32 + 255) / 256, sm_count * 16);
    gather_rows_kernel<<<grid, 256, 0, st>>>(src, ids_dev, dst, n, d);
    EVS_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_synth_fill(float* out, long long n, int d, unsigned long long seed, long long row_base, int sm_count,
                              cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    long long work = (d & 3) == 0 ? n * (d >> 2) : n * d;
    int grid = clamp_grid((work + 255) / 256, sm_count * 16);
    synth_fill_kernel<<<grid, 256, 0, st>>>(out, n, d, seed, row_base);
    EVS_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_merge_partials(int nparts, long long nq, int k, const double* scores, const long long* ids,
                                  long long part_stride, float* D, long long* I, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    size_t smem = (size_t)nparts * k * 24;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(merge_partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    merge_partials_kernel<<<(unsigned)nq, 256, smem, st>>>(nparts, nq, k, scores, ids, part_stride, D, I);
    EVS_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_publish_partials(const Exchange& x, long long nq, int k, const double* scores, const long long* ids,
                                    cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    const long long count = nq * k;
    int grid = (int)((count + 255) / 256 < 64 ? (count + 255) / 256 : 64);
    publish_partials_kernel<<<grid, 256, 0, st>>>(x, count, scores, ids);
    EVS_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_merge_exchange(const Exchange& x, long long nq, int k, float* D, long long* I, int* timed_out,
                                  cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    size_t smem = (size_t)x.world * k * 24;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(merge_exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t le = launch_pdl(merge_exchange_kernel, dim3((unsigned)nq), dim3(256), smem, st, x, nq, k, D, I, timed_out);
    g_kernel_launches.fetch_add(1);
    if (le != cudaSuccess) return le;
    return cudaGetLastError();
}

cudaError_t launch_finalize(const FinalizeArgs& a, cudaStream_t st) {
    if (a.nq <= 0) return cudaSuccess;
    FinalizeParams p;
    p.lists = reinterpret_cast<const u64*>(a.lists);
    p.L = a.L;
    p.kp = a.kp;
    p.xb = a.xb;
    p.xb_is_bf16 = a.xb_is_bf16;
    p.xq = a.xq;
    p.d = a.d;
    p.k = a.k;
    p.id_base = a.id_base;
    p.D = a.D;
    p.I = reinterpret_cast<long long*>(a.I);
    p.P_scores = a.P_scores;
    p.P_ids = reinterpret_cast<long long*>(a.P_ids);
    p.margins = a.margins;
    p.x = a.x;
    size_t smem = (size_t)finalize_surv_slots(a.L, a.kp) * 8 + (size_t)a.L * 8 + (size_t)a.kp * 24 + (size_t)a.d * 8 + 16;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // small batches: 1024 threads shorten the single CTA's critical path; large batches: 256 threads
    // so that several queries share an SM
    const int threads = a.nq <= 296 ? 1024 : 256;
    cudaError_t le = launch_pdl(finalize_kernel, dim3((unsigned)a.nq), dim3((unsigned)threads), smem, st, p);
    g_kernel_launches.fetch_add(1);
    if (le != cudaSuccess) return le;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// scan dispatch
// ---------------------------------------------------------------------------------------------
template <typename T, int NQ, int NV>
static cudaError_t launch_scan_t(const ScanArgs& a, ScanPlan* plan, cudaStream_t st) {
    ScanParams p;
    p.xb = a.xb;
    p.n = a.n;
    p.d = a.d;
    p.xq = a.xq;
    p.q0 = a.q0;
    p.lists = reinterpret_cast<u64*>(a.lists);
    p.kp = a.kp;
    p.tile_rows = plan->tile_rows;
    p.stages = plan->stages;
    p.lists_stride_q = plan->grid * a.kp;
    if (plan->variant == 2) {
        auto kern = scan_ring_kernel<T, NQ, NV>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes);
        kern<<<plan->grid, plan->threads, plan->smem_bytes, st>>>(p);
    } else {
        auto kern = scan_direct_kernel<T, NQ, NV>;
        if (plan->smem_bytes > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes);
        kern<<<plan->grid, plan->threads, plan->smem_bytes, st>>>(p);
    }
    EVS_LAUNCH_CHECK();
    return cudaSuccess;
}

template <typename T, int NV>
static cudaError_t launch_scan_nq(const ScanArgs& a, ScanPlan* plan, cudaStream_t st) {
    switch (a.nq_pass) {
        case 1: return launch_scan_t<T, 1, NV>(a, plan, st);
        case 2: return launch_scan_t<T, 2, NV>(a, plan, st);
        case 3: return launch_scan_t<T, 3, NV>(a, plan, st);
        case 4: return launch_scan_t<T, 4, NV>(a, plan, st);
        default: return cudaErrorInvalidValue;
    }
}

template <typename T>
static cudaError_t launch_scan_generic(const ScanArgs& a, ScanPlan* plan, cudaStream_t st) {
    ScanParams p;
    p.xb = a.xb;
    p.n = a.n;
    p.d = a.d;
    p.xq = a.xq;
    p.lists = reinterpret_cast<u64*>(a.lists);
    p.kp = a.kp;
    p.tile_rows = 0;
    p.stages = 0;
    p.lists_stride_q = plan->grid * a.kp;
    for (int qi = 0; qi < a.nq_pass; qi++) {
        p.q0 = a.q0 + qi;
        scan_generic_kernel<T><<<plan->grid, plan->threads, plan->smem_bytes, st>>>(p);
        EVS_LAUNCH_CHECK();
    }
    return cudaSuccess;
}

// number of 16-byte vectors per lane per row for the vectorised kernels, 0 = use the generic kernel
static int vectors_per_lane(int d, int is_bf16) {
    int unit = is_bf16 ? 256 : 128;
    if (d % unit) return 0;
    int nv = d / unit;
    if (is_bf16) return (nv >= 1 && nv <= 4) ? nv : 0;
    return (nv == 1 || nv == 2 || nv == 3 || nv == 4 || nv == 6 || nv == 8) ? nv : 0;
}

// Decide variant, grid, block and shared memory for a scan over `n` rows (does not launch).
cudaError_t plan_scan(long long n, int d, int is_bf16, int kp, int nq_pass, int sm_count, const ScanTuning& tune,
                      ScanPlan* plan) {
    const int nv = vectors_per_lane(d, is_bf16);
    const size_t esz = is_bf16 ? 2 : 4;
    plan->nv = nv;
    if (nv == 0) {
        plan->variant = 0;
        plan->threads = 256;
        plan->smem_bytes = (size_t)(plan->threads / 32) * 2 * kp * 8;
        plan->grid = clamp_grid((n + 7) / 8, sm_count * 4);
        plan->tile_rows = plan->stages = 0;
        return cudaSuccess;
    }
    int variant = tune.scan_variant ? tune.scan_variant : 1;
    if (variant == 2) {
        const int cw = 8;
        size_t row_bytes = (size_t)d * esz;
        int tr = tune.tile_rows > 0 ? tune.tile_rows : (int)(32768 / row_bytes);
        if (tr < 1) tr = 1;
        size_t sel = (size_t)cw * nq_pass * 2 * kp * 8;
        int stages = tune.stages > 0 ? tune.stages : 4;
        const size_t budget = 200 * 1024;
        while (stages > 2 && (size_t)stages * tr * row_bytes + sel + 2 * stages * 8 > budget) stages--;
        while (tr > 1 && (size_t)stages * tr * row_bytes + sel + 2 * stages * 8 > budget) tr--;
        size_t smem = (size_t)stages * tr * row_bytes + 2 * stages * 8 + sel;
        if (smem > budget) variant = 1;
        else {
            plan->variant = 2;
            plan->threads = 32 * (cw + 1);
            plan->tile_rows = tr;
            plan->stages = stages;
            plan->smem_bytes = smem;
            long long ntiles = (n + tr - 1) / tr;
            plan->grid = clamp_grid(ntiles, sm_count);
            return cudaSuccess;
        }
    }
    plan->variant = 1;
    plan->threads = 256;
    plan->tile_rows = plan->stages = 0;
    plan->smem_bytes = (size_t)(plan->threads / 32) * nq_pass * 2 * kp * 8;
    // measured on B200 at 10M x 512 (profiles/r01_tune_scan_*): fp32 rows peak at 2 CTAs/SM, bf16 rows
    // (half the bytes in flight per row) need 4
    int per_sm = tune.ctas_per_sm > 0 ? tune.ctas_per_sm : (is_bf16 ? 4 : 2);
    long long groups = (n + 3) / 4;
    plan->grid = clamp_grid((groups + 7) / 8, sm_count * per_sm);
    return cudaSuccess;
}

cudaError_t launch_scan(const ScanArgs& a, ScanPlan* plan, cudaStream_t st) {
    if (a.n <= 0) return cudaErrorInvalidValue;
    if (plan->nv == 0) return a.is_bf16 ? launch_scan_generic<__nv_bfloat16>(a, plan, st) : launch_scan_generic<float>(a, plan, st);
    if (a.is_bf16) {
        switch (plan->nv) {
            case 1: return launch_scan_nq<__nv_bfloat16, 1>(a, plan, st);
            case 2: return launch_scan_nq<__nv_bfloat16, 2>(a, plan, st);
            case 3: return launch_scan_nq<__nv_bfloat16, 3>(a, plan, st);
            case 4: return launch_scan_nq<__nv_bfloat16, 4>(a, plan, st);
        }
    } else {
        switch (plan->nv) {
            case 1: return launch_scan_nq<float, 1>(a, plan, st);
            case 2: return launch_scan_nq<float, 2>(a, plan, st);
            case 3: return launch_scan_nq<float, 3>(a, plan, st);
            case 4: return launch_scan_nq<float, 4>(a, plan, st);
            case 6: return launch_scan_nq<float, 6>(a, plan, st);
            case 8: return launch_scan_nq<float, 8>(a, plan, st);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace evs
