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
};

__device__ __forceinline__ bool cand_better(double sa, long long ia, double sb, long long ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

__device__ __forceinline__ double canon32_dot_bf16(const __nv_bfloat16* __restrict__ x, const float* __restrict__ q,
                                                   int d, int lane) {
    double acc = 0.0;
    const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
    for (int i = lane; i < d; i += 32) acc = fma((double)__uint_as_float(((uint32_t)xs[i]) << 16), (double)q[i], acc);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, off);
    return acc;
}

// grid = nq, block = 1024.  dynamic smem: 32*kp*8 (warp lists) + kp*8 (scores) + kp*8 (ids)
__global__ void __launch_bounds__(1024) finalize_kernel(FinalizeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kp = p.kp;
    u64* A = reinterpret_cast<u64*>(smem_raw);            // [32][kp]
    double* sc = reinterpret_cast<double*>(A + 32 * kp);  // [kp]
    long long* id = reinterpret_cast<long long*>(sc + kp);  // [kp]
    __shared__ int s_nvalid;
    const int qi = blockIdx.x;
    const u64* lists = p.lists + (size_t)qi * p.L * kp;

    // 1. every warp folds its share of the L sorted lists into its own sorted top-kp
    u64* Aw = A + (size_t)warp * kp;
    for (int i = lane; i < kp; i += 32) Aw[i] = (warp < p.L) ? lists[(size_t)warp * kp + i] : 0ull;
    for (int l = warp + 32; l < p.L; l += 32) warp_merge_top(Aw, lists + (size_t)l * kp, kp, lane);
    // 2. tree over the 32 warps
    for (int step = 1; step < 32; step <<= 1) {
        __syncthreads();
        if ((warp % (2 * step)) == 0) warp_merge_top(Aw, A + (size_t)(warp + step) * kp, kp, lane);
    }
    __syncthreads();

    // 3. canonical re-score of the kp candidates (one warp per candidate)
    const float* q = p.xq + (size_t)qi * p.d;
    for (int c = warp; c < kp; c += 32) {
        u64 key = A[c];
        double s = -DBL_MAX;
        long long row = -1;
        if (key != 0ull) {
            row = (long long)key_row(key);
            if (p.xb_is_bf16)
                s = canon32_dot_bf16(reinterpret_cast<const __nv_bfloat16*>(p.xb) + (size_t)row * p.d, q, p.d, lane);
            else
                s = canon32_dot(reinterpret_cast<const float*>(p.xb) + (size_t)row * p.d, q, p.d, lane);
        }
        if (lane == 0) {
            sc[c] = s;
            id[c] = row;
        }
    }
    if (threadIdx.x == 0) s_nvalid = 0;
    __syncthreads();

    // 4. rank by counting under (score desc, id asc); ids are unique so ranks are a permutation
    const int t = threadIdx.x;
    if (t < kp && id[t] >= 0) {
        atomicAdd(&s_nvalid, 1);
        const double st = sc[t];
        const long long it = id[t];
        int rank = 0;
        for (int j = 0; j < kp; j++) rank += (id[j] >= 0 && cand_better(sc[j], id[j], st, it)) ? 1 : 0;
        if (rank < p.k) {
            if (p.D) {
                p.D[(size_t)qi * p.k + rank] = (float)st;
                p.I[(size_t)qi * p.k + rank] = it + p.id_base;
            } else {
                p.P_scores[(size_t)qi * p.k + rank] = st;
                p.P_ids[(size_t)qi * p.k + rank] = it + p.id_base;
            }
            if (rank == p.k - 1 && p.margins) {
                // all kp slots taken -> rows outside the list scored <= the worst retained scan score
                float worst = key_score(A[kp - 1]);
                p.margins[qi] = (A[kp - 1] != 0ull) ? (float)(st - (double)worst) : INFINITY;
            }
        }
    }
    __syncthreads();
    // 5. padding (-FLT_MAX,-1) / (-DBL_MAX,-1) for the slots no candidate ranked into
    const int nvalid = s_nvalid;
    for (int r = nvalid + t; r < p.k; r += blockDim.x) {
        if (p.D) {
            p.D[(size_t)qi * p.k + r] = -FLT_MAX;
            p.I[(size_t)qi * p.k + r] = -1;
        } else {
            p.P_scores[(size_t)qi * p.k + r] = -DBL_MAX;
            p.P_ids[(size_t)qi * p.k + r] = -1;
        }
    }
    if (t == 0 && p.margins && nvalid < p.k) p.margins[qi] = INFINITY;
}

// =============================================================================================
// merge of shard partials: scores/ids laid out [part][nq][k]; grid = nq, block = 256
// =============================================================================================
__global__ void __launch_bounds__(256) merge_partials_kernel(int nparts, long long nq, int k,
                                                            const double* __restrict__ scores,
                                                            const long long* __restrict__ ids, long long part_stride,
                                                            float* __restrict__ D, long long* __restrict__ I) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = nparts * k;
    double* sc = reinterpret_cast<double*>(smem_raw);
    long long* id = reinterpret_cast<long long*>(sc + m);
    __shared__ int s_nvalid;
    const long long qi = blockIdx.x;
    if (threadIdx.x == 0) s_nvalid = 0;
    for (int e = threadIdx.x; e < m; e += blockDim.x) {
        int part = e / k, r = e % k;
        size_t src = (size_t)part * part_stride + (size_t)qi * k + r;
        sc[e] = scores[src];
        id[e] = ids[src];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < m; e += blockDim.x) {
        if (id[e] < 0) continue;
        atomicAdd(&s_nvalid, 1);
        const double st = sc[e];
        const long long it = id[e];
        int rank = 0;
        for (int j = 0; j < m; j++) rank += (id[j] >= 0 && cand_better(sc[j], id[j], st, it)) ? 1 : 0;
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
    extern __shared__ __align__(16) unsigned char smem_raw[];
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
    size_t smem = (size_t)nparts * k * 16;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(merge_partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    merge_partials_kernel<<<(unsigned)nq, 256, smem, st>>>(nparts, nq, k, scores, ids, part_stride, D, I);
    EVS_LAUNCH_CHECK();
    return cudaSuccess;
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
    size_t smem = (size_t)32 * a.kp * 8 + (size_t)a.kp * 16;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    finalize_kernel<<<(unsigned)a.nq, 1024, smem, st>>>(p);
    EVS_LAUNCH_CHECK();
    return cudaSuccess;
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
