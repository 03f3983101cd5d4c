// evs_scan.cuh -- streaming score + fused per-warp top-k' selection for 1..4 queries per pass.
//
// Replaces the inner loop of faiss exhaustive_inner_product_seq + HeapBlockResultHandler that
// index.search() runs at /root/reference/oldapp.py:2005 and :2112.
//
// HBM-bound: every database row is read exactly once per pass; the score matrix is never written.
// A warp owns a row: the 32 lanes split d in 16-byte vectors, multiply against query values held
// in registers, and butterfly-reduce.  Each warp keeps its own candidate buffer in shared memory
// (no atomics, no block barriers in the steady state): a key enters if it beats the warp's running
// threshold tau; when the buffer (2*KP keys) fills, the warp bitonic-sorts it, keeps KP and raises
// tau.  At the end the CTA tree-merges its warps' lists and writes one sorted KP-list per query.
//
// Two variants share all of that and differ only in how rows reach the lanes:
//   scan_direct : ld.global.nc.L1::no_allocate.v4 straight into registers, 4 rows in flight / warp
//   scan_ring   : a producer lane streams row tiles into a shared-memory ring with the bulk async
//                 copy engine (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP); consumer warps
//                 read conflict-free 16-byte vectors from the ring.
#pragma once
#include "evs_common.cuh"
#include "evs_finalize.cuh"

namespace evs {

struct ScanParams {
    const void* xb;    // database rows, row-major [n][d], fp32 or bf16
    long long n;       // rows in this shard
    int d;             // dimension
    const float* xq;   // queries fp32 [*, d] on the device
    int q0;            // first query handled by this launch
    u64* lists;        // out: [query][gridDim.x][kp] keys sorted descending, 0 = empty
    int lists_stride_q;  // = gridDim.x * kp
    int kp;            // candidates kept per list (64 or 128)
    int tile_rows;     // ring variant: rows per stage
    int stages;        // ring variant: ring depth
    // direct variant only:
    unsigned* ticket;  // non-null (single-query launches): the LAST CTA to retire finalises the query in this launch
    unsigned* next_chunk;  // non-null: rows are dealt dynamically, `chunk_groups` row groups per grab (evens out the tail)
    int chunk_groups;
    const int* qmap;   // guard re-run: the launch walks the queries qmap[0 .. *nactive), NQ at a time, and writes the
    const int* nactive;  //   lists of queue slot s at lists[s * lists_stride_q ...]; null = queries q0 .. q0+NQ-1
    int qcap;          // capacity of qmap
    unsigned long long* cta_clock;  // diagnostics (option "scan_clock"): [gridDim.x][2] globaltimer at CTA start / end of its scan loop
};

template <typename T> struct Elem;
template <> struct Elem<float> { static constexpr int VEC = 4; };           // fp32: 4 per 16 B
template <> struct Elem<__nv_bfloat16> { static constexpr int VEC = 8; };  // bf16: 8 per 16 B

// ---------------------------------------------------------------------------------------------
// Query registers: lane l holds, for vector j, the VEC query values that pair with the row
// elements at [VEC*(l + 32 j), VEC*(l + 32 j) + VEC).
// ---------------------------------------------------------------------------------------------
template <typename T, int NQ, int NV>
struct QueryRegs {
    static constexpr int VEC = Elem<T>::VEC;
    float v[NQ][NV][VEC];
    __device__ __forceinline__ void load(const float* __restrict__ xq, int q0, int d, int lane) {
        int qidx[NQ];
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) qidx[qi] = q0 + qi;
        load_idx(xq, qidx, d, lane);
    }
    __device__ __forceinline__ void load_idx(const float* __restrict__ xq, const int (&qidx)[NQ], int d, int lane) {
#pragma unroll
        for (int qi = 0; qi < NQ; qi++)
#pragma unroll
            for (int j = 0; j < NV; j++) {
                const float4* src = reinterpret_cast<const float4*>(xq + (size_t)qidx[qi] * d + VEC * (lane + 32 * j));
#pragma unroll
                for (int h = 0; h < VEC / 4; h++) {
                    float4 t = src[h];
                    v[qi][j][4 * h + 0] = t.x;
                    v[qi][j][4 * h + 1] = t.y;
                    v[qi][j][4 * h + 2] = t.z;
                    v[qi][j][4 * h + 3] = t.w;
                }
            }
    }
};

// unpack one 16-byte vector of row elements to fp32
__device__ __forceinline__ void unpack(const float4& raw, float (&x)[4]) {
    x[0] = raw.x; x[1] = raw.y; x[2] = raw.z; x[3] = raw.w;
}
__device__ __forceinline__ void unpack(const uint4& raw, float (&x)[8]) {
    x[0] = __uint_as_float(raw.x << 16); x[1] = __uint_as_float(raw.x & 0xFFFF0000u);
    x[2] = __uint_as_float(raw.y << 16); x[3] = __uint_as_float(raw.y & 0xFFFF0000u);
    x[4] = __uint_as_float(raw.z << 16); x[5] = __uint_as_float(raw.z & 0xFFFF0000u);
    x[6] = __uint_as_float(raw.w << 16); x[7] = __uint_as_float(raw.w & 0xFFFF0000u);
}
template <typename T> struct RawVec;
template <> struct RawVec<float> { typedef float4 type; };
template <> struct RawVec<__nv_bfloat16> { typedef uint4 type; };

// Scan score of one row against NQ queries.  Fixed order: four partial sums per lane (vector
// component mod 4) over j, combined as (p0+p1)+(p2+p3), then the xor butterfly 16,8,4,2,1.
template <typename T, int NQ, int NV>
__device__ __forceinline__ void row_scores(const typename RawVec<T>::type (&raw)[NV], const QueryRegs<T, NQ, NV>& q,
                                           float (&out)[NQ]) {
    constexpr int VEC = Elem<T>::VEC;
    float p[NQ][4];
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) p[qi][0] = p[qi][1] = p[qi][2] = p[qi][3] = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        float x[VEC];
        unpack(raw[j], x);
#pragma unroll
        for (int qi = 0; qi < NQ; qi++)
#pragma unroll
            for (int e = 0; e < VEC; e++) p[qi][e & 3] = fmaf(x[e], q.v[qi][j][e], p[qi][e & 3]);
    }
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) {
        float s = (p[qi][0] + p[qi][1]) + (p[qi][2] + p[qi][3]);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        out[qi] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Per-warp candidate selection state (registers: count + threshold; shared memory: the buffer).
// ---------------------------------------------------------------------------------------------
template <int NQ>
struct WarpSelect {
    u64* buf;   // shared memory, [NQ][capw] for this warp
    int capw;   // 2 * kp
    int kp;
    int cnt[NQ];
    u64 tau[NQ];

    __device__ __forceinline__ void init(u64* warp_buf, int kp_) {
        buf = warp_buf;
        kp = kp_;
        capw = 2 * kp_;
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) {
            cnt[qi] = 0;
            tau[qi] = 0ull;
        }
    }
    // sort the buffer of query qi, keep the best kp, raise the threshold
    __device__ __forceinline__ void compact(int qi, int lane) {
        u64* b = buf + qi * capw;
        __syncwarp();
        for (int i = cnt[qi] + lane; i < capw; i += 32) b[i] = 0ull;
        warp_bitonic_sort_desc(b, capw, lane);
        u64 t = b[kp - 1];
        tau[qi] = umax64(tau[qi], t);
        cnt[qi] = cnt[qi] < kp ? cnt[qi] : kp;
    }
    // warp-uniform: every lane passes the same score
    __device__ __forceinline__ void offer(int qi, float score, uint32_t row, int lane) {
        u64 key = make_key(score, row);
        if (key > tau[qi]) {
            if (lane == 0) buf[qi * capw + cnt[qi]] = key;
            cnt[qi]++;
            if (cnt[qi] == capw) compact(qi, lane);
        }
    }
    __device__ __forceinline__ void finish(int lane) {
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) compact(qi, lane);
    }
};

// CTA epilogue shared by both variants: tree-merge the consumer warps' sorted lists, write one list
// per query.  `nwarps_sel` warps own buffers sel_base + warp * NQ * capw.  Called by ALL threads.
template <int NQ>
__device__ __forceinline__ void cta_merge_and_store(u64* sel_base, int nwarps_sel, int kp, const ScanParams& p, int slot0 = -1,
                                                    int nvalid = NQ) {
    if (slot0 < 0) slot0 = p.q0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int capw = 2 * kp;
    for (int step = 1; step < nwarps_sel; step <<= 1) {
        __syncthreads();
        if (warp < nwarps_sel && (warp % (2 * step)) == 0 && warp + step < nwarps_sel) {
#pragma unroll
            for (int qi = 0; qi < NQ; qi++)
                warp_merge_top(sel_base + (size_t)(warp * NQ + qi) * capw,
                               sel_base + (size_t)((warp + step) * NQ + qi) * capw, kp, lane);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nvalid * kp; i += blockDim.x) {
        int qi = i / kp, r = i % kp;
        p.lists[(size_t)(slot0 + qi) * p.lists_stride_q + (size_t)blockIdx.x * kp + r] = sel_base[(size_t)qi * capw + r];
    }
}

// ---------------------------------------------------------------------------------------------
// Variant 1: direct loads.  grid = ctas, block = 32 * warps.  Warps walk groups of RPG rows.
// dynamic smem: max(warps * NQ * 2*kp * 8, finalize_smem_bytes when the finalise is fused) bytes.
//
// Launched with programmatic stream serialisation: the CTAs may become resident while the preceding kernel of the
// stream drains and wait (griddepcontrol.wait) until it has completed before they read anything.
//   * rows are dealt statically (group g to warp g mod W) or -- next_chunk != null -- dynamically: a warp's first chunk
//     is static, every further one comes from a global counter, fetched one chunk ahead so that the atomic's latency
//     hides behind the loads.  A static deal leaves the SMs that the memory system serves more slowly still streaming
//     while the others idle; at 1.25M rows per GPU (the metric at N = 8) that tail is several percent of the scan.
//   * ticket != null (single-query searches): the last CTA to retire finalises the query right here (merge of the
//     per-CTA lists, canonical fp64 re-score, ranking, output / peer stores): no second launch on the latency path.
//   * nactive != null: guard re-run of the queries queued by the first finalise (evs_finalize.cuh), NQ at a time.
// ---------------------------------------------------------------------------------------------
template <typename T, int NQ, int NV>
struct ScanOcc {  // CTAs per SM the register allocation is held to (the measured optimum: 2 for fp32 rows, 4 for bf16 rows)
    static constexpr int value = NQ == 1 ? ((sizeof(T) == 2 && NV <= 2) ? 4 : 2) : 1;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <typename T, int NQ, int NV>
__global__ void __launch_bounds__(256, ScanOcc<T, NQ, NV>::value) scan_direct_kernel(ScanParams p, FinalizeParams f) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_last;
    typedef typename RawVec<T>::type raw_t;
    // rows in flight per warp, sized so that query registers + raw vectors stay near 100 registers
    constexpr int QREGS = NQ * NV * Elem<T>::VEC;
    constexpr int RPG_RAW = (100 - QREGS) / (NV * 4);
    constexpr int RPG = RPG_RAW < 1 ? 1 : (RPG_RAW > 4 ? 4 : RPG_RAW);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    u64* sel_base = reinterpret_cast<u64*>(smem_raw);

    // The queries (and, for a guard re-run, the queue) may be the output of the preceding kernel of the stream, which may
    // itself have triggered this launch early: nothing is read before it has completed.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // only now may the next kernel of the stream (finalise / merge / the next search's scan) become resident: its own
    // prologue may read the queries too, and they are complete from here on
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    unsigned long long t_start = 0, t_fin = 0, t_merged = 0;
    if (p.cta_clock && threadIdx.x == 0) t_start = globaltimer_ns();

    int nact = NQ;
    if (p.nactive) {
        nact = *p.nactive;
        if (nact > p.qcap) nact = p.qcap;
        if (nact <= 0) return;  // nothing queued (the usual case): grid-uniform
    }
    const long long total_warps = (long long)gridDim.x * nwarps;
    const long long gw = (long long)blockIdx.x * nwarps + warp;
    const long long ngroups = (p.n + RPG - 1) / RPG;
    const size_t row_vecs = (size_t)p.d / Elem<T>::VEC;  // 16-byte vectors per row
    const raw_t* base = reinterpret_cast<const raw_t*>(p.xb);

    for (int g0 = 0; g0 < nact; g0 += NQ) {
        int qidx[NQ];
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) {
            int sl = g0 + qi < nact ? g0 + qi : nact - 1;  // a short last group repeats its last query (lists not stored)
            qidx[qi] = p.nactive ? p.qmap[sl] : p.q0 + sl;
        }
        QueryRegs<T, NQ, NV> q;
        q.load_idx(p.xq, qidx, p.d, lane);
        WarpSelect<NQ> sel;
        sel.init(sel_base + (size_t)warp * NQ * 2 * p.kp, p.kp);

        auto do_group = [&](long long g) {
            const long long r0 = g * RPG;
            raw_t raw[RPG][NV];
#pragma unroll
            for (int r = 0; r < RPG; r++) {
                long long row = r0 + r < p.n ? r0 + r : p.n - 1;  // clamp: tail rows are re-read, not offered
                const raw_t* src = base + (size_t)row * row_vecs + lane;
#pragma unroll
                for (int j = 0; j < NV; j++) {
                    if constexpr (sizeof(T) == 4) raw[r][j] = ldg_stream_f4(src + 32 * j);
                    else raw[r][j] = ldg_stream_u4(src + 32 * j);
                }
            }
#pragma unroll
            for (int r = 0; r < RPG; r++) {
                float s[NQ];
                row_scores<T, NQ, NV>(raw[r], q, s);
                if (r0 + r < p.n) {
#pragma unroll
                    for (int qi = 0; qi < NQ; qi++) sel.offer(qi, s[qi], (uint32_t)(r0 + r), lane);
                }
            }
        };

        if (p.next_chunk != nullptr && p.nactive == nullptr) {
            const long long C = p.chunk_groups;
            const long long nchunks = (ngroups + C - 1) / C;
            long long chunk = gw;
            while (chunk < nchunks) {
                long long next = 0;
                if (lane == 0) next = total_warps + (long long)atomicAdd(p.next_chunk, 1u);  // consumed after this chunk
                const long long gend = (chunk + 1) * C < ngroups ? (chunk + 1) * C : ngroups;
                for (long long g = chunk * C; g < gend; g++) do_group(g);
                chunk = __shfl_sync(0xffffffffu, next, 0);
            }
        } else {
            for (long long g = gw; g < ngroups; g += total_warps) do_group(g);
        }
        if (p.cta_clock && threadIdx.x == 0 && g0 == 0) {
            p.cta_clock[2 * blockIdx.x] = t_start;
            p.cta_clock[2 * blockIdx.x + 1] = globaltimer_ns();
        }
        sel.finish(lane);
        __syncwarp();
        if (p.cta_clock && threadIdx.x == 0) t_fin = globaltimer_ns();
        const int left = nact - g0;
        cta_merge_and_store<NQ>(sel_base, nwarps, p.kp, p, p.nactive ? g0 : p.q0, left < NQ ? left : NQ);
        __syncthreads();  // the selection buffers are reused by the next group / by the finalise below
        if (p.cta_clock && threadIdx.x == 0) t_merged = globaltimer_ns();
    }

    if constexpr (NQ == 1) {
        if (p.ticket != nullptr) {
            __threadfence();  // this thread's list stores are visible device-wide before the CTA takes its ticket
            __syncthreads();
            if (threadIdx.x == 0) s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1 ? 1 : 0;
            __syncthreads();
            if (s_last) {  // CTA-uniform: every other CTA's list is complete
                if (threadIdx.x == 0) {
                    *p.ticket = 0u;  // ready for the next search (its CTAs take tickets only after this kernel has completed)
                    if (p.next_chunk) *p.next_chunk = 0u;
                }
                __threadfence();
                if (p.cta_clock && threadIdx.x == 0) {
                    unsigned long long* x = p.cta_clock + 2 * gridDim.x;
                    x[0] = globaltimer_ns();  // this (the last) CTA holds its ticket
                    x[2] = t_fin;             // ... had sorted its warps' buffers
                    x[3] = t_merged;          // ... had merged them and stored its list
                    x[4] = p.cta_clock[2 * blockIdx.x + 1];  // ... had left its scan loop (warp 0)
                }
                finalize_prologue(f, p.q0, smem_raw);
                __syncthreads();
                finalize_query(f, p.q0, f.lists + (size_t)p.q0 * f.L * f.kp, smem_raw);
                if (p.cta_clock && threadIdx.x == 0) p.cta_clock[2 * gridDim.x + 1] = globaltimer_ns();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Variant 2: bulk-async ring.  block = 32 * (consumer warps + 1) <= 288; the LAST warp is the producer.
// dynamic smem layout (bytes):
//   [0, stages*stage_bytes)                       ring, stage_bytes = tile_rows * d * sizeof(T) (mult. of 128)
//   [.., + 2*stages*8)                            full[stages], empty[stages] mbarriers
//   [.., + cw*NQ*2*kp*8)                          per-warp candidate buffers
// Tiles are dealt to CTAs round-robin: tile t = blockIdx.x + i * gridDim.x.
// ---------------------------------------------------------------------------------------------
template <typename T, int NQ, int NV>
__global__ void __launch_bounds__(288) scan_ring_kernel(ScanParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef typename RawVec<T>::type raw_t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cw = (blockDim.x >> 5) - 1;  // consumer warps
    const int S = p.stages, TR = p.tile_rows;
    const size_t row_bytes = (size_t)p.d * sizeof(T);
    const size_t stage_bytes = (size_t)TR * row_bytes;
    unsigned char* ring = smem_raw;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)S * stage_bytes);
    uint64_t* empty = full + S;
    u64* sel_base = reinterpret_cast<u64*>(empty + S);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], cw);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const long long ntiles = (p.n + TR - 1) / TR;

    if (warp == cw) {
        // ---------------- producer: one lane feeds the ring ----------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(p.xb);
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                mbar_wait(&empty[stage], phase ^ 1u);
                long long rows = p.n - t * TR;
                if (rows > TR) rows = TR;
                uint32_t bytes = (uint32_t)(rows * row_bytes);
                mbar_arrive_expect_tx(&full[stage], bytes);
                // one bulk copy per <= 32 KiB piece (rows are contiguous in HBM)
                const unsigned char* g = src + (size_t)t * stage_bytes;
                unsigned char* s = ring + (size_t)stage * stage_bytes;
                uint32_t off = 0;
                while (off < bytes) {
                    uint32_t piece = bytes - off < 32768u ? bytes - off : 32768u;
                    bulk_g2s(s + off, g + off, piece, &full[stage]);
                    off += piece;
                }
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------- consumers ----------------
        QueryRegs<T, NQ, NV> q;
        q.load(p.xq, p.q0, p.d, lane);
        WarpSelect<NQ> sel;
        sel.init(sel_base + (size_t)warp * NQ * 2 * p.kp, p.kp);
        int stage = 0;
        uint32_t phase = 0;
        const size_t row_vecs = row_bytes / 16;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            mbar_wait(&full[stage], phase);
            const raw_t* tile = reinterpret_cast<const raw_t*>(ring + (size_t)stage * stage_bytes);
            long long rows = p.n - t * TR;
            if (rows > TR) rows = TR;
            const long long row_base = t * TR;
#pragma unroll 2
            for (int r = warp; r < (int)rows; r += cw) {
                raw_t raw[NV];
                const raw_t* src = tile + (size_t)r * row_vecs + lane;
#pragma unroll
                for (int j = 0; j < NV; j++) raw[j] = src[32 * j];
                float s[NQ];
                row_scores<T, NQ, NV>(raw, q, s);
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) sel.offer(qi, s[qi], (uint32_t)(row_base + r), lane);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == S) {
                stage = 0;
                phase ^= 1u;
            }
        }
        sel.finish(lane);
    }
    cta_merge_and_store<NQ>(sel_base, cw, p.kp, p);
}

// ---------------------------------------------------------------------------------------------
// Generic-d fallback (d not a multiple of the vector width * 32): scalar coalesced loads, queries
// read from global through L1.  Correct for any d >= 1; not a roofline kernel.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float load_elem(const T* p);
template <>
__device__ __forceinline__ float load_elem<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_elem<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __uint_as_float(((uint32_t) * reinterpret_cast<const unsigned short*>(p)) << 16);
}

template <typename T>
__global__ void __launch_bounds__(256) scan_generic_kernel(ScanParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    u64* sel_base = reinterpret_cast<u64*>(smem_raw);
    WarpSelect<1> sel;
    sel.init(sel_base + (size_t)warp * 2 * p.kp, p.kp);
    const float* q = p.xq + (size_t)p.q0 * p.d;
    const T* xb = reinterpret_cast<const T*>(p.xb);
    const long long total_warps = (long long)gridDim.x * nwarps;
    for (long long row = (long long)blockIdx.x * nwarps + warp; row < p.n; row += total_warps) {
        const T* x = xb + (size_t)row * p.d;
        float s = 0.f;
        for (int i = lane; i < p.d; i += 32) s = fmaf(load_elem<T>(x + i), __ldg(q + i), s);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        sel.offer(0, s, (uint32_t)row, lane);
    }
    sel.finish(lane);
    cta_merge_and_store<1>(sel_base, nwarps, p.kp, p);
}

}  // namespace evs
