// evs_tc.cu -- tensor-core scan for query batches: tcgen05.mma + TMEM accumulators + TMA-staged tiles.
//
// Same job as evs_scan.cuh (the inner loop of index.search(), /root/reference/oldapp.py:2005, :2112)
// for batches of more than a few queries, where scoring really is a dense contraction
//     S[row][query] = sum_i xb[row][i] * xq[query][i]
// One database pass serves up to NQ_MAX queries (instead of 4 with the CUDA-core GEMV):
//   A operand = database tile, M = 128 rows x 128-byte K chunks, streamed from HBM by TMA
//               (cp.async.bulk.tensor.2d, 128B swizzle) through a shared-memory ring;
//   B operand = the query block, N = 16..128 queries, loaded once and kept resident in shared memory;
//   D         = 128 rows (TMEM lanes) x N queries (TMEM columns) fp32, double-buffered in TMEM so the
//               MMA of tile i+1 overlaps the epilogue of tile i.
// bf16 rows use kind::f16 (bf16 x bf16 -> fp32); fp32 rows use kind::tf32 directly on the fp32 bits.
// Scan precision only has to find the k' candidates: the final ranking is the canonical fp64
// re-score, so results are identical to the other scan paths.
//
// Warp roles (256 threads, 1 CTA per SM, persistent over row tiles dealt round-robin):
//   warp 0  TMA producer (one lane)      warp 1  MMA issuer (one lane)      warp 2  TMEM alloc/dealloc
//   warps 4-7  epilogue: warp e owns TMEM lanes 32e..32e+31 = rows 32e..32e+31 of the tile.
// Epilogue modes:
//   MODE_MAX     per (32-row group, query) maximum -> gmax      (threshold pre-pass over sampled tiles)
//   MODE_SELECT  every score >= tau0[query] is appended to the (CTA, query) candidate buffer
//   MODE_DUMP    raw scores to global memory (tests)
//   MODE_HEAP    small batches (<= 32 queries, k' = 64): every (CTA, query) keeps a running top-k' in shared memory --
//                a score enters if it beats the CTA's current k'-th best for that query; when more than 128 keys have
//                piled up the owning epilogue warp bitonic-sorts the slots, keeps 64 and raises the threshold -- and
//                the CTA writes one sorted k' list per query at the end, exactly what the GEMV scan writes.  No
//                gather, no overflow case and so no host synchronisation.  For 5..32 queries the thresholds start from
//                the pre-pass bound tau0 (it spares the sorts of the first tiles); for <= 4 queries there is no
//                pre-pass either and the whole search is this launch plus finalize_kernel.
// tau0[query] = k'-th largest group maximum of the pre-pass: at least k' distinct rows score >= tau0,
// so it is a valid lower bound of the k'-th best score and the SELECT pass keeps a superset of the
// top k'.  A (CTA, query) buffer that is full sends its extra keys to the query's spill list (clustered rows); if that
// overflows too, overflow[query] is raised and the host re-runs the query through the GEMV path (exactness never
// depends on the data).
#include <cuda.h>
#include <float.h>

#include "evs_internal.h"
#include "evs_common.cuh"
#include "evs_tc_common.cuh"
#include "evs_finalize.cuh"

namespace evs {

struct TcParams {
    long long n;          // rows in the shard
    int d;                // dimension
    int nq;               // valid queries of this launch (all blocks)
    int npad;             // N of the MMA = queries per block (multiple of 16)
    int nblocks;          // query blocks walked inside the launch: block b = queries [b*npad, (b+1)*npad)
    int nqp;              // nblocks * npad (padded query count; row pitch of the per-query arrays)
    int nk;               // 128-byte K chunks per row
    int stages;           // ring depth
    long long ntiles;     // tiles this launch walks: tile = (blockIdx.x + i*gridDim.x) * tile_stride
    long long tile_stride;
    // MODE_MAX
    uint32_t* gmax;       // [ntiles*4][nqp] ordered-uint maxima per 32-row group
    // MODE_SELECT
    const float* tau0;    // [nqp]
    u64* cand;            // [nqp][gridDim.x][cap]: the buffers of one query are contiguous for the gather
    int cap;
    int* counts;          // [nqp][gridDim.x]
    int* overflow;        // [nqp] set to 1 when a buffer AND the query's spill list overflowed
    int* spill_cnt;       // [nqp] keys offered to the spill list of the query
    u64* spill;           // [nqp][TC_SPILL_CAP]: where a full (CTA, query) buffer sends its extra keys (clustered rows)
    int kp_sel;           // k' the thresholds are taken for (in-launch pre-pass)
    // MODE_HEAP
    u64* lists;           // [nq][gridDim.x][64] sorted descending, 0 = empty
    // MODE_DUMP
    float* dump;          // [n][nqp]
    // MODE_HEAP / MODE_SELECT with the threshold pre-pass INSIDE the launch (inline_pre = 1): every CTA first scores one sampled
    // tile (tile blockIdx.x * pre_stride) per query block and writes its group maxima, a grid barrier makes them visible,
    // the CTAs that own a query compute its tau0, a second barrier publishes the thresholds -- two launches and ~15 us less
    // per search than the MODE_MAX launch + tc_tau0_kernel
    int inline_pre;
    long long pre_stride;
    float* tau0_w;        // the tau0 array, writable
    unsigned* bar;        // [0] arrivals, [1] generation, [2] sticky failure flag (a barrier timed out once: never used again)
};

// ---------------------------------------------------------------------------------------------
// the kernel.  dynamic shared memory (1024-byte aligned base):
//   [0, nk*npad*128)                  resident query block, chunk-major: chunk c at c*npad*128
//   [.., + stages*16384)              ring of database tiles (128 rows x 128 B)
//   then barriers, TMEM base address, tau0[npad], cnt[npad]
// ---------------------------------------------------------------------------------------------

// Grid barrier for the persistent scan (one CTA per SM, all resident): sense-reversing on two global words, called by ONE
// thread per CTA.  A watchdog (~2 s) sets the sticky failure flag instead of hanging the GPU; the caller then falls back to
// thresholds that are valid without the pre-pass (and the host never uses the in-launch pre-pass on this handle again).
__device__ __forceinline__ void tc_grid_barrier(unsigned* bar, unsigned nctas) {
    volatile unsigned* gen = bar + 1;
    __threadfence();
    const unsigned g = *gen;
    if (atomicAdd(bar, 1u) == nctas - 1u) {
        bar[0] = 0u;
        __threadfence();
        atomicAdd(bar + 1, 1u);
    } else {
        const long long t0 = clock64();
        while (*gen == g) {
            if (clock64() - t0 > 4000000000ll) {
                *reinterpret_cast<volatile unsigned*>(bar + 2) = 1u;
                break;
            }
            __nanosleep(100);
        }
    }
    __threadfence();
}

// X3 (fp32 rows only): 3xTF32 split scan.  tf32 keeps 11 significant bits of each operand, a ~1e-3 relative error that
// is far above fp32 noise; here every operand is split  v = hi + lo  (hi = the tf32 part, lo = v - hi exactly) and the
// products A_hi*Q_hi + A_hi*Q_lo + A_lo*Q_hi are accumulated: the dropped terms are ~2^-20 relative, i.e. fp32-class
// scores from the tensor cores.  The queries are split once (tc_split_queries); the resident query block holds, per K
// chunk, the NP hi rows followed by the NP lo rows, so that ONE MMA with N = 2 NP multiplies the tile by both halves
// (accumulator columns [0, NP) = A_hi*Q_hi, [NP, 2 NP) = A_hi*Q_lo; the tensor core narrows the raw fp32 values to their
// tf32 part by itself: kind::tf32 ignores the low 13 mantissa bits) and a second MMA with N = NP adds A_lo*Q_hi into
// columns [0, NP); the epilogue adds the two column halves.
// The A operands of both MMAs come from TENSOR MEMORY (tcgen05.mma with [a_tmem]): eight converter warps (8-15, two per
// TMEM lane quarter, alternating stages) read each staged tile once from shared memory -- a thread owns a row, its eight
// 16-byte reads undo the 128-byte swizzle -- and write the raw values and their lo parts into a six-deep ring of TMEM
// columns with tcgen05.st.  Shared memory then carries, per 16 KB stage, the TMA write (16 KB), the converters' read
// (16 KB) and the MMAs' reads of the query block (12 KB): 44 KB.  The first two versions kept A_lo in a shared-memory
// side ring (TMA write 16 + converter read 16 + write 16 + two A reads of 16 each + 12 = 92 KB per stage against the
// ~82 KB the SM can move while HBM delivers the stage): they ran at 0.53x and 0.64x of the HBM roofline.
template <typename T, int MODE, bool X3>
__global__ void __launch_bounds__(X3 ? 512 : 256, 1)
tc_scan_kernel(const __grid_constant__ CUtensorMap tm_db, const __grid_constant__ CUtensorMap tm_q, TcParams p) {
    static_assert(!X3 || sizeof(T) == 4, "the 3xTF32 split applies to fp32 rows");
    constexpr bool TF32 = sizeof(T) == 4;
    constexpr int QH = X3 ? 2 : 1;             // resident query copies (hi, lo): chunk c holds [hi: NP rows][lo: NP rows]
    constexpr int AS = TC_X3_TMEM_STAGES;      // X3: TMEM ring of A operands, 64 columns per stage (raw | lo)
    constexpr uint32_t A_COL0 = 128;           // X3: the ring starts behind the two accumulator buffers (2 x 3 NP <= 96)
    constexpr int EC = 128 / sizeof(T);        // elements per 128-byte chunk
    constexpr int KSTEP_BYTES = 32;            // one MMA consumes 32 bytes of K per row (16 bf16 / 8 tf32)
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages, NK = p.nk, NP = p.npad;

    // 128B-swizzled operands need 1024-byte aligned bases: align by hand (1 KiB of slack is allocated)
    unsigned char* q_smem = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);  // X3: [hi: NK chunks][lo: NK chunks]
    unsigned char* ring = q_smem + (size_t)QH * NK * NP * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)S * TC_STAGE_BYTES);
    uint64_t* q_full = bars;             // 1: the query block has landed
    uint64_t* q_empty = bars + 1;        // 1: every MMA that reads the query block has completed
    uint64_t* full = bars + 2;           // S
    uint64_t* empty = full + S;          // S
    uint64_t* acc_full = empty + S;      // 2
    uint64_t* acc_empty = acc_full + 2;  // 2
    uint64_t* a_full = acc_empty + 2;    // AS (X3): the converter warps have filled the TMEM stage
    uint64_t* a_empty = a_full + AS;     // AS (X3): the MMAs that read the TMEM stage have completed
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(a_empty + AS);
    float* tau_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_base_smem + 4) + 15) & ~(uintptr_t)15);  // float4 reads
    int* cnt_s = reinterpret_cast<int*>(tau_s + NP);
    u64* heap_s = reinterpret_cast<u64*>(cnt_s + NP);  // MODE_HEAP: [NP][TC_HEAP_SLOTS]; 16-byte aligned (NP % 16 == 0)

    const int ACC = (X3 ? 3 : 1) * NP;  // accumulator columns per buffer (X3: A_hi*Q_hi | A_hi*Q_lo | A_lo*Q_hi)
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(2 * ACC)) tmem_cols <<= 1;
    if (X3) tmem_cols = 512;  // accumulators [0, 128) + AS x 64 columns of A operands

    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_db);
        prefetch_tmap(&tm_q);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], X3 ? 4 : 1);  // X3: released by the four converter warps that read the stage
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 4);
        }
        for (int a = 0; a < AS; a++) {
            mbar_init(&a_full[a], 4);  // one arrival per converter warp of the stage's group
            mbar_init(&a_empty[a], 1);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_base_smem, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    // everything above (barriers, TMEM, descriptor prefetch) may have overlapped the preceding kernel of the stream
    // (programmatic launch); the queries (bf16 copy), thresholds and counters below are that kernel's output
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // dependents (tau0 / finalise, whose prologues read the queries) may get resident from here on: the kernel that
    // produced the queries -- possibly one that triggered THIS launch early -- has completed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // tiles of this CTA (the same list for every query block)
    const long long my_tiles = p.ntiles > blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // in-launch pre-pass: one leading sample tile per query block (decided once, grid-uniformly: the sticky failure flag is
    // only ever set by a launch that is already past this point)
    const int pre = ((MODE == MODE_HEAP || MODE == MODE_SELECT) && p.inline_pre && *reinterpret_cast<volatile unsigned*>(p.bar + 2) == 0u) ? 1 : 0;
    const long long n_iter = my_tiles + pre;
    auto tile_of = [&](long long i) -> long long {
        return i < pre ? (long long)blockIdx.x * p.pre_stride : (blockIdx.x + (i - pre) * gridDim.x) * p.tile_stride;
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0 && n_iter > 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int b = 0; b < p.nblocks; b++) {
                if (b > 0) mbar_wait(q_empty, (uint32_t)((b - 1) & 1));  // previous block's MMAs are done with it
                mbar_arrive_expect_tx(q_full, (uint32_t)(QH * NK * NP * 128));
                for (int c = 0; c < NK; c++) tma_load_2d(q_smem + (size_t)c * QH * NP * 128, &tm_q, c * EC, b * NP, q_full);
                if (X3)  // the lo halves of the queries are rows [nqp, 2 nqp) of the split query matrix
                    for (int c = 0; c < NK; c++)
                        tma_load_2d(q_smem + ((size_t)c * QH + 1) * NP * 128, &tm_q, c * EC, p.nqp + b * NP, q_full);
                for (long long i = 0; i < n_iter; i++) {
                    const long long tile = tile_of(i);
                    const int row0 = (int)(tile * TC_BM);
                    for (int c = 0; c < NK; c++) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        mbar_arrive_expect_tx(&full[stage], TC_STAGE_BYTES);
                        tma_load_2d(ring + (size_t)stage * TC_STAGE_BYTES, &tm_db, c * EC, row0, &full[stage]);
                        if (++stage == S) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0 && n_iter > 0) {
            const uint32_t idesc = make_idesc(TF32, TC_BM, NP);
            const uint32_t idesc2 = make_idesc(TF32, TC_BM, 2 * NP);  // X3: both query halves in one MMA
            int stage = 0, ts = 0;
            uint32_t phase = 0, tphase = 0;
            long long it = 0;  // tile counter across blocks: accumulator buffer and its phase
            for (int b = 0; b < p.nblocks; b++) {
                mbar_wait(q_full, (uint32_t)(b & 1));
                tc_fence_after();
                for (long long i = 0; i < n_iter; i++, it++) {
                    const int a = (int)(it & 1);
                    const uint32_t aphase = (uint32_t)((it >> 1) & 1);
                    mbar_wait(&acc_empty[a], aphase ^ 1u);  // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(a * ACC);
                    for (int c = 0; c < NK; c++) {
                        const uint32_t b_addr = smem_u32(q_smem + (size_t)c * QH * NP * 128);
                        if (X3) {
                            mbar_wait(&a_full[ts], tphase);  // the converters have put this K chunk (raw | lo) into TMEM
                            tc_fence_after();
                            const uint32_t a_raw = tmem_base + A_COL0 + (uint32_t)(ts * 64);
                            // raw tile x [Q_hi; Q_lo] (N = 2 NP) into columns [0, 2 NP), lo tile x Q_hi (N = NP) into columns
                            // [2 NP, 3 NP): two INDEPENDENT accumulation chains issued alternately.  MMAs this small are bound
                            // by their latency (~100 cycles each when every one waits for the previous one's accumulator:
                            // ncu had the converters waiting for the TMEM ring half of the time with both products chained
                            // through the same columns)
#pragma unroll
                            for (int k = 0; k < 128 / KSTEP_BYTES; k++) {
                                const uint64_t bd = smem_desc_sw128(b_addr + k * KSTEP_BYTES);
                                umma_ts_tf32(d_tmem, a_raw + 8 * k, bd, idesc2, (uint32_t)((c | k) != 0));
                                umma_ts_tf32(d_tmem + 2 * NP, a_raw + 32 + 8 * k, bd, idesc, (uint32_t)((c | k) != 0));
                            }
                            umma_commit(&a_empty[ts]);  // frees the TMEM stage when these MMAs have read it
                            if (++ts == AS) {
                                ts = 0;
                                tphase ^= 1u;
                            }
                        } else {
                            mbar_wait(&full[stage], phase);
                            tc_fence_after();
                            const uint32_t a_addr = smem_u32(ring + (size_t)stage * TC_STAGE_BYTES);
#pragma unroll
                            for (int k = 0; k < 128 / KSTEP_BYTES; k++) {
                                umma<TF32>(d_tmem, smem_desc_sw128(a_addr + k * KSTEP_BYTES),
                                           smem_desc_sw128(b_addr + k * KSTEP_BYTES), idesc, (uint32_t)((c | k) != 0));
                            }
                            umma_commit(&empty[stage]);  // frees the ring slot when these MMAs have read it
                            if (++stage == S) {
                                stage = 0;
                                phase ^= 1u;
                            }
                        }
                    }
                    umma_commit(&acc_full[a]);  // accumulator complete -> epilogue
                }
                umma_commit(q_empty);  // the query block may be overwritten once all of the above completed
            }
        }
        __syncwarp();
    } else if (X3 && warp >= 8) {
        // ================= converter warps (3xTF32): staged tile -> TMEM as the two A operands (raw | lo) =================
        const int cw = warp - 8;        // 0..7
        const int quarter = cw & 3;     // = warp % 4: the TMEM lanes this warp may access
        const int grp = cw >> 2;        // the two groups of four warps take alternate stages
        const int r = quarter * 32 + lane;  // this thread's row of the tile = its TMEM lane
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + A_COL0;
        int stage = 0, ts = 0;
        uint32_t phase = 0, tphase = 0;
        long long itc = 0;
        for (int b = 0; b < p.nblocks; b++) {
            for (long long i = 0; i < n_iter; i++) {
                for (int c = 0; c < NK; c++, itc++) {
                    if ((int)(itc & 1) == grp) {
                        mbar_wait(&full[stage], phase);           // the raw fp32 tile has landed (TMA)
                        mbar_wait(&a_empty[ts], tphase ^ 1u);     // the MMAs that read this TMEM stage are done
                        tc_fence_after();
                        // row r of the 128-byte-swizzled tile: logical 16-byte chunk u sits at chunk u ^ (r & 7); a quarter warp
                        // (8 consecutive rows) covers all 32 banks once per load: conflict-free
                        const float4* row = reinterpret_cast<const float4*>(ring + (size_t)stage * TC_STAGE_BYTES + (size_t)r * 128);
                        uint32_t v[32];
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const float4 t = row[u ^ (r & 7)];
                            v[4 * u + 0] = __float_as_uint(t.x);
                            v[4 * u + 1] = __float_as_uint(t.y);
                            v[4 * u + 2] = __float_as_uint(t.z);
                            v[4 * u + 3] = __float_as_uint(t.w);
                        }
                        // raw values as they are: kind::tf32 reads exactly their upper 19 bits (tests/test_gpu_tensorcore.py
                        // checks the 3xTF32 scores to 2e-6, which a rounding MMA would miss)
                        tmem_st32(t_lane + (uint32_t)(ts * 64), v);
                        // lo = v - (v with the low 13 mantissa bits cleared): exact.  (inf - inf would turn an infinite element
                        // into NaN: keep hi only.)
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const float f = __uint_as_float(v[j]);
                            v[j] = fabsf(f) <= FLT_MAX ? __float_as_uint(f - __uint_as_float(v[j] & 0xFFFFE000u)) : 0u;
                        }
                        tmem_st32(t_lane + (uint32_t)(ts * 64 + 32), v);
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(&empty[stage]);  // every lane has its row in registers: the ring slot may be refilled
                            mbar_arrive(&a_full[ts]);
                        }
                    }
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1u;
                    }
                    if (++ts == AS) {
                        ts = 0;
                        tphase ^= 1u;
                    }
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= epilogue =================
        const int e = warp & 3;                 // TMEM lane quarter
        const int row_in_tile = e * 32 + lane;  // this thread's row
        const int et = threadIdx.x - 128;       // 0..127 among the epilogue threads
        long long it = 0;
        for (int b = 0; b < p.nblocks; b++) {
            const int qb = b * NP;  // first query of the block
            if (MODE == MODE_SELECT) {
                asm volatile("bar.sync 1, 128;" ::: "memory");  // all epilogue warps have left the previous block
                for (int c = et; c < NP; c += 128) {
                    tau_s[c] = (!pre && qb + c < p.nq) ? p.tau0[qb + c] : INFINITY;  // pre: set behind the sample tile below
                    cnt_s[c] = 0;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            if (MODE == MODE_HEAP) {  // one block per launch in this mode
                for (int c = et; c < NP; c += 128) {
                    // start from the pre-pass bound when there is one (it spares the early sorts: with it a CTA sees
                    // ~10 admissions per query in all); padded queries never admit anything
                    tau_s[c] = (!pre && qb + c < p.nq) ? (p.tau0 ? p.tau0[qb + c] : -INFINITY) : INFINITY;
                    cnt_s[c] = 0;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            for (long long i = 0; i < n_iter; i++, it++) {
                const int a = (int)(it & 1);
                const uint32_t aphase = (uint32_t)((it >> 1) & 1);
                const bool sample = i < pre;  // the leading sample tile: group maxima only, then the thresholds
                const long long tile = tile_of(i);
                const long long row = tile * TC_BM + row_in_tile;
                const bool row_ok = row < p.n;
                mbar_wait(&acc_full[a], aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(a * ACC);
                for (int c0 = 0; c0 < NP; c0 += 16) {
                    uint32_t v[16];
                    if (X3) {  // score = A_hi Q_hi + (A_lo Q_hi + A_hi Q_lo)
                        uint32_t w[16], z[16];
                        tmem_ld16_nowait(taddr + c0, v);
                        tmem_ld16_nowait(taddr + NP + c0, w);
                        tmem_ld16_nowait(taddr + 2 * NP + c0, z);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j++)
                            v[j] = __float_as_uint(__uint_as_float(v[j]) + (__uint_as_float(w[j]) + __uint_as_float(z[j])));
                    } else {
                        tmem_ld16(taddr + c0, v);
                    }
                    if (MODE == MODE_DUMP) {
                        if (row_ok) {
#pragma unroll
                            for (int j = 0; j < 16; j++) p.dump[(size_t)row * p.nqp + qb + c0 + j] = __uint_as_float(v[j]);
                        }
                    } else if (MODE == MODE_MAX || sample) {
                        const long long g = sample ? ((long long)blockIdx.x * 4 + e) : ((blockIdx.x + i * gridDim.x) * 4 + e);  // 32-row group index
                        float mine = -INFINITY;  // lane j keeps column j's maximum, then one coalesced 64-byte store
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const float m = warp_max_f32(row_ok ? __uint_as_float(v[j]) : -INFINITY);
                            mine = (lane == j) ? m : mine;
                        }
                        if (lane < 16) p.gmax[(size_t)g * p.nqp + qb + c0 + lane] = group_max_to_ordered(mine);
                    } else if (MODE == MODE_HEAP) {
                        uint32_t mask = 0;
#pragma unroll
                        for (int j4 = 0; j4 < 4; j4++) {
                            const float4 t = *reinterpret_cast<const float4*>(&tau_s[c0 + 4 * j4]);
                            mask |= (__uint_as_float(v[4 * j4 + 0]) >= t.x ? 1u : 0u) << (4 * j4 + 0);
                            mask |= (__uint_as_float(v[4 * j4 + 1]) >= t.y ? 1u : 0u) << (4 * j4 + 1);
                            mask |= (__uint_as_float(v[4 * j4 + 2]) >= t.z ? 1u : 0u) << (4 * j4 + 2);
                            mask |= (__uint_as_float(v[4 * j4 + 3]) >= t.w ? 1u : 0u) << (4 * j4 + 3);
                        }
                        if (!row_ok) mask = 0;
                        uint32_t cols = __reduce_or_sync(0xffffffffu, mask);  // columns admitted by any row of this warp
                        while (cols) {
                            const int j = __ffs(cols) - 1;
                            cols &= cols - 1;
                            float sc = 0.f;
#pragma unroll
                            for (int jj = 0; jj < 16; jj++)
                                if (jj == j) sc = __uint_as_float(v[jj]);  // j is warp-uniform
                            const bool hit = (mask >> j) & 1u;
                            const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                            const int src = __ffs(bal) - 1;
                            int base = 0;
                            if (lane == src) base = atomicAdd(&cnt_s[c0 + j], __popc(bal));
                            base = __shfl_sync(0xffffffffu, base, src);
                            if (hit) {
                                const u64 key = make_key(sc, (uint32_t)row);  // 0 (dropped by the sort) for NaN: cannot happen, NaN >= tau is false
                                heap_s[(size_t)(c0 + j) * TC_HEAP_SLOTS + base + __popc(bal & ((1u << lane) - 1u))] = key;
                            }
                        }
                    } else {
                        // branch-free filter: 16 compares into a bit mask, one warp-uniform test per group;
                        // admissions are ~1e-3 of the scores, so the insert path below is rare
                        uint32_t mask = 0;
#pragma unroll
                        for (int j4 = 0; j4 < 4; j4++) {
                            const float4 t = *reinterpret_cast<const float4*>(&tau_s[c0 + 4 * j4]);
                            mask |= (__uint_as_float(v[4 * j4 + 0]) >= t.x ? 1u : 0u) << (4 * j4 + 0);
                            mask |= (__uint_as_float(v[4 * j4 + 1]) >= t.y ? 1u : 0u) << (4 * j4 + 1);
                            mask |= (__uint_as_float(v[4 * j4 + 2]) >= t.z ? 1u : 0u) << (4 * j4 + 2);
                            mask |= (__uint_as_float(v[4 * j4 + 3]) >= t.w ? 1u : 0u) << (4 * j4 + 3);
                        }
                        if (!row_ok) mask = 0;
                        if (__any_sync(0xffffffffu, mask != 0)) {
                            while (mask) {
                                const int j = __ffs(mask) - 1;
                                mask &= mask - 1;
                                const int c = c0 + j;
                                float s = 0.f;
#pragma unroll
                                for (int jj = 0; jj < 16; jj++)
                                    if (jj == j) s = __uint_as_float(v[jj]);
                                int slot = atomicAdd(&cnt_s[c], 1);
                                if (slot < p.cap) {
                                    p.cand[((size_t)(qb + c) * gridDim.x + blockIdx.x) * p.cap + slot] = make_key(s, (uint32_t)row);
                                } else {
                                    const int s2 = atomicAdd(&p.spill_cnt[qb + c], 1);
                                    if (s2 < TC_SPILL_CAP) p.spill[(size_t)(qb + c) * TC_SPILL_CAP + s2] = make_key(s, (uint32_t)row);
                                    else p.overflow[qb + c] = 1;
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[a]);
                if (sample) {
                    // ---- thresholds from the sample tiles of all CTAs (in-launch pre-pass) ----
                    __threadfence();  // this thread's group maxima are visible device-wide
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (et == 0) tc_grid_barrier(p.bar, gridDim.x);
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (e == 0) {  // tau0 of the queries of this block that this CTA owns: k'-th largest of the 4 * grid group maxima
                        const int G = (int)gridDim.x * 4;  // <= 1024 (checked by the host)
                        for (int c = (int)blockIdx.x; c < NP; c += (int)gridDim.x) {
                            const int qq = qb + c;
                            if (qq >= p.nq) continue;  // warp-uniform
                            uint32_t vals[32];
#pragma unroll
                            for (int u = 0; u < 32; u++) {
                                const int g = lane + 32 * u;
                                vals[u] = g < G ? __ldcg(p.gmax + (size_t)g * p.nqp + qq) : 0u;
                            }
                            uint32_t prefix = 0u;  // bit by bit: the largest v with count(values >= v) >= k'
                            for (int bit = 31; bit >= 0; bit--) {
                                const uint32_t cand = prefix | (1u << bit);
                                int cnt = 0;
#pragma unroll
                                for (int u = 0; u < 32; u++) cnt += vals[u] >= cand ? 1 : 0;
                                cnt = __reduce_add_sync(0xffffffffu, cnt);
                                if (cnt >= p.kp_sel) prefix = cand;
                            }
                            if (lane == 0) p.tau0_w[qq] = prefix != 0u ? ordered_to_score(prefix) : -INFINITY;
                        }
                    }
                    __threadfence();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (et == 0) tc_grid_barrier(p.bar, gridDim.x);
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const bool failed = *reinterpret_cast<volatile unsigned*>(p.bar + 2) != 0u;
                    for (int c = et; c < NP; c += 128) {
                        float t = (qb + c < p.nq) ? __ldcg(p.tau0_w + qb + c) : INFINITY;
                        if (failed) {  // a barrier timed out: thresholds that need no pre-pass
                            if (MODE == MODE_HEAP) t = (qb + c < p.nq) ? -INFINITY : INFINITY;
                            else {
                                t = INFINITY;  // admit nothing; every query of the block is re-run exactly by the repair
                                if (qb + c < p.nq) p.overflow[qb + c] = 1;
                            }
                        }
                        tau_s[c] = t;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                } else if (MODE == MODE_HEAP) {
                    // every warp has appended this tile's admissions (at most 128 per query: room is guaranteed because a
                    // query never starts a tile with more than 128 keys); warp e now tidies the queries c = e, e+4, ...
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    for (int c = e; c < NP; c += 4) {
                        const int n = cnt_s[c];
                        if (n > 128) {  // warp-uniform
                            u64* hc = heap_s + (size_t)c * TC_HEAP_SLOTS;
                            for (int i2 = n + lane; i2 < TC_HEAP_SLOTS; i2 += 32) hc[i2] = 0ull;
                            warp_bitonic_sort_desc(hc, TC_HEAP_SLOTS, lane);
                            if (lane == 0) {
                                tau_s[c] = key_score(hc[63]);  // the CTA's 64th best so far: nothing below it can matter
                                cnt_s[c] = 64;
                            }
                        }
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
            }
            if (MODE == MODE_HEAP) {
                // the CTA's answer: one sorted list of 64 keys per query, [query][cta][64] like the GEMV scan's lists
                for (int c = e; c < NP; c += 4) {
                    if (qb + c >= p.nq) continue;
                    const int n = cnt_s[c];
                    u64* hc = heap_s + (size_t)c * TC_HEAP_SLOTS;
                    const int pow2 = n <= 64 ? 64 : (n <= 128 ? 128 : TC_HEAP_SLOTS);
                    for (int i2 = n + lane; i2 < pow2; i2 += 32) hc[i2] = 0ull;
                    warp_bitonic_sort_desc(hc, pow2, lane);
                    u64* dst = p.lists + ((size_t)(qb + c) * gridDim.x + blockIdx.x) * 64;
                    for (int i2 = lane; i2 < 64; i2 += 32) dst[i2] = hc[i2];
                }
            }
            if (MODE == MODE_SELECT) {
                asm volatile("bar.sync 1, 128;" ::: "memory");  // every epilogue warp has finished the block
                for (int c = et; c < NP; c += 128) {
                    int n = cnt_s[c];
                    p.counts[(size_t)(qb + c) * gridDim.x + blockIdx.x] = n < p.cap ? n : p.cap;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// tau0[query] = kp-th largest of the pre-pass group maxima, by a warp-level radix select (4 passes of 8 bits
// over the ordered-uint maxima, a 256-bin histogram per warp in shared memory).  A warp per query, 8 queries
// per CTA: the 8 warps read the same 32-byte sectors of gmax ([group][query] layout).
// grid = nqp / 8, block = 256.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_tau0_kernel(const uint32_t* __restrict__ gmax, int groups, int npad, int nq, int kp,
                                                      float* __restrict__ tau0, int staged) {
    __shared__ int hist_all[8][256];
    extern __shared__ uint32_t tile[];  // staged: [8 queries][groups + 1], filled with whole 32-byte sectors of gmax
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + warp;
    int* hist = hist_all[warp];
    const int pitch = groups + 1;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // gmax is the pre-pass kernel's output
    if (staged) {
        // eight loads in flight per thread (one at a time this loop was 32 dependent L2 round trips: 10 of the kernel's 15 us)
        const int q = threadIdx.x & 7;
        for (int i0 = threadIdx.x >> 3; i0 < groups; i0 += 32 * 8) {
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + 32 * u;
                v[u] = i < groups ? __ldg(gmax + (size_t)i * npad + blockIdx.x * 8 + q) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + 32 * u;
                if (i < groups) tile[q * pitch + i] = v[u];
            }
        }
        __syncthreads();
    }
    const uint32_t* mine_vals = tile + warp * pitch;
    uint32_t prefix = 0;  // the high bits of the answer found so far
    int krem = kp;        // rank of the answer among the values that share `prefix`
    bool found = groups >= kp;
    // (a bit-by-bit counting search without shared-memory atomics was tried instead of the radix passes: 92 vs 66 us for
    // 4096 queries, 17.7 vs 11.5 us for 64)
    for (int pass = 0; pass < 4 && found; pass++) {
        const int shift = 24 - 8 * pass;
#pragma unroll
        for (int b = 0; b < 8; b++) hist[lane * 8 + b] = 0;
        __syncwarp();
        for (int i = lane; i < groups; i += 32) {
            const uint32_t v = staged ? mine_vals[i] : __ldg(gmax + (size_t)i * npad + c);
            const bool in = pass == 0 || (v >> (shift + 8)) == prefix;
            if (in) atomicAdd(&hist[(v >> shift) & 255u], 1);  // (electing one lane per bin with match.any was 5x slower)
        }
        __syncwarp();
        // lane l owns bins [8l, 8l+8); walk from the top bin down until krem values are covered
        int mine[8], lsum = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            mine[b] = hist[lane * 8 + b];
            lsum += mine[b];
        }
        int above = lsum;  // inclusive suffix sum over lanes >= this one
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_down_sync(0xffffffffu, above, off);
            if (lane + off < 32) above += o;
        }
        above -= lsum;  // values in bins of higher lanes
        const bool here = above < krem && krem <= above + lsum;
        const uint32_t bal = __ballot_sync(0xffffffffu, here);
        if (bal == 0u) {  // fewer than krem values share the prefix: cannot happen for groups >= kp, be safe
            found = false;
            break;
        }
        const int src = __ffs(bal) - 1;
        int bin = 0, newk = 0;
        if (lane == src) {
            int acc = above;
#pragma unroll
            for (int b = 7; b >= 0; b--) {
                if (acc < krem && krem <= acc + mine[b]) {
                    bin = lane * 8 + b;
                    newk = krem - acc;
                }
                acc += mine[b];
            }
        }
        bin = __shfl_sync(0xffffffffu, bin, src);
        krem = __shfl_sync(0xffffffffu, newk, src);
        prefix = (prefix << 8) | (uint32_t)bin;
        __syncwarp();
    }
    if (lane == 0) {
        // fewer than kp groups (or NaN-only groups, ordered value 0): no usable bound -> admit everything
        const uint32_t o = found ? prefix : 0u;
        tau0[c] = (c < nq && o != 0u) ? ordered_to_score(o) : ((c < nq) ? -INFINITY : INFINITY);
    }
}

// ---------------------------------------------------------------------------------------------
// gather the (CTA, query) candidate buffers of one query into a sorted top-kp list
// (same list format the GEMV scan writes: [query][1][kp], sorted descending, 0 = empty).
// The union holds ~k' * n / sampled_rows keys (about a thousand) of which kp are wanted, so instead of sorting
// it: (1) compact the valid keys into shared memory (the buffers of a query are contiguous: [query][cta][cap]),
// (2) cut them into 2 kp strided chunks, T = kp-th largest chunk maximum -- at least kp keys are >= T, so the top
// kp are among the keys >= T, typically ~1.5 kp of them -- (3) rank those survivors by counting.  Many survivors
// (clustered candidates) fall back to a bitonic sort.  grid = nq, block = 256.
// ---------------------------------------------------------------------------------------------
constexpr int GATHER_ALL_MAX = 4096;  // keys of one query held in shared memory at most; more -> overflow (exact re-run).  The launch
                                      // sizes the array (`gall`, a power of two in [1024, 4096]) from the expected candidate
                                      // count: a 4096-query batch expects ~1100 keys per query, and 2048 slots instead of
                                      // 4096 let eight CTAs share an SM instead of four
constexpr int GATHER_SURV = 1024;  // also holds the buffer maxima: nctas + 1 <= GATHER_SURV

// FUSE: the kernel goes on to finalise the query itself (canonical re-score, ranking, output, margin / guard:
// finalize_rank_emit) from the list it has just built in shared memory: one launch and one list round trip less per search
// (the 4096-query batch spent 83 + 175 us in gather + finalise, a 64-query batch 16 + 17 us of its 395).
__host__ __device__ inline size_t gather_smem_bytes(int nctas, int kp, int gall) {
    return ((size_t)(gall + GATHER_SURV + 2 * kp) * 8 + (size_t)(nctas + 1) * 4 + 15) & ~(size_t)15;
}
template <bool FUSE>
__global__ void __launch_bounds__(FUSE ? 1024 : 256) tc_gather_kernel(const u64* __restrict__ cand, const int* __restrict__ counts, int nctas,
                                                        int cap, int kp, u64* __restrict__ lists, int* __restrict__ overflow,
                                                        const int* __restrict__ spill_cnt, const u64* __restrict__ spill, int gall,
                                                        FinalizeParams f) {
    extern __shared__ __align__(16) unsigned char sraw[];
    u64* all = reinterpret_cast<u64*>(sraw);       // gall
    u64* surv = all + gall;                        // GATHER_SURV
    u64* cmax = surv + GATHER_SURV;                // 2 * kp
    int* cnt = reinterpret_cast<int*>(cmax + 2 * kp);  // nctas + 1
    // FUSE: FinalizeShared | A[kp] | sc[kp] | id[kp] | ok[kp] | qs[d] behind the gather's own arrays
    FinalizeShared* fsh = reinterpret_cast<FinalizeShared*>(sraw + gather_smem_bytes(nctas, kp, gall));
    u64* fA = reinterpret_cast<u64*>(fsh + 1);
    double* fsc = reinterpret_cast<double*>(fA + kp);
    long long* fid = reinterpret_cast<long long*>(fsc + kp);
    u64* fok = reinterpret_cast<u64*>(fid + kp);
    double* fqs = reinterpret_cast<double*>(fok + kp);
    __shared__ int s_n, s_m;
    __shared__ u64 s_T;
    const int c = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nt = blockDim.x, nwarps = nt >> 5;
    if (FUSE) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        finalize_prologue_at(f, c, fsh, fqs);  // the query does not come from the scan: widened while the scan drains
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    const u64* base = cand + (size_t)c * nctas * cap;
    for (int b = threadIdx.x; b < nctas; b += nt) cnt[b] = counts[(size_t)c * nctas + b];
    // the query's spill list is one more buffer (index nctas): normally empty
    const u64* sbase = spill + (size_t)c * TC_SPILL_CAP;
    if (threadIdx.x == 0) {
        const int sc = spill_cnt[c];
        cnt[nctas] = sc < TC_SPILL_CAP ? sc : TC_SPILL_CAP;
    }
    const int nbuf = nctas + 1;
    auto buf = [&](int b) { return b < nctas ? base + (size_t)b * cap : sbase; };
    if (threadIdx.x == 0) {
        s_n = 0;
        s_m = 0;
        s_T = 0ull;
    }
    __syncthreads();
    // how many keys does this query have?  (block reduction of the counts)
    {
        int part = 0;
        for (int b = threadIdx.x; b < nbuf; b += nt) part += cnt[b];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (lane == 0 && part) atomicAdd(&s_m, part);
    }
    __syncthreads();
    const int grand = s_m;
    __syncthreads();
    if (threadIdx.x == 0) s_m = 0;
    u64 T1 = 0ull;
    if (grand > gall) {
        // 0. too many keys for shared memory (large shards: the pre-pass samples n/128 rows, so ~140 kp keys pass):
        //    first a threshold from the buffer maxima -- T1 = kp-th largest head, at least kp keys are >= T1 and
        //    only ~1.5 % of the keys are -- then the compaction below keeps the keys >= T1 only.
        u64* heads = surv;  // nctas <= GATHER_SURV slots, free until step 2
        for (int b = warp; b < nbuf; b += nwarps) {
            const int n = cnt[b];
            const u64* src = buf(b);
            u64 m = 0ull;
            for (int i = lane; i < n; i += 32) m = umax64(m, src[i]);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) m = umax64(m, __shfl_xor_sync(0xffffffffu, m, off));
            if (lane == 0) heads[b] = m;
        }
        __syncthreads();
        for (int b = threadIdx.x; b < nbuf; b += nt) {
            const u64 hb = heads[b];
            if (hb == 0ull) continue;
            int r = 0;
            for (int j = 0; j < nbuf; j++) r += heads[j] > hb ? 1 : 0;
            if (r == kp - 1) s_T = hb;  // fewer than kp non-empty buffers: stays 0, everything is kept
        }
        __syncthreads();
        T1 = s_T;
        __syncthreads();
        if (threadIdx.x == 0) s_T = 0ull;
    }
    // 1. compaction: a warp takes 4 buffers at a time so that their loads are in flight together
    for (int b0 = warp * 4; b0 < nbuf; b0 += nwarps * 4) {
        u64 key[4];
        int nb[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            nb[u] = (b0 + u < nbuf) ? cnt[b0 + u] : 0;
            key[u] = (lane < nb[u]) ? buf(b0 + u)[lane] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            for (int i0 = 0; i0 < nb[u]; i0 += 32) {  // warp-uniform; more than one round only for counts > 32
                const u64 k = i0 == 0 ? key[u] : ((i0 + lane < nb[u]) ? buf(b0 + u)[i0 + lane] : 0ull);
                const bool valid = k != 0ull && k >= T1;
                const unsigned m = __ballot_sync(0xffffffffu, valid);
                if (m == 0u) continue;
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&s_n, __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                const int dst = pos + __popc(m & ((1u << lane) - 1u));
                if (valid && dst < gall) all[dst] = k;
            }
        }
    }
    __syncthreads();
    int total = s_n;
    if (total > gall) {
        if (threadIdx.x == 0) overflow[c] = 1;  // the caller re-runs this query through the GEMV scan
        total = gall;
    }
    u64* out = FUSE ? fA : lists + (size_t)c * kp;
    // 2. threshold from strided chunk maxima (only worth it when there are clearly more than kp keys)
    const int nch = 2 * kp;
    const u64* src = all;
    int nsrc = total;
    if (total > 4 * kp) {
        for (int j = threadIdx.x; j < nch; j += nt) {
            u64 m = 0ull;
            for (int i = j; i < total; i += nch) m = umax64(m, all[i]);
            cmax[j] = m;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < nch; j += nt) {
            const u64 mj = cmax[j];
            int r = 0;
            for (int i = 0; i < nch; i++) r += cmax[i] > mj ? 1 : 0;
            if (r == kp - 1) s_T = mj;  // keys are unique and every chunk is non-empty, so exactly one chunk has this rank
        }
        __syncthreads();
        const u64 T = s_T;
        for (int i0 = warp * 32; i0 < total; i0 += nt) {
            const u64 k = (i0 + lane < total) ? all[i0 + lane] : 0ull;
            const bool keep = k != 0ull && k >= T;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (m == 0u) continue;
            int pos = 0;
            if (lane == 0) pos = atomicAdd(&s_m, __popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            const int dst = pos + __popc(m & ((1u << lane) - 1u));
            if (keep && dst < GATHER_SURV) surv[dst] = k;
        }
        __syncthreads();
        if (s_m <= GATHER_SURV) {
            src = surv;
            nsrc = s_m;
        }  // else: too many survivors for the side array -> sort everything below
    }
    if (nsrc <= GATHER_SURV) {
        // 3a. rank by counting
        for (int i = nsrc + threadIdx.x; i < kp; i += nt) out[i] = 0ull;  // slots no key ranks into
        for (int i = threadIdx.x; i < nsrc; i += nt) {
            const u64 key = src[i];
            int r = 0;
            for (int j = 0; j < nsrc; j++) r += src[j] > key ? 1 : 0;
            if (r < kp) out[r] = key;
        }
    } else {
        // 3b. bitonic sort of everything, descending
        int pow2 = 1024;
        while (pow2 < total) pow2 <<= 1;  // <= gall (a power of two)
        for (int i = total + threadIdx.x; i < pow2; i += nt) all[i] = 0ull;
        for (int size = 2; size <= pow2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                for (int t = threadIdx.x; t < (pow2 >> 1); t += nt) {
                    const int i = bitonic_low(t, stride);
                    cmpx_desc(all, i, i + stride, (i & size) == 0);
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kp; i += nt) out[i] = all[i];
    }
    if (FUSE) {
        if (threadIdx.x == 0) {
            if (f.guard_slot) f.guard_slot[c] = -1;  // overwritten by the guard (barriers in between) if the query is queued
            if (c == 0 && f.guard_count_next) *f.guard_count_next = 0;  // ready for the next guarded search of this handle
        }
        finalize_qnorm2(f, fsh, fqs);
        __syncthreads();  // the list and |q|^2 are complete
        if (!f.xb_is_bf16) {  // the re-score reads the candidates' rows (fp32 master copy): all their lines on the way to L2 at once
            const int lines = (f.d * 4 + 127) / 128;
            for (int i = threadIdx.x; i < kp * lines; i += nt) {
                const u64 key = fA[i / lines];
                if (key != 0ull) {
                    const char* row = reinterpret_cast<const char*>(f.xb) + (size_t)key_row(key) * f.d * 4 + (size_t)(i % lines) * 128;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                }
            }
        }
        finalize_rank_emit<2>(f, c, fA, fsh, fsc, fid, fok, fqs, reinterpret_cast<unsigned char*>(all));
    }
}

__global__ void f32_to_bf16_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long count) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}

// 3xTF32: split the queries once.  dst = fp32 [2 * nqp][d]: rows [0, nqp) hold hi = tf32(q) (round to nearest), rows
// [nqp, 2 nqp) hold lo = q - hi (exact); rows beyond nq are zero (the TMA box of a partial block reads them).
__global__ void tc_split_queries_kernel(const float* __restrict__ xq, float* __restrict__ dst, int nq, int nqp, int d) {
    const long long total = (long long)nqp * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / d);
        float hi = 0.f, lo = 0.f;
        if (row < nq) {
            const float v = xq[i];
            if (fabsf(v) <= FLT_MAX) {
                uint32_t u;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
                hi = __uint_as_float(u & 0xFFFFE000u);
                lo = v - hi;
            } else {
                hi = v;
            }
        }
        dst[i] = hi;
        dst[total + i] = lo;
    }
}

// =============================================================================================
// host side
// =============================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D row-major [rows][d] tensor, box = 128 bytes of K x box_rows rows, 128-byte swizzle
cudaError_t tc_make_tmap(CUtensorMap* map, const void* base, long long rows, int d, bool is_f32, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return cudaErrorNotSupported;
    const size_t esz = is_f32 ? 4 : 2;
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)d * esz};
    cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                     gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// the same through the handle's descriptor cache (encoding two descriptors costs a few microseconds per search)
static cudaError_t tc_get_tmap(TmapCache* cache, CUtensorMap* map, const void* base, long long rows, int d, bool is_f32, int box_rows) {
    if (cache) {
        for (auto& e : cache->e)
            if (e.base == base && e.rows == rows && e.d == d && e.f32 == (int)is_f32 && e.box == box_rows) {
                *map = e.map;
                return cudaSuccess;
            }
    }
    cudaError_t err = tc_make_tmap(map, base, rows, d, is_f32, box_rows);
    if (err != cudaSuccess || !cache) return err;
    TmapCache::Entry& e = cache->e[cache->next];
    cache->next = (cache->next + 1) % 16;
    e.base = base;
    e.rows = rows;
    e.d = d;
    e.f32 = (int)is_f32;
    e.box = box_rows;
    e.map = *map;
    return cudaSuccess;
}

static cudaError_t tc_split_queries(const float* xq, void* dst, int nq, int nqp, int d, cudaStream_t st) {
    const long long total = (long long)nqp * d;
    long long blocks = (total + 255) / 256;
    tc_split_queries_kernel<<<(int)(blocks < 1024 ? blocks : 1024), 256, 0, st>>>(xq, reinterpret_cast<float*>(dst), nq, nqp, d);
    g_kernel_launches.fetch_add(1);
    return cudaGetLastError();
}

cudaError_t tc_queries_to_bf16(const float* xq, void* dst, long long count, cudaStream_t st) {
    long long blocks = (count + 255) / 256;
    f32_to_bf16_rows_kernel<<<(int)(blocks < 4096 ? blocks : 4096), 256, 0, st>>>(xq, reinterpret_cast<__nv_bfloat16*>(dst), count);
    g_kernel_launches.fetch_add(1);
    return cudaGetLastError();
}

cudaError_t tc_launch_tau0(const uint32_t* gmax, int groups, int gpow2, int nqp, int nq, int kp, float* tau0, cudaStream_t st) {
    (void)gpow2;
    if (nqp % 8) return cudaErrorInvalidValue;  // npad is a multiple of 16
    const size_t tile_bytes = (size_t)8 * (groups + 1) * 4;
    const int staged = tile_bytes <= 160 * 1024;
    if (staged && tile_bytes > 40 * 1024)
        cudaFuncSetAttribute(tc_tau0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes);
    cudaError_t e = launch_pdl(tc_tau0_kernel, dim3((unsigned)(nqp / 8)), dim3(256), staged ? tile_bytes : (size_t)0, st, gmax, groups, nqp, nq,
                               kp, tau0, staged);
    g_kernel_launches.fetch_add(1);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t tc_launch_gather(const u64* cand, const int* counts, int nctas, int nqp, int cap, int kp, int cap_total, int nq,
                             u64* lists, int* overflow, const int* spill_cnt, const u64* spill, const FinalizeParams* fin, cudaStream_t st) {
    (void)nqp;
    int gall = 1024;
    while (gall < cap_total && gall < GATHER_ALL_MAX) gall <<= 1;  // cap_total: what the plan expects per query, with slack
    size_t gs = gather_smem_bytes(nctas, kp, gall);
    if (nctas + 1 > GATHER_SURV) return cudaErrorInvalidValue;
    if (fin != nullptr) {
        FinalizeParams f = *fin;
        if (f.overflow) f.overflow = overflow;  // the scan's own flags (set by the select pass or by this very CTA)
        gs += sizeof(FinalizeShared) + (4 * (size_t)kp + (size_t)f.d) * 8 + 16;
        if (gs > 200 * 1024) return cudaErrorInvalidValue;
        static size_t have[16] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (gs > have[dev & 15]) {
            cudaError_t ae = cudaFuncSetAttribute(tc_gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs);
            if (ae != cudaSuccess) return ae;
            have[dev & 15] = gs;
        }
        // small batches: 1024 threads shorten the single CTA's critical path (one re-score round); large ones: 256 threads so
        // that several queries share an SM
        cudaError_t le = launch_pdl(tc_gather_kernel<true>, dim3((unsigned)nq), dim3(nq <= 296 ? 1024 : 256), gs, st, cand, counts, nctas, cap, kp,
                                    lists, overflow, spill_cnt, spill, gall, f);
        g_kernel_launches.fetch_add(1);
        if (le != cudaSuccess) return le;
        return cudaGetLastError();
    }
    if (gs > 200 * 1024) return cudaErrorInvalidValue;
    if (gs > 40 * 1024) cudaFuncSetAttribute(tc_gather_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs);
    tc_gather_kernel<false><<<nq, 256, gs, st>>>(cand, counts, nctas, cap, kp, lists, overflow, spill_cnt, spill, gall, FinalizeParams{});
    g_kernel_launches.fetch_add(1);
    return cudaGetLastError();
}

// largest query block (multiple of 16, <= 128) whose resident copy leaves room for a useful ring
int tc_max_queries(int d, int is_bf16) {
    const size_t esz = is_bf16 ? 2 : 4;
    if (((size_t)d * esz) % 128) return 0;  // K must be whole 128-byte chunks
    size_t per_query = (size_t)d * esz;
    int n = (int)((128 * 1024) / per_query) / 16 * 16;
    if (n > 128) n = 128;
    return n < 16 ? 0 : n;
}

static size_t tc_smem_bytes(int nk, int npad, int stages) {
    return (size_t)nk * npad * 128 + (size_t)stages * TC_STAGE_BYTES + (size_t)(2 + 2 * stages + 4 + 2 * TC_X3_TMEM_STAGES) * 8 + 32 + (size_t)npad * 8 + 1024;
}
static size_t tc_smem_bytes_heap(int nk, int npad, int stages) {
    return tc_smem_bytes(nk, npad, stages) + (size_t)npad * TC_HEAP_SLOTS * 8;
}
// 3xTF32: two resident query copies (hi, lo); the A operands live in tensor memory
static size_t tc_smem_bytes_x3(int nk, int npad, int stages, bool heap) {
    return tc_smem_bytes(nk, npad, stages) + (size_t)nk * npad * 128 + (heap ? (size_t)npad * TC_HEAP_SLOTS * 8 : 0);
}

template <typename T, int MODE, bool X3>
static cudaError_t launch_tc_mode(const CUtensorMap& tdb, const CUtensorMap& tq, const TcParams& p, int grid, size_t smem,
                                  cudaStream_t st) {
    auto kern = tc_scan_kernel<T, MODE, X3>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = launch_pdl(kern, dim3((unsigned)grid), dim3(X3 ? 512 : 256), smem, st, tdb, tq, p);
    g_kernel_launches.fetch_add(1);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_tc(bool is_bf16, int x3, const CUtensorMap& tdb, const CUtensorMap& tq, const TcParams& p, int grid,
                             size_t smem, cudaStream_t st) {
    if (x3) {
        if (is_bf16) return cudaErrorInvalidValue;
        if constexpr (MODE == MODE_SELECT) return cudaErrorInvalidValue;  // the 3xTF32 scan serves the on-chip-heap batches only
        else return launch_tc_mode<float, MODE, true>(tdb, tq, p, grid, smem, st);
    }
    return is_bf16 ? launch_tc_mode<__nv_bfloat16, MODE, false>(tdb, tq, p, grid, smem, st)
                   : launch_tc_mode<float, MODE, false>(tdb, tq, p, grid, smem, st);
}

int g_tc_max_stages = 8;  // option "tc_stages"
int g_tc_heap_max_nq = 32;  // option "tc_heap_max_nq": batches up to this size keep their top-k' on chip (MODE_HEAP); 0 = never
int g_tc_heap_pure_max_nq = 0;  // option "tc_heap_pure_max_nq": ... and up to this size without the threshold pre-pass.  Off by
                                // default: with the 16384-row sample the seeded heaps win at 1M rows for every batch size
                                // (bf16, 2 queries: 0.195 vs 0.229 ms) and lose 1-2 % at 10M rows for <= 3 queries
int g_tc_inline_pre = 1;   // option "tc_inline_pre": one-CTA kernel, <= 2 query blocks: thresholds from a sample tile per CTA INSIDE the scan launch
int g_tc_sample_rows = 0;  // option "tc_sample_rows": rows the threshold pre-pass scores at least (0 = auto)

// auto: 65536 rows for large batches (the select pass pays for every admitted candidate: 1.1 k' n / sample per query),
// 32768 for up to 256 queries, where the pre-pass itself is the larger cost, 16384 for up to 32 (the on-chip heaps take
// the extra admissions in their stride); measured: scripts/tc_tune.py and the round-1 sample sweep in DESIGN.md
int tc_sample_rows(int nq) { return g_tc_sample_rows > 0 ? g_tc_sample_rows : (nq <= 32 ? 16384 : (nq <= 256 ? 32768 : 65536)); }
static size_t tc_smem_bytes_plain(const TcPlan& pl) {
    return pl.x3 ? tc_smem_bytes_x3(pl.nk, pl.npad, pl.stages, false) : tc_smem_bytes(pl.nk, pl.npad, pl.stages);
}

// queries one 3xTF32 pass serves: both query halves, the on-chip heaps and >= 4 raw stages must fit
int tc_x3_max_queries(int d) {
    if (((size_t)d * 4) % 128) return 0;
    const int nk = (int)((size_t)d * 4 / 128);
    for (int npad = 32; npad >= 16; npad -= 16)
        if (tc_smem_bytes_x3(nk, npad, 4, true) <= 227 * 1024) return npad;
    return 0;
}
static int pick_stages(int nk, int npad) {
    int stages = g_tc_max_stages;
    while (stages > 2 && tc_smem_bytes(nk, npad, stages) > 226 * 1024) stages--;
    return stages;
}

size_t tc_workspace_bytes(const TcPlan& pl) { return pl.off_end; }

// Plan a launch set for `nq` queries over `n` rows: blocks of npad <= tc_max_queries queries are walked
// inside one persistent launch, so the fixed costs (launches, prologues, pre-pass) are paid once.
cudaError_t tc_plan(long long n, int d, int is_bf16, int nq, int kp, int sm_count, int x3, TcPlan* pl) {
    const size_t esz = is_bf16 ? 2 : 4;
    const int nbmax = x3 ? tc_x3_max_queries(d) : tc_max_queries(d, is_bf16);
    if (nbmax == 0 || nq <= 0) return cudaErrorInvalidValue;
    if (x3 && (is_bf16 || kp != 64 || nq > nbmax)) return cudaErrorInvalidValue;  // one on-chip-heap block of fp32 rows
    pl->x3 = x3 ? 1 : 0;
    pl->npad = nq >= nbmax ? nbmax : (nq + 15) / 16 * 16;
    pl->nblocks = (nq + pl->npad - 1) / pl->npad;
    pl->nqp = pl->nblocks * pl->npad;
    pl->nk = (int)((size_t)d * esz / 128);
    pl->stages = pick_stages(pl->nk, pl->npad);
    pl->smem = tc_smem_bytes(pl->nk, pl->npad, pl->stages);
    if (pl->smem > 227 * 1024) return cudaErrorInvalidValue;
    // small batches with k' = 64: running top-k' per (CTA, query) in shared memory, if >= 4 ring stages still fit
    pl->heap = 0;
    if (x3) {
        int stages = g_tc_max_stages;
        while (stages > 2 && tc_smem_bytes_x3(pl->nk, pl->npad, stages, true) > 226 * 1024) stages--;
        if (stages < 4 || tc_smem_bytes_x3(pl->nk, pl->npad, stages, true) > 227 * 1024) return cudaErrorInvalidValue;
        pl->heap = nq <= g_tc_heap_pure_max_nq ? 1 : 2;
        pl->stages = stages;
        pl->smem = tc_smem_bytes_x3(pl->nk, pl->npad, stages, true);
    } else if (kp == 64 && nq <= g_tc_heap_max_nq && pl->nblocks == 1) {
        int stages = g_tc_max_stages;
        while (stages > 2 && tc_smem_bytes_heap(pl->nk, pl->npad, stages) > 226 * 1024) stages--;
        if (stages >= 4 && tc_smem_bytes_heap(pl->nk, pl->npad, stages) <= 227 * 1024) {
            pl->heap = nq <= g_tc_heap_pure_max_nq ? 1 : 2;  // 2: thresholds from the pre-pass first
            pl->stages = stages;
            pl->smem = tc_smem_bytes_heap(pl->nk, pl->npad, stages);
        }
    }
    pl->ntiles = (n + TC_BM - 1) / TC_BM;
    pl->grid = (int)(pl->ntiles < sm_count ? pl->ntiles : sm_count);
    // pre-pass sample: every `stride`-th tile, at least 512 tiles (or all of them)
    long long want = pl->ntiles / 128;
    if (want < tc_sample_rows(nq) / TC_BM) want = tc_sample_rows(nq) / TC_BM;
    // one-CTA kernel (<= 128 queries per block): the pre-pass is latency-bound (~10 us per tile per CTA), so up to 1.2 million
    // rows it takes one tile per SM (18 944 rows on 148 SMs) instead of a second round for a few CTAs; the select pass
    // absorbs the 1.7x admissions unnoticed at these batch sizes
    if (g_tc_sample_rows == 0 && nq > 32 && want > sm_count && pl->ntiles / 128 <= sm_count) want = sm_count;
    if (want > pl->ntiles) want = pl->ntiles;
    pl->pre_stride = pl->ntiles / want;
    pl->pre_tiles = (pl->ntiles + pl->pre_stride - 1) / pl->pre_stride;
    pl->pre_grid = (int)(pl->pre_tiles < sm_count ? pl->pre_tiles : sm_count);
    pl->groups = (int)(pl->pre_tiles * 4);
    // In-launch pre-pass (tc_scan_kernel, inline_pre): every CTA of the persistent grid scores ONE sampled tile first and the
    // thresholds are agreed through two grid barriers -- needs the full grid resident (one CTA per SM), k' <= 4 * grid group
    // maxima held in registers by one warp (<= 1024), and a CTA per query of a block.
    pl->inline_pre = 0;
    if (g_tc_inline_pre && g_tc_sample_rows == 0 && (pl->heap == 2 || (!pl->heap && pl->nblocks <= 2)) && pl->grid == sm_count &&
        pl->grid * 4 <= 1024 && pl->grid * 4 >= 2 * kp && pl->grid >= pl->npad && pl->ntiles >= 2 * (long long)pl->grid &&
        pl->ntiles / 128 <= pl->grid) {  // larger shards want a larger sample than one tile per CTA (the separate pre-pass scores
                                         // n / 128 rows: at 10M rows the 18 944-row sample cost more in admissions than it saved)
        pl->inline_pre = 1;
        pl->pre_tiles = pl->grid;
        pl->pre_stride = pl->ntiles / pl->grid;
        pl->pre_grid = pl->grid;
        pl->groups = pl->grid * 4;
    }
    int g2 = 1;
    while (g2 < pl->groups) g2 <<= 1;
    pl->gpow2 = g2;
    pl->kp = kp;
    // candidate capacity per (CTA, query): the pre-pass threshold admits about 1.1 * kp * n / sampled_rows
    // rows per query; give every CTA 4x its share plus slack (an overflow only costs a re-run)
    double expect = 1.1 * kp * (double)n / ((double)pl->groups * 32.0) / pl->grid;
    int cap = 32;
    while (cap < 4.0 * expect + 24.0 && cap < 1024) cap <<= 1;
    pl->cap = cap;
    pl->cap_total = (int)(3.0 * expect * pl->grid) + 256;  // keys per query the gather should hold in shared memory (3x the expectation)
    // workspace layout
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = (off + bytes + 255) & ~(size_t)255;
        return o;
    };
    pl->off_gmax = take((size_t)pl->groups * pl->nqp * 4);
    pl->off_tau0 = take((size_t)pl->nqp * 4);
    pl->off_counts = take((size_t)pl->grid * pl->nqp * 4);
    pl->off_overflow = take((size_t)pl->nqp * 4);
    pl->off_spill_cnt = take((size_t)pl->nqp * 4);
    pl->off_cand = take((size_t)pl->grid * pl->nqp * pl->cap * 8);
    pl->off_spill = take((size_t)pl->nqp * TC_SPILL_CAP * 8);
    pl->off_qbf16 = take(x3 ? (size_t)2 * pl->nqp * d * 4 : (size_t)pl->nqp * d * 2);  // bf16 copy of the queries / their hi + lo split
    pl->off_end = off;
    return cudaSuccess;
}

static cudaError_t tc_prepare(const TcArgs& a, const TcPlan& pl, unsigned char* ws, CUtensorMap* tdb, CUtensorMap* tq,
                              TcParams* p, cudaStream_t st) {
    cudaError_t e = tc_get_tmap(a.tmaps, tdb, a.xb, a.n, a.d, !a.is_bf16, TC_BM);
    if (e != cudaSuccess) return e;
    const void* qsrc = a.xq;
    long long qrows = a.nq;
    if (a.is_bf16) {
        void* qb = ws + pl.off_qbf16;
        if ((e = tc_queries_to_bf16(a.xq, qb, (long long)a.nq * a.d, st)) != cudaSuccess) return e;
        qsrc = qb;
    } else if (pl.x3) {
        void* qs = ws + pl.off_qbf16;
        if ((e = tc_split_queries(a.xq, qs, a.nq, pl.nqp, a.d, st)) != cudaSuccess) return e;
        qsrc = qs;
        qrows = 2LL * pl.nqp;
    }
    if ((e = tc_get_tmap(a.tmaps, tq, qsrc, qrows, a.d, !a.is_bf16, pl.npad)) != cudaSuccess) return e;
    *p = TcParams{};
    p->n = a.n;
    p->d = a.d;
    p->nq = a.nq;
    p->npad = pl.npad;
    p->nblocks = pl.nblocks;
    p->nqp = pl.nqp;
    p->nk = pl.nk;
    p->stages = pl.stages;
    p->gmax = reinterpret_cast<uint32_t*>(ws + pl.off_gmax);
    p->tau0 = reinterpret_cast<const float*>(ws + pl.off_tau0);
    p->cand = reinterpret_cast<u64*>(ws + pl.off_cand);
    p->cap = pl.cap;
    p->counts = reinterpret_cast<int*>(ws + pl.off_counts);
    p->overflow = reinterpret_cast<int*>(ws + pl.off_overflow);
    p->spill_cnt = reinterpret_cast<int*>(ws + pl.off_spill_cnt);
    p->spill = reinterpret_cast<u64*>(ws + pl.off_spill);
    return cudaSuccess;
}

// Scan `a.nq` queries with the tensor-core path; writes one sorted kp-list per query into `lists`
// ([nq][kp], the format finalize_kernel takes with L = 1) and sets overflow[q] = 1 where the result
// must not be trusted.  `ws` is a device workspace of tc_workspace_bytes(pl).
cudaError_t tc_scan_block(const TcArgs& a, const TcPlan& pl, unsigned char* ws, cudaStream_t st) {
    CUtensorMap tdb, tq;
    TcParams p;
    cudaError_t e = tc_prepare(a, pl, ws, &tdb, &tq, &p, st);
    if (e != cudaSuccess) return e;
    const bool inline_pre = pl.inline_pre && a.bar != nullptr;
    if (inline_pre) {
        p.inline_pre = 1;
        p.pre_stride = pl.pre_stride;
        p.tau0_w = reinterpret_cast<float*>(ws + pl.off_tau0);
        p.bar = a.bar;
        p.kp_sel = pl.kp;
    }
    if (pl.heap) {
        // small batch: the CTAs keep their own top-k' per query on chip and write [nq][grid][64] lists
        if (pl.heap == 2 && inline_pre) {
            // thresholds from inside the scan launch
        } else if (pl.heap == 2) {  // pre-pass thresholds first: the running top-k' then hardly ever needs a sort
            p.ntiles = pl.pre_tiles;
            p.tile_stride = pl.pre_stride;
            if ((e = launch_tc<MODE_MAX>(a.is_bf16, pl.x3, tdb, tq, p, pl.pre_grid, tc_smem_bytes_plain(pl), st)) != cudaSuccess) return e;
            if ((e = tc_launch_tau0(p.gmax, pl.groups, pl.gpow2, pl.nqp, a.nq, pl.kp, reinterpret_cast<float*>(ws + pl.off_tau0), st)) !=
                cudaSuccess)
                return e;
        } else {
            p.tau0 = nullptr;
        }
        p.ntiles = pl.ntiles;
        p.tile_stride = 1;
        p.lists = reinterpret_cast<u64*>(a.lists);
        if ((e = launch_tc<MODE_HEAP>(a.is_bf16, pl.x3, tdb, tq, p, pl.grid, pl.smem, st)) != cudaSuccess) return e;
        if (a.overflow_out) e = cudaMemsetAsync(a.overflow_out, 0, (size_t)a.nq * 4, st);  // this mode cannot overflow
        return e;
    }
    // 1. threshold pre-pass over the sampled tiles (separate launches unless it runs inside the select launch)
    if (!inline_pre) {
        p.ntiles = pl.pre_tiles;
        p.tile_stride = pl.pre_stride;
        e = launch_tc<MODE_MAX>(a.is_bf16, 0, tdb, tq, p, pl.pre_grid, pl.smem, st);
        if (e != cudaSuccess) return e;
        if ((e = tc_launch_tau0(p.gmax, pl.groups, pl.gpow2, pl.nqp, a.nq, pl.kp, reinterpret_cast<float*>(ws + pl.off_tau0), st)) !=
            cudaSuccess)
            return e;
    }
    // overflow flags and spill counters start at zero (adjacent in the workspace)
    if ((e = cudaMemsetAsync(ws + pl.off_overflow, 0, pl.off_cand - pl.off_overflow, st)) != cudaSuccess) return e;
    // 2. selection pass over every tile, all query blocks in one persistent launch
    p.ntiles = pl.ntiles;
    p.tile_stride = 1;
    e = launch_tc<MODE_SELECT>(a.is_bf16, 0, tdb, tq, p, pl.grid, pl.smem, st);
    if (e != cudaSuccess) return e;
    // 3. per query: gather + sort -> top-kp list
    if ((e = tc_launch_gather(p.cand, p.counts, pl.grid, pl.nqp, pl.cap, pl.kp, pl.cap_total, a.nq,
                              reinterpret_cast<u64*>(a.lists), p.overflow, p.spill_cnt, p.spill, a.fin, st)) != cudaSuccess)
        return e;
    if (a.overflow_out)
        e = cudaMemcpyAsync(a.overflow_out, ws + pl.off_overflow, (size_t)a.nq * 4, cudaMemcpyDeviceToDevice, st);
    return e;
}

// tests: raw tensor-core scores of every row against the queries, [n][nqp] fp32
cudaError_t tc_dump_scores(const TcArgs& a, const TcPlan& pl, unsigned char* ws, float* out, cudaStream_t st) {
    CUtensorMap tdb, tq;
    TcParams p;
    cudaError_t e = tc_prepare(a, pl, ws, &tdb, &tq, &p, st);
    if (e != cudaSuccess) return e;
    p.ntiles = pl.ntiles;
    p.tile_stride = 1;
    p.dump = out;
    return launch_tc<MODE_DUMP>(a.is_bf16, pl.x3, tdb, tq, p, pl.grid, tc_smem_bytes_plain(pl), st);
}

}  // namespace evs
