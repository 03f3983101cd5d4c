// evs_tc2.cu -- CTA-pair tensor-core scan for LARGE query batches (BASELINE config 3: 1M x 512 bf16, 4096 queries).
//
// Same job as evs_tc.cu -- the inner loop of index.search() (/root/reference/oldapp.py:2005, :2112) as a dense
// contraction S[row][query] with the top-k' selection fused into the epilogue -- re-shaped for the regime where
// the batch is much larger than one MMA's N.  evs_tc.cu walks the whole database once per block of 128 queries:
// at 4096 queries that is 32 passes over HBM and the kernel is HBM-bound (ncu: 77 % DRAM, 41 % tensor pipe).
// Here:
//   * two CTAs on the two SMs of a TPC form a pair (cluster of 2) and issue ONE tcgen05.mma.cta_group::2 per
//     K step: M = 256 database rows (128 staged by each CTA), N = up to 256 queries (half resident in each CTA's
//     shared memory), accumulators 128 lanes x N columns in each CTA's TMEM, double-buffered (2 x 256 = all 512
//     columns).  Every database byte staged into shared memory now meets 256 queries instead of 128: half the
//     L2 -> SM traffic and half the shared-memory operand bandwidth per flop.
//   * work is cut into items (database slice, query block), numbered slice-major and dealt round-robin to the
//     74 pairs, so the pairs that run at the same time share a handful of slices (a few MB each): the database
//     streams from HBM once per launch and is re-read from the 126 MB L2 by the other query blocks.
//   * the query block of the next item is re-loaded (128 KB per CTA, from L2) only when it changes.  (Reloading it K
//     chunk by K chunk behind the MMAs, with per-chunk barriers, was measured: the extra waits and commits in the
//     MMA issue loop cost more (+4 % on the select kernel) than the ~3 % drain they remove.)
// Warp roles per CTA (384 threads): warp 0 TMA producer (one lane), warp 1 MMA issuer (one lane, leader CTA
// only), warp 2 TMEM alloc/dealloc, warps 4-11 epilogue: two warps per TMEM lane quarter (e = warp & 3), which
// take alternate 32-column groups, so that every SM sub-partition has two epilogue warps to overlap the
// tcgen05.ld latency of one with the filter arithmetic of the other (one warp per quarter left the MMA issuer
// waiting on acc_empty 40 % of the time: ncu, profiles/r01_ncu_tc2_select_v1.txt).
// Barriers: full[s] and q_full collect the TMA bytes of BOTH CTAs on the leader's barrier (cp.async.bulk.tensor
// .cta_group::2 with the leader's barrier address); empty[s], q_empty and acc_full[a] are signalled in both CTAs
// by tcgen05.commit ... multicast::cluster; acc_empty[a] lives in the leader and counts the 16 epilogue warps of
// the pair (the peer's arrive remotely through mapa + mbarrier.arrive.shared::cluster), each as soon as its last
// tcgen05.ld of the tile has completed -- before it filters that last group.
// SELECT epilogue: scores are packed to bf16x2 (cvt.rn.bf16x2.f32) and compared two at a time with the
// bf16-rounded thresholds (set.ge.u32.bf16x2; rounding is monotonic, so score >= tau implies bf16(score) >=
// bf16(tau): no false negatives); the 16 compare results are folded into a 32-bit column mask with one LOP3
// each.  Columns flagged by any lane are then visited warp-uniformly: exact fp32 test, one shared-memory
// atomic per (warp, column) for the slots, 8-byte key stores.
// Epilogue modes and the threshold scheme (pre-pass maxima -> tau0 -> branch-free filter -> candidate buffers ->
// gather) are those of evs_tc.cu; exactness never depends on the data (overflow -> GEMV re-run by the caller).
#include <cuda.h>
#include <float.h>

#include "evs_internal.h"
#include "evs_common.cuh"
#include "evs_tc_common.cuh"

namespace evs {

struct Tc2Params {
    long long n;           // rows in the shard
    int d;
    int nq;                // valid queries
    int npad;              // queries per block = N of the MMA (multiple of 32, <= 256)
    int half;              // npad / 2: query rows resident per CTA
    int nqb;               // query blocks
    int nqp;               // nqb * npad: padded query count, row pitch of the per-query arrays
    int nk;                // 128-byte K chunks per row
    int stages;            // ring depth
    long long ntiles;      // pair-tiles (256 rows) in the walked list; entry j is pair-tile j * tile_stride
    long long tile_stride;
    int slice_tiles;       // pair-tiles per slice
    long long nslices;
    int block_major;       // item order: 0 = slice-major, dealt round-robin (select pass: concurrent pairs share slices through L2);
                           // 1 = block-major, a contiguous run per pair (pre-pass: its 64 MB sample is L2-resident anyway, and a
                           // pair then keeps the same query block for ~all of its items instead of re-loading it every 3 tiles)
    // MODE_MAX
    uint32_t* gmax;        // [ntiles * 8][nqp] ordered-uint maxima per 32-row group
    // MODE_SELECT
    const float* tau0;     // [nqp]
    u64* cand;             // [nqp][gridDim.x][cap]: the buffers of one query are contiguous for the gather
    int cap;
    int* counts;           // [nqp][gridDim.x], zeroed before the launch; persists across the items of a CTA
    int* overflow;         // [nqp]
    int* spill_cnt;        // [nqp]
    u64* spill;            // [nqp][TC_SPILL_CAP]: extra keys of full (CTA, query) buffers (clustered rows)
    // MODE_DUMP
    float* dump;           // [n][nqp]
};

constexpr int TC2_THREADS = 384;      // 4 control warps + 8 epilogue warps
constexpr int TC2_EPI_THREADS = 256;

// 32 scores of one row (32 consecutive queries) against the bf16-rounded thresholds of those queries (64 bytes in
// shared memory): bit b < 16 of the result <-> column 2b, bit 16 + b <-> column 2b + 1
__device__ __forceinline__ uint32_t prefilter_mask_bf16(const uint32_t (&v)[32], const uint4* tau_b) {
    uint32_t mask = 0;
#pragma unroll
    for (int i4 = 0; i4 < 4; i4++) {
        const uint4 t4 = tau_b[i4];
        const uint32_t tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i4 * 4 + u;  // column pair (2i, 2i + 1)
            uint32_t pk, r;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(__uint_as_float(v[2 * i + 1])), "f"(__uint_as_float(v[2 * i])));
            asm("set.ge.u32.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(pk), "r"(tt[u]));
            mask |= r & ((1u << i) | (1u << (16 + i)));
        }
    }
    return mask;
}

// v[j] for a warp-uniform j: a jump table instead of 32 predicated moves
__device__ __forceinline__ float pick_uniform(const uint32_t (&v)[32], int j) {
    uint32_t r;
    switch (j) {
#define EVS_PICK(J) case J: r = v[J]; break;
        EVS_PICK(0) EVS_PICK(1) EVS_PICK(2) EVS_PICK(3) EVS_PICK(4) EVS_PICK(5) EVS_PICK(6) EVS_PICK(7)
        EVS_PICK(8) EVS_PICK(9) EVS_PICK(10) EVS_PICK(11) EVS_PICK(12) EVS_PICK(13) EVS_PICK(14) EVS_PICK(15)
        EVS_PICK(16) EVS_PICK(17) EVS_PICK(18) EVS_PICK(19) EVS_PICK(20) EVS_PICK(21) EVS_PICK(22) EVS_PICK(23)
        EVS_PICK(24) EVS_PICK(25) EVS_PICK(26) EVS_PICK(27) EVS_PICK(28) EVS_PICK(29) EVS_PICK(30)
#undef EVS_PICK
        default: r = v[31]; break;
    }
    return __uint_as_float(r);
}

template <typename T, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
tc2_scan_kernel(const __grid_constant__ CUtensorMap tm_db, const __grid_constant__ CUtensorMap tm_q, Tc2Params p) {
    constexpr bool TF32 = sizeof(T) == 4;
    constexpr int EC = 128 / sizeof(T);  // elements per 128-byte chunk
    constexpr int KSTEP_BYTES = 32;      // one MMA consumes 32 bytes of K per row
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages, NK = p.nk, NP = p.npad, H = p.half;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const long long pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const long long nitems = p.nslices * p.nqb;
    // step -> item (slice, query block) of this pair; false when the pair has no more items
    const long long run = (nitems + npairs - 1) / npairs;  // block-major: items [pair * run, (pair + 1) * run)
    auto item_at = [&](long long step, long long& sl, int& bl) -> bool {
        if (p.block_major) {
            const long long it = pair * run + step;
            if (step >= run || it >= nitems) return false;
            bl = (int)(it / p.nslices);
            sl = it % p.nslices;
        } else {
            const long long it = pair + step * npairs;
            if (it >= nitems) return false;
            sl = it / p.nqb;
            bl = (int)(it % p.nqb);
        }
        return true;
    };

    // shared memory: [resident query half: NK chunks x H rows x 128 B][ring: S x 16 KB][barriers][tmem base][tau][cnt]
    if ((smem_u32(smem) & 1023u) != 0u) __trap();  // 128B-swizzled operands need 1024-byte aligned bases
    unsigned char* q_smem = smem;
    unsigned char* ring = q_smem + (size_t)NK * H * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)S * TC_STAGE_BYTES);
    uint64_t* q_full = bars;             // leader's: the query block (both halves) has landed
    uint64_t* q_empty = bars + 1;        // both: every MMA that reads the resident query block has completed
    uint64_t* full = bars + 2;           // S, leader's: stage s of BOTH CTAs has landed
    uint64_t* empty = full + S;          // S, both
    uint64_t* acc_full = empty + S;      // 2, both
    uint64_t* acc_empty = acc_full + 2;  // 2, leader's, 16 arrivals
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* tau_s = reinterpret_cast<float*>(tmem_base_smem + 4);
    int* cnt_s = reinterpret_cast<int*>(tau_s + NP);
    __nv_bfloat16* tau_b = reinterpret_cast<__nv_bfloat16*>(cnt_s + NP);  // 16-byte aligned: NP is a multiple of 32

    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(2 * NP)) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_db);
        prefetch_tmap(&tm_q);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 16);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_base_smem, tmem_cols);
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated, before any remote arrive or multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (lane == 0) {
            const uint32_t q_full_leader = mapa_u32(smem_u32(q_full), 0);
            int stage = 0;
            uint32_t phase = 0;
            int cur_b = -1;
            uint32_t qloads = 0;
            for (long long step = 0;; step++) {
                long long s;
                int b;
                if (!item_at(step, s, b)) break;
                if (b != cur_b) {
                    if (qloads > 0) mbar_wait(q_empty, (qloads - 1) & 1u);  // the MMAs of the previous block are done with it
                    if (leader) mbar_arrive_expect_tx(q_full, (uint32_t)(2 * NK * H * 128));
                    for (int c = 0; c < NK; c++)
                        tma_load_2d_pair(q_smem + (size_t)c * H * 128, &tm_q, c * EC, b * NP + (int)cta_rank * H, q_full_leader);
                    cur_b = b;
                    qloads++;
                }
                const long long t0 = s * p.slice_tiles;
                const long long t1 = (t0 + p.slice_tiles < p.ntiles) ? t0 + p.slice_tiles : p.ntiles;
                for (long long t = t0; t < t1; t++) {
                    const int row0 = (int)(t * p.tile_stride * (2 * TC_BM) + cta_rank * TC_BM);
                    for (int c = 0; c < NK; c++) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        if (leader) mbar_arrive_expect_tx(&full[stage], 2 * TC_STAGE_BYTES);
                        tma_load_2d_pair(ring + (size_t)stage * TC_STAGE_BYTES, &tm_db, c * EC, row0,
                                         mapa_u32(smem_u32(&full[stage]), 0));
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
        // ================= MMA issuer (leader CTA only) =================
        if (lane == 0 && leader) {
            const uint32_t idesc = make_idesc(TF32, 2 * TC_BM, NP);
            int stage = 0;
            uint32_t phase = 0;
            long long tcount = 0;  // tile counter: accumulator buffer and its phase
            int cur_b = -1;
            uint32_t qloads = 0;
            for (long long step = 0;; step++) {
                long long s;
                int b;
                if (!item_at(step, s, b)) break;
                if (b != cur_b) {
                    mbar_wait(q_full, qloads & 1u);
                    tc_fence_after();
                    cur_b = b;
                    qloads++;
                }
                const long long t0 = s * p.slice_tiles;
                const long long t1 = (t0 + p.slice_tiles < p.ntiles) ? t0 + p.slice_tiles : p.ntiles;
                for (long long t = t0; t < t1; t++, tcount++) {
                    const int a = (int)(tcount & 1);
                    mbar_wait(&acc_empty[a], (uint32_t)((tcount >> 1) & 1) ^ 1u);  // both epilogues have drained it
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(a * NP);
                    for (int c = 0; c < NK; c++) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(ring + (size_t)stage * TC_STAGE_BYTES);
                        const uint32_t b_addr = smem_u32(q_smem + (size_t)c * H * 128);
#pragma unroll
                        for (int k = 0; k < 128 / KSTEP_BYTES; k++) {
                            umma_pair<TF32>(d_tmem, smem_desc_sw128(a_addr + k * KSTEP_BYTES),
                                            smem_desc_sw128(b_addr + k * KSTEP_BYTES), idesc, (uint32_t)((c | k) != 0));
                        }
                        umma_commit_pair(&empty[stage], 3);  // frees the ring slot in both CTAs
                        if (++stage == S) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    umma_commit_pair(&acc_full[a], 3);  // accumulator complete -> both epilogues
                }
                long long s_next;
                int b_next;
                if (item_at(step + 1, s_next, b_next) && b_next != b) umma_commit_pair(q_empty, 3);  // block may be overwritten
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ================= epilogue (both CTAs), 8 warps =================
        const int e = warp & 3;                 // TMEM lane quarter of this warp
        const int h = (warp - 4) >> 2;          // which of the quarter's two warps: takes 32-column groups h, h+2, ...
        const int row_in_tile = (int)cta_rank * TC_BM + e * 32 + lane;
        const int et = threadIdx.x - 128;       // 0..255 among the epilogue threads
        const int ngroups = NP / 32;
        const uint32_t acc_empty_leader0 = mapa_u32(smem_u32(&acc_empty[0]), 0);
        const uint32_t acc_empty_leader1 = mapa_u32(smem_u32(&acc_empty[1]), 0);
        const uint32_t lt_mask = (1u << lane) - 1u;
        long long tcount = 0;
        int cur_b = -1;
        for (long long step = 0;; step++) {
            long long s;
            int b;
            if (!item_at(step, s, b)) break;
            const int qb = b * NP;
            if (MODE == MODE_SELECT && b != cur_b) {
                asm volatile("bar.sync 1, 256;" ::: "memory");  // all epilogue warps have left the previous block
                if (cur_b >= 0) {
                    for (int c = et; c < NP; c += TC2_EPI_THREADS) {
                        const int n = cnt_s[c];
                        p.counts[(size_t)(cur_b * NP + c) * gridDim.x + blockIdx.x] = n < p.cap ? n : p.cap;
                    }
                }
                for (int c = et; c < NP; c += TC2_EPI_THREADS) {
                    const float tq = (qb + c < p.nq) ? p.tau0[qb + c] : INFINITY;
                    tau_s[c] = tq;
                    tau_b[c] = __float2bfloat16_rn(tq);
                    cnt_s[c] = p.counts[(size_t)(qb + c) * gridDim.x + blockIdx.x];
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            cur_b = b;
            const long long t0 = s * p.slice_tiles;
            const long long t1 = (t0 + p.slice_tiles < p.ntiles) ? t0 + p.slice_tiles : p.ntiles;
            for (long long t = t0; t < t1; t++, tcount++) {
                const int a = (int)(tcount & 1);
                const long long row = t * p.tile_stride * (2 * TC_BM) + row_in_tile;
                const bool row_ok = row < p.n;
                mbar_wait(&acc_full[a], (uint32_t)((tcount >> 1) & 1));
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(a * NP);
                uint32_t v[2][32];
                if (h < ngroups) tmem_ld32_nowait(taddr + h * 32, v[0]);
                else if (lane == 0) {  // fewer groups than warps in this quarter: nothing to read
                    tc_fence_before();
                    mbar_arrive_cluster_relaxed(a ? acc_empty_leader1 : acc_empty_leader0);
                }
#pragma unroll
                for (int gi = 0; gi < 4; gi++) {  // at most 8 groups per tile, 4 per warp
                    const int g = h + 2 * gi;
                    if (g < ngroups) {
                        tmem_ld_wait();
                        if (g + 2 < ngroups) {
                            tmem_ld32_nowait(taddr + (g + 2) * 32, v[(gi + 1) & 1]);  // in flight while this group is filtered
                        } else {
                            // this warp's last read of the accumulator is complete: hand it back before filtering
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster_relaxed(a ? acc_empty_leader1 : acc_empty_leader0);
                        }
                        const uint32_t(&vc)[32] = v[gi & 1];
                        const int c0 = g * 32;
                        if (MODE == MODE_DUMP) {
                            if (row_ok) {
#pragma unroll
                                for (int j = 0; j < 32; j++) p.dump[(size_t)row * p.nqp + qb + c0 + j] = __uint_as_float(vc[j]);
                            }
                        } else if (MODE == MODE_MAX) {
                            // column maxima over the warp's 32 rows; lane j keeps column j's, then ONE coalesced store
                            // (32 predicated single-lane stores cost ptxas-dependent branch/address code per column:
                            // measured 2.4x on the whole pre-pass)
                            const long long gr = (t * 2 + cta_rank) * 4 + e;  // 32-row group index in the walked list
                            float mine = -INFINITY;
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                const float m = warp_max_f32(row_ok ? __uint_as_float(vc[j]) : -INFINITY);
                                mine = (lane == j) ? m : mine;
                            }
                            p.gmax[(size_t)gr * p.nqp + qb + c0 + lane] = group_max_to_ordered(mine);
                        } else {
                            uint32_t mask = prefilter_mask_bf16(vc, reinterpret_cast<const uint4*>(tau_b + c0));
                            if (!row_ok) mask = 0;
                            uint32_t cols = __reduce_or_sync(0xffffffffu, mask);  // columns flagged by any row of this warp
                            while (cols) {
                                const int bit = __ffs(cols) - 1;
                                cols &= cols - 1;
                                const int j = bit < 16 ? 2 * bit : 2 * (bit - 16) + 1;
                                const int c = c0 + j;
                                const float sc = pick_uniform(vc, j);
                                const bool hit = ((mask >> bit) & 1u) && sc >= tau_s[c];  // exact test
                                const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                                if (bal) {
                                    const int src = __ffs(bal) - 1;
                                    int base = 0;
                                    if (lane == src) base = atomicAdd(&cnt_s[c], __popc(bal));
                                    base = __shfl_sync(0xffffffffu, base, src);
                                    if (hit) {
                                        const int slot = base + __popc(bal & lt_mask);
                                        if (slot < p.cap) {
                                            p.cand[((size_t)(qb + c) * gridDim.x + blockIdx.x) * p.cap + slot] = make_key(sc, (uint32_t)row);
                                        } else {
                                            const int s2 = atomicAdd(&p.spill_cnt[qb + c], 1);
                                            if (s2 < TC_SPILL_CAP) p.spill[(size_t)(qb + c) * TC_SPILL_CAP + s2] = make_key(sc, (uint32_t)row);
                                            else p.overflow[qb + c] = 1;
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        if (MODE == MODE_SELECT && cur_b >= 0) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int c = et; c < NP; c += TC2_EPI_THREADS) {
                const int n = cnt_s[c];
                p.counts[(size_t)(cur_b * NP + c) * gridDim.x + blockIdx.x] = n < p.cap ? n : p.cap;
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();  // the peer may still be reading its TMEM / arriving on our barriers
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, tmem_cols);
    }
}

// =============================================================================================
// host side
// =============================================================================================
static size_t tc2_smem_bytes(int nk, int half, int stages) {
    return (size_t)nk * half * 128 + (size_t)stages * TC_STAGE_BYTES + (size_t)(2 + 2 * stages + 4) * 8 + 16 + (size_t)half * 20;
}

// query rows resident per CTA: the largest multiple of 16 (<= 128) that leaves room for >= 4 ring stages
int tc2_max_half(int d, int is_bf16) {
    const size_t esz = is_bf16 ? 2 : 4;
    if (((size_t)d * esz) % 128) return 0;
    const int nk = (int)((size_t)d * esz / 128);
    for (int half = 128; half >= 16; half -= 16)
        if (tc2_smem_bytes(nk, half, 4) <= 227 * 1024) return half;
    return 0;
}

int g_tc2_slice_tiles = 0;  // option "tc2_slice_tiles" (0 = auto)

cudaError_t tc2_plan(long long n, int d, int is_bf16, int nq, int kp, int sm_count, Tc2Plan* pl) {
    const size_t esz = is_bf16 ? 2 : 4;
    const int hmax = tc2_max_half(d, is_bf16);
    if (hmax == 0 || nq <= 0 || sm_count < 2) return cudaErrorInvalidValue;
    const int nmax = 2 * hmax;
    pl->nqb = (nq + nmax - 1) / nmax;
    pl->npad = ((nq + pl->nqb - 1) / pl->nqb + 31) / 32 * 32;  // balanced blocks, halves stay multiples of 16
    pl->half = pl->npad / 2;
    pl->nqp = pl->nqb * pl->npad;
    pl->nk = (int)((size_t)d * esz / 128);
    int stages = g_tc_max_stages;
    while (stages > 2 && tc2_smem_bytes(pl->nk, pl->half, stages) > 227 * 1024) stages--;
    pl->stages = stages;
    pl->smem = tc2_smem_bytes(pl->nk, pl->half, stages);
    if (pl->smem > 227 * 1024) return cudaErrorInvalidValue;
    const int npairs_max = sm_count / 2;
    pl->ntiles = (n + 2 * TC_BM - 1) / (2 * TC_BM);
    auto slices_for = [&](long long ntiles, int* slice_tiles, long long* nslices, int* grid) {
        // ~16 rounds of items per pair for balance, and the slices in flight (npairs / nqb of them, each read by the
        // nqb pairs that run its query blocks) must stay L2-resident: at most ~64 MB of rows in flight, 32 pair-tiles
        // per slice
        long long ts = ntiles * pl->nqb / ((long long)npairs_max * 16);
        const long long tile_bytes = 2LL * TC_BM * d * (long long)esz;
        long long l2_cap = (64LL << 20) * pl->nqb / ((long long)npairs_max * tile_bytes);
        if (l2_cap < 2) l2_cap = 2;
        if (ts > l2_cap) ts = l2_cap;
        if (ts > 32) ts = 32;
        if (g_tc2_slice_tiles > 0) ts = g_tc2_slice_tiles;
        if (ts < 1) ts = 1;
        *slice_tiles = (int)ts;
        *nslices = (ntiles + ts - 1) / ts;
        long long items = *nslices * pl->nqb;
        *grid = 2 * (int)(items < npairs_max ? items : npairs_max);
    };
    slices_for(pl->ntiles, &pl->slice_tiles, &pl->nslices, &pl->grid);
    // pre-pass sample: every `stride`-th pair-tile, at least 256 pair-tiles = 65536 rows (or all of them)
    long long want = pl->ntiles / 128;
    if (want < tc_sample_rows(nq) / (2 * TC_BM)) want = tc_sample_rows(nq) / (2 * TC_BM);
    if (want > pl->ntiles) want = pl->ntiles;
    pl->pre_stride = pl->ntiles / want;
    pl->pre_tiles = (pl->ntiles + pl->pre_stride - 1) / pl->pre_stride;
    slices_for(pl->pre_tiles, &pl->pre_slice_tiles, &pl->pre_nslices, &pl->pre_grid);
    pl->groups = (int)(pl->pre_tiles * 8);
    int g2 = 1;
    while (g2 < pl->groups) g2 <<= 1;
    pl->gpow2 = g2;
    pl->kp = kp;
    // candidate capacity per (CTA, query): the threshold admits about 1.1 * kp * n / sampled_rows rows per query;
    // a CTA sees, for one query block, at most ceil(items / (pairs * nqb)) + 1 slices of slice_tiles * 128 rows
    const long long items = pl->nslices * pl->nqb;
    const long long pairs = pl->grid / 2;
    const long long per_block = (items + pairs * pl->nqb - 1) / (pairs * pl->nqb) + 1;
    const double rows_seen = (double)per_block * pl->slice_tiles * TC_BM;
    const double expect = 1.1 * kp / ((double)pl->groups * 32.0) * rows_seen;
    int cap = 32;
    while (cap < 4.0 * expect + 24.0 && cap < 1024) cap <<= 1;
    pl->cap = cap;
    pl->cap_total = (int)(3.0 * 1.1 * kp / ((double)pl->groups * 32.0) * (double)n) + 256;  // keys per query the gather should hold
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
    pl->off_qbf16 = take((size_t)pl->nqp * d * 2);
    pl->off_end = off;
    return cudaSuccess;
}

size_t tc2_workspace_bytes(const Tc2Plan& pl) { return pl.off_end; }

template <typename T, int MODE>
static cudaError_t launch_tc2_mode(const CUtensorMap& tdb, const CUtensorMap& tq, const Tc2Params& p, int grid, size_t smem,
                                   cudaStream_t st) {
    auto kern = tc2_scan_kernel<T, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, TC2_THREADS, smem, st>>>(tdb, tq, p);  // __cluster_dims__(2,1,1): grid is even
    g_kernel_launches.fetch_add(1);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_tc2(bool is_bf16, const CUtensorMap& tdb, const CUtensorMap& tq, const Tc2Params& p, int grid, size_t smem,
                              cudaStream_t st) {
    return is_bf16 ? launch_tc2_mode<__nv_bfloat16, MODE>(tdb, tq, p, grid, smem, st)
                   : launch_tc2_mode<float, MODE>(tdb, tq, p, grid, smem, st);
}

static cudaError_t tc2_prepare(const TcArgs& a, const Tc2Plan& pl, unsigned char* ws, CUtensorMap* tdb, CUtensorMap* tq,
                               Tc2Params* p, cudaStream_t st) {
    cudaError_t e = tc_make_tmap(tdb, a.xb, a.n, a.d, !a.is_bf16, TC_BM);
    if (e != cudaSuccess) return e;
    const void* qsrc = a.xq;
    if (a.is_bf16) {
        void* qb = ws + pl.off_qbf16;
        if ((e = tc_queries_to_bf16(a.xq, qb, (long long)a.nq * a.d, st)) != cudaSuccess) return e;
        qsrc = qb;
    }
    if ((e = tc_make_tmap(tq, qsrc, a.nq, a.d, !a.is_bf16, pl.half)) != cudaSuccess) return e;
    *p = Tc2Params{};
    p->n = a.n;
    p->d = a.d;
    p->nq = a.nq;
    p->npad = pl.npad;
    p->half = pl.half;
    p->nqb = pl.nqb;
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

// Scan `a.nq` queries with the CTA-pair kernel: one sorted kp-list per query into a.lists ([nq][kp], the format
// finalize_kernel takes with L = 1); overflow[q] = 1 where the result must not be trusted.
cudaError_t tc2_scan(const TcArgs& a, const Tc2Plan& pl, unsigned char* ws, cudaStream_t st) {
    CUtensorMap tdb, tq;
    Tc2Params p;
    cudaError_t e = tc2_prepare(a, pl, ws, &tdb, &tq, &p, st);
    if (e != cudaSuccess) return e;
    // 1. threshold pre-pass over the sampled pair-tiles
    p.ntiles = pl.pre_tiles;
    p.tile_stride = pl.pre_stride;
    p.slice_tiles = pl.pre_slice_tiles;
    p.nslices = pl.pre_nslices;
    p.block_major = 1;
    if ((e = launch_tc2<MODE_MAX>(a.is_bf16, tdb, tq, p, pl.pre_grid, pl.smem, st)) != cudaSuccess) return e;
    if ((e = tc_launch_tau0(p.gmax, pl.groups, pl.gpow2, pl.nqp, a.nq, pl.kp, reinterpret_cast<float*>(ws + pl.off_tau0), st)) !=
        cudaSuccess)
        return e;
    // the per-(CTA, query) counts, overflow flags and spill counters start at zero (adjacent in the workspace)
    if ((e = cudaMemsetAsync(ws + pl.off_counts, 0, pl.off_cand - pl.off_counts, st)) != cudaSuccess) return e;
    // 2. selection pass over every item
    p.ntiles = pl.ntiles;
    p.tile_stride = 1;
    p.slice_tiles = pl.slice_tiles;
    p.nslices = pl.nslices;
    p.block_major = 0;
    if ((e = launch_tc2<MODE_SELECT>(a.is_bf16, tdb, tq, p, pl.grid, pl.smem, st)) != cudaSuccess) return e;
    // 3. per query: gather + sort -> top-kp list
    if ((e = tc_launch_gather(p.cand, p.counts, pl.grid, pl.nqp, pl.cap, pl.kp, pl.cap_total, a.nq,
                              reinterpret_cast<u64*>(a.lists), p.overflow, p.spill_cnt, p.spill, a.fin, st)) != cudaSuccess)
        return e;
    if (a.overflow_out)
        e = cudaMemcpyAsync(a.overflow_out, ws + pl.off_overflow, (size_t)a.nq * 4, cudaMemcpyDeviceToDevice, st);
    return e;
}

// tests: raw scores of every row against the queries, [n][nqp] fp32
cudaError_t tc2_dump_scores(const TcArgs& a, const Tc2Plan& pl, unsigned char* ws, float* out, cudaStream_t st) {
    CUtensorMap tdb, tq;
    Tc2Params p;
    cudaError_t e = tc2_prepare(a, pl, ws, &tdb, &tq, &p, st);
    if (e != cudaSuccess) return e;
    p.ntiles = pl.ntiles;
    p.tile_stride = 1;
    p.slice_tiles = pl.slice_tiles;
    p.nslices = pl.nslices;
    p.dump = out;
    return launch_tc2<MODE_DUMP>(a.is_bf16, tdb, tq, p, pl.grid, pl.smem, st);
}

}  // namespace evs
