// evs_internal.h -- host-side declarations shared by evs_kernels.cu and evs_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/evs.h"

namespace evs {

extern std::atomic<long long> g_kernel_launches;

struct ScanTuning {
    int scan_variant = 0;  // 0 auto, 1 direct loads, 2 bulk-async ring
    int tile_rows = 0;     // ring: rows per stage (0 = ~32 KiB)
    int stages = 0;        // ring: depth (0 = 4)
    int ctas_per_sm = 0;   // direct: CTAs per SM (0 = 2 fp32 / 4 bf16)
    int tc_min_nq = 2;     // query batches of at least this many use the tensor-core scan (0 = never); measured
                           // (scripts/small_nq_sweep.py): from 2 queries on it beats the CUDA-core GEMV for fp32 and bf16 rows
    int tc_pair_min_nq = 129;  // ... and of at least this many the CTA-pair kernel (evs_tc2.cu); 0 = never
};

struct ScanPlan {
    int variant;  // 0 generic-d, 1 direct, 2 ring
    int nv;       // 16-byte vectors per lane per row (0 for generic)
    int grid, threads;
    size_t smem_bytes;
    int tile_rows, stages;
};

struct ScanArgs {
    const void* xb;
    int is_bf16;
    long long n;
    int d;
    const float* xq;
    int q0;
    int nq_pass;  // 1..4 queries handled by this launch
    void* lists;  // u64 [nq_chunk][grid][kp]
    int kp;
};

// Peer-store exchange of shard partials (row sharding): every rank owns a symmetric buffer
//   [2 parities][world shards][slot_bytes]  gather slots (float64 scores [nq*k] then int64 ids [nq*k])
//   [2 parities][world shards] uint64       arrival flags holding the search sequence number
// and sees all ranks' buffers through peer-mapped pointers.
struct Exchange {
    unsigned char* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int rank = 0, world = 0, parity = 0;
    unsigned long long seq = 0;
    size_t slot_bytes = 0;
    unsigned* done = nullptr;  // local counter of finished finalize CTAs
    long long nq_total = 0;    // queries of the whole search (a search may be finalised in several launches)
    long long q_off = 0;       // first query of this launch
};

struct FinalizeArgs {
    const void* lists;
    int L, kp;
    const void* xb;
    int xb_is_bf16;
    const float* xq;
    long long nq;
    int d, k;
    long long id_base;
    float* D;
    int64_t* I;
    double* P_scores;
    int64_t* P_ids;
    float* margins;
    Exchange x;
};

// tensor-core scan (evs_tc.cu)
struct TcPlan {
    int npad, nblocks, nqp, nk, stages, grid, pre_grid, groups, gpow2, cap, cap_total, kp;
    int heap;  // 1: MODE_HEAP (small batch, lists [nq][grid][64] come straight out of the scan, L = grid for finalize)
    size_t smem;
    long long ntiles, pre_tiles, pre_stride;
    size_t off_gmax, off_tau0, off_counts, off_overflow, off_spill_cnt, off_cand, off_spill, off_qbf16, off_end;
};
struct TcArgs {
    const void* xb;   // database rows (fp32 or bf16)
    int is_bf16;
    long long n;
    int d;
    const float* xq;  // the queries of this launch set, fp32 [nq][d]
    int nq;
    void* lists;      // out: u64 [nq][kp]
    int* overflow_out;  // out (optional): int [nq]
};
// CTA-pair tensor-core scan for large batches (evs_tc2.cu)
struct Tc2Plan {
    int npad, half, nqb, nqp, nk, stages, grid, pre_grid, groups, gpow2, cap, cap_total, kp;
    int slice_tiles, pre_slice_tiles;
    size_t smem;
    long long ntiles, nslices, pre_tiles, pre_stride, pre_nslices;
    size_t off_gmax, off_tau0, off_counts, off_overflow, off_spill_cnt, off_cand, off_spill, off_qbf16, off_end;
};
extern int g_tc2_slice_tiles;
int tc2_max_half(int d, int is_bf16);
cudaError_t tc2_plan(long long n, int d, int is_bf16, int nq, int kp, int sm_count, Tc2Plan* pl);
size_t tc2_workspace_bytes(const Tc2Plan& pl);
cudaError_t tc2_scan(const TcArgs& a, const Tc2Plan& pl, unsigned char* ws, cudaStream_t st);
cudaError_t tc2_dump_scores(const TcArgs& a, const Tc2Plan& pl, unsigned char* ws, float* out, cudaStream_t st);
constexpr int TC_SPILL_CAP = 1024;  // per-query spill list behind the (CTA, query) candidate buffers
extern int g_tc_max_stages;
extern int g_tc_sample_rows;
extern int g_tc_heap_max_nq;
extern int g_tc_heap_pure_max_nq;
int tc_sample_rows(int nq);
int tc_max_queries(int d, int is_bf16);
cudaError_t tc_plan(long long n, int d, int is_bf16, int nq, int kp, int sm_count, TcPlan* pl);
size_t tc_workspace_bytes(const TcPlan& pl);
cudaError_t tc_scan_block(const TcArgs& a, const TcPlan& pl, unsigned char* ws, cudaStream_t st);
cudaError_t tc_dump_scores(const TcArgs& a, const TcPlan& pl, unsigned char* ws, float* out, cudaStream_t st);

// Launch with programmatic stream serialisation: the kernel may become resident while the preceding kernel of the stream is
// still draining (its CTAs call griddepcontrol.launch_dependents early); the kernel itself must execute griddepcontrol.wait
// before it reads anything the predecessor wrote.  Hides launch latency and prologues on the multi-kernel search paths.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

cudaError_t plan_scan(long long n, int d, int is_bf16, int kp, int nq_pass, int sm_count, const ScanTuning& tune,
                      ScanPlan* plan);
int max_queries_per_pass(int d, int is_bf16);
cudaError_t launch_scan(const ScanArgs& a, ScanPlan* plan, cudaStream_t st);
cudaError_t launch_finalize(const FinalizeArgs& a, cudaStream_t st);
cudaError_t launch_publish_partials(const Exchange& x, long long nq, int k, const double* scores, const long long* ids,
                                    cudaStream_t st);
cudaError_t launch_merge_exchange(const Exchange& x, long long nq, int k, float* D, long long* I, int* timed_out,
                                  cudaStream_t st);
cudaError_t launch_merge_partials(int nparts, long long nq, int k, const double* scores, const long long* ids,
                                  long long part_stride, float* D,
                                  long long* I, cudaStream_t st);
cudaError_t launch_l2_normalize(void* x, long long n, int d, int dtype, int sm_count, cudaStream_t st);
cudaError_t launch_f32_to_bf16(const float* src, void* dst, long long count, int sm_count, cudaStream_t st);
cudaError_t launch_to_f32(const void* src, int dtype, float* dst, long long count, int sm_count, cudaStream_t st);
cudaError_t launch_gather_rows(const float* src, const long long* ids_dev, float* dst, long long n, int d, int sm_count,
                               cudaStream_t st);
cudaError_t launch_synth_fill(float* out, long long n, int d, unsigned long long seed, long long row_base, int sm_count,
                              cudaStream_t st);

}  // namespace evs
