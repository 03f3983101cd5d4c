// evs_api.cu -- the C ABI of libevs.so (include/evs.h): index handles, add / search / persistence.
// No CPU fallback anywhere: without a CUDA device every compute entry point returns EVS_ENODEV.
#include <errno.h>
#include <float.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "evs_internal.h"

using namespace evs;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char t_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            int c__ = (e__ == cudaErrorMemoryAllocation) ? EVS_ENOMEM                                     \
                      : (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver) ? EVS_ENODEV     \
                                                                                         : EVS_ECUDA;    \
            return fail(c__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                                 \
    } while (0)

static ScanTuning g_tune;
static std::mutex g_tune_mu;
static int g_profile_scans = 0;
static std::atomic<long long> g_tc_fallbacks{0};
static std::atomic<long long> g_exact_reruns{0};   // queries re-run with the fp32 GEMV scan because the tf32 scan's margin was not certifying
static int g_tf32_guard_eps_e6 = 150;             // option "tf32_guard_eps_e6": margin (x 1e-6) below which a tf32-scanned result is re-run; 0 = off  // queries re-run through the GEMV scan after a tensor-core overflow  // record CUDA events around every search's scan launches

// ---------------------------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------------------------
static const int kQueryChunk = 256;     // GEMV path: queries finalised per launch (bounds the list workspace)
static const int kTcQueryChunk = 4096;  // tensor-core path: queries per launch set

struct evs_index {
    int d = 0, device = 0, storage = EVS_STORE_F32;
    int64_t ntotal = 0, capacity = 0, id_base = 0;
    float* xb32 = nullptr;            // fp32 rows (the master copy; what index.faiss holds)
    void* xb16 = nullptr;             // derived bf16 rows (EVS_STORE_BF16_F32)
    int sm_count = 0;
    cudaStream_t stream = nullptr;    // the handle's own stream
    cudaEvent_t ws_free = nullptr;    // recorded after each search: the workspace may be reused after it
    std::mutex mu;                    // serialises add/search on this handle (Flask threads)
    // workspace (device)
    float* q_dev = nullptr;  size_t q_cap = 0;         // staged queries [chunk][d]
    void* lists = nullptr;   size_t lists_cap = 0;     // candidate lists
    float* D_dev = nullptr;  int64_t* I_dev = nullptr; size_t out_cap = 0;  // [chunk][k]
    float* margins_dev = nullptr; size_t margins_cap = 0;   // [nq of last search]
    unsigned char* tc_ws = nullptr; size_t tc_ws_cap = 0;   // tensor-core scan workspace
    int* tc_overflow = nullptr; size_t tc_overflow_cap = 0; // [nq] overflow flags of the tensor-core scan
    int* tc_overflow_pin = nullptr; size_t tc_overflow_pin_cap = 0;
    int64_t last_nq = 0;
    // optional per-search timing of the scan stage (option "profile_scans")
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    // pinned host staging
    float* q_pin = nullptr;  size_t q_pin_cap = 0;
    float* D_pin = nullptr;  int64_t* I_pin = nullptr; size_t out_pin_cap = 0;
    float* m_pin = nullptr;  size_t m_pin_cap = 0;     // margins of the last host search (tf32 guard)
};

static int use_device(int device) {
    CU(cudaSetDevice(device));
    return EVS_OK;
}

template <typename T>
static int ensure_dev(T** ptr, size_t* cap, size_t need_elems) {
    if (*cap >= need_elems && *ptr) return EVS_OK;
    if (*ptr) CU(cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    CU(cudaMalloc(reinterpret_cast<void**>(ptr), need_elems * sizeof(T)));
    *cap = need_elems;
    return EVS_OK;
}

static int pick_kp(int64_t k) { return k <= 48 ? 64 : 128; }

namespace evs {
int max_queries_per_pass(int d, int is_bf16) {
    // keep the query registers of the vectorised kernels at <= 64 per lane
    int per_lane = d / 32;  // query values held per lane per query
    if (per_lane <= 0) return 1;
    int m = 64 / per_lane;
    if (m > 4) m = 4;
    if (m < 1) m = 1;
    (void)is_bf16;
    return m;
}
}  // namespace evs

// ---------------------------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------------------------
extern "C" int evs_version(void) { return EVS_VERSION; }
extern "C" const char* evs_last_error(void) { return t_err; }

extern "C" int evs_device_count(int* count) {
    if (!count) return fail(EVS_EINVAL, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return EVS_OK;
}

extern "C" int64_t evs_kernel_launches(void) { return (int64_t)g_kernel_launches.load(); }

extern "C" int evs_set_option(const char* name, int64_t value) {
    if (!name) return fail(EVS_EINVAL, "name is NULL");
    std::lock_guard<std::mutex> lk(g_tune_mu);
    if (!strcmp(name, "scan_variant")) {
        if (value < 0 || value > 2) return fail(EVS_EINVAL, "scan_variant must be 0, 1 or 2");
        g_tune.scan_variant = (int)value;
    } else if (!strcmp(name, "tile_rows")) {
        if (value < 0 || value > 1024) return fail(EVS_EINVAL, "tile_rows out of range");
        g_tune.tile_rows = (int)value;
    } else if (!strcmp(name, "stages")) {
        if (value < 0 || value > 16) return fail(EVS_EINVAL, "stages out of range");
        g_tune.stages = (int)value;
    } else if (!strcmp(name, "ctas_per_sm")) {
        if (value < 0 || value > 8) return fail(EVS_EINVAL, "ctas_per_sm out of range");
        g_tune.ctas_per_sm = (int)value;
    } else if (!strcmp(name, "tc_min_nq")) {
        if (value < 0) return fail(EVS_EINVAL, "tc_min_nq must be >= 0");
        g_tune.tc_min_nq = (int)value;
    } else if (!strcmp(name, "tc_pair_min_nq")) {
        if (value < 0) return fail(EVS_EINVAL, "tc_pair_min_nq must be >= 0");
        g_tune.tc_pair_min_nq = (int)value;
    } else if (!strcmp(name, "tc2_slice_tiles")) {
        if (value < 0 || value > 4096) return fail(EVS_EINVAL, "tc2_slice_tiles must be in [0, 4096]");
        g_tc2_slice_tiles = (int)value;
    } else if (!strcmp(name, "tc_heap_max_nq")) {
        if (value < 0 || value > 128) return fail(EVS_EINVAL, "tc_heap_max_nq must be in [0, 128]");
        g_tc_heap_max_nq = (int)value;
    } else if (!strcmp(name, "tc_heap_pure_max_nq")) {
        if (value < 0 || value > 128) return fail(EVS_EINVAL, "tc_heap_pure_max_nq must be in [0, 128]");
        g_tc_heap_pure_max_nq = (int)value;
    } else if (!strcmp(name, "tf32_guard_eps_e6")) {
        if (value < 0 || value > 100000) return fail(EVS_EINVAL, "tf32_guard_eps_e6 must be in [0, 100000]");
        g_tf32_guard_eps_e6 = (int)value;
    } else if (!strcmp(name, "tc_sample_rows")) {
        if (value != 0 && (value < 1024 || value > (1 << 24))) return fail(EVS_EINVAL, "tc_sample_rows must be 0 (auto) or in [1024, 2^24]");
        g_tc_sample_rows = (int)value;
    } else if (!strcmp(name, "tc_stages")) {
        if (value < 2 || value > 14) return fail(EVS_EINVAL, "tc_stages must be in [2, 14]");
        g_tc_max_stages = (int)value;
    } else if (!strcmp(name, "profile_scans")) {
        g_profile_scans = value ? 1 : 0;
    } else {
        return fail(EVS_EINVAL, "unknown option '%s'", name);
    }
    return EVS_OK;
}

extern "C" int evs_get_option(const char* name, int64_t* value) {
    if (!name || !value) return fail(EVS_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(g_tune_mu);
    if (!strcmp(name, "scan_variant")) *value = g_tune.scan_variant;
    else if (!strcmp(name, "tile_rows")) *value = g_tune.tile_rows;
    else if (!strcmp(name, "stages")) *value = g_tune.stages;
    else if (!strcmp(name, "ctas_per_sm")) *value = g_tune.ctas_per_sm;
    else if (!strcmp(name, "tc_min_nq")) *value = g_tune.tc_min_nq;
    else if (!strcmp(name, "tc_pair_min_nq")) *value = g_tune.tc_pair_min_nq;
    else if (!strcmp(name, "tc2_slice_tiles")) *value = g_tc2_slice_tiles;
    else if (!strcmp(name, "tc_heap_max_nq")) *value = g_tc_heap_max_nq;
    else if (!strcmp(name, "tc_heap_pure_max_nq")) *value = g_tc_heap_pure_max_nq;
    else if (!strcmp(name, "tc_sample_rows")) *value = g_tc_sample_rows;
    else if (!strcmp(name, "tc_stages")) *value = g_tc_max_stages;
    else if (!strcmp(name, "profile_scans")) *value = g_profile_scans;
    else if (!strcmp(name, "tc_fallbacks")) *value = g_tc_fallbacks.load();  // read-only counter
    else if (!strcmp(name, "exact_reruns")) *value = g_exact_reruns.load();  // read-only counter
    else if (!strcmp(name, "tf32_guard_eps_e6")) *value = g_tf32_guard_eps_e6;
    else return fail(EVS_EINVAL, "unknown option '%s'", name);
    return EVS_OK;
}

// ---------------------------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------------------------
extern "C" int evs_index_create(int d, int device, int storage, evs_index** out) {
    if (!out) return fail(EVS_EINVAL, "out is NULL");
    *out = nullptr;
    if (d <= 0) return fail(EVS_EINVAL, "d must be > 0 (got %d)", d);
    if (storage != EVS_STORE_F32 && storage != EVS_STORE_BF16_F32) return fail(EVS_EINVAL, "unknown storage %d", storage);
    int ndev = 0;
    evs_device_count(&ndev);
    if (ndev <= 0) return fail(EVS_ENODEV, "no CUDA device: libevs has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(EVS_EINVAL, "device %d out of range (have %d)", device, ndev);
    int rc = use_device(device);
    if (rc) return rc;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(EVS_ENODEV, "device %d is sm_%d%d; libevs is built for sm_100a only", device, prop.major, prop.minor);
    evs_index* idx = new (std::nothrow) evs_index();
    if (!idx) return fail(EVS_ENOMEM, "out of host memory");
    idx->d = d;
    idx->device = device;
    idx->storage = storage;
    idx->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&idx->ws_free, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete idx;
        return fail(EVS_ECUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
    }
    *out = idx;
    return EVS_OK;
}

extern "C" int evs_index_free(evs_index* idx) {
    if (!idx) return EVS_OK;
    cudaSetDevice(idx->device);
    if (idx->stream) cudaStreamSynchronize(idx->stream);
    cudaFree(idx->xb32);
    cudaFree(idx->xb16);
    cudaFree(idx->q_dev);
    cudaFree(idx->lists);
    cudaFree(idx->I_dev);  // D_dev points into the same allocation
    cudaFree(idx->margins_dev);
    cudaFree(idx->tc_ws);
    cudaFree(idx->tc_overflow);
    cudaFreeHost(idx->tc_overflow_pin);
    cudaFreeHost(idx->q_pin);
    cudaFreeHost(idx->I_pin);  // D_pin points into the same allocation
    cudaFreeHost(idx->m_pin);
    for (auto& pe : idx->prof_events) {
        cudaEventDestroy(pe.first);
        cudaEventDestroy(pe.second);
    }
    if (idx->ws_free) cudaEventDestroy(idx->ws_free);
    if (idx->stream) cudaStreamDestroy(idx->stream);
    delete idx;
    return EVS_OK;
}

#define GETTER(name, type, expr)                                            \
    extern "C" int name(const evs_index* idx, type* out) {                  \
        if (!idx || !out) return fail(EVS_EINVAL, #name ": NULL argument"); \
        *out = (expr);                                                      \
        return EVS_OK;                                                      \
    }
GETTER(evs_index_d, int, idx->d)
GETTER(evs_index_ntotal, int64_t, idx->ntotal)
GETTER(evs_index_device, int, idx->device)
GETTER(evs_index_storage, int, idx->storage)
GETTER(evs_index_id_base, int64_t, idx->id_base)

extern "C" int evs_index_set_id_base(evs_index* idx, int64_t base) {
    if (!idx) return fail(EVS_EINVAL, "idx is NULL");
    if (base < 0) return fail(EVS_EINVAL, "id base must be >= 0");
    idx->id_base = base;
    return EVS_OK;
}

// grow the row storage to hold at least `rows` rows (device-to-device copy of what is there)
static int grow_locked(evs_index* idx, int64_t rows) {
    if (rows <= idx->capacity) return EVS_OK;
    if (rows >= (int64_t)0xFFFFFFFFll) return fail(EVS_ELIMIT, "a shard holds at most 2^32-2 rows");
    const size_t d = (size_t)idx->d;
    float* n32 = nullptr;
    void* n16 = nullptr;
    CU(cudaMalloc(reinterpret_cast<void**>(&n32), (size_t)rows * d * sizeof(float)));
    if (idx->storage == EVS_STORE_BF16_F32) {
        cudaError_t e = cudaMalloc(&n16, (size_t)rows * d * 2);
        if (e != cudaSuccess) {
            cudaFree(n32);
            return fail(EVS_ENOMEM, "cudaMalloc of the bf16 copy failed: %s", cudaGetErrorString(e));
        }
    }
    if (idx->ntotal > 0) {
        CU(cudaMemcpyAsync(n32, idx->xb32, (size_t)idx->ntotal * d * sizeof(float), cudaMemcpyDeviceToDevice, idx->stream));
        if (n16) CU(cudaMemcpyAsync(n16, idx->xb16, (size_t)idx->ntotal * d * 2, cudaMemcpyDeviceToDevice, idx->stream));
        CU(cudaStreamSynchronize(idx->stream));
    }
    cudaFree(idx->xb32);
    cudaFree(idx->xb16);
    idx->xb32 = n32;
    idx->xb16 = n16;
    idx->capacity = rows;
    return EVS_OK;
}

static int grow_for_add_locked(evs_index* idx, int64_t n) {
    int64_t need = idx->ntotal + n;
    if (need <= idx->capacity) return EVS_OK;
    int64_t cap = idx->capacity + idx->capacity / 2;  // 1.5x amortised growth
    if (cap < need) cap = need;
    return grow_locked(idx, cap);
}

// after new fp32 rows [ntotal, ntotal+n) are in place (enqueued on idx->stream): derive bf16, publish
static int finish_add_locked(evs_index* idx, int64_t n) {
    const size_t d = (size_t)idx->d;
    if (idx->storage == EVS_STORE_BF16_F32) {
        CU(launch_f32_to_bf16(idx->xb32 + (size_t)idx->ntotal * d,
                              reinterpret_cast<unsigned char*>(idx->xb16) + (size_t)idx->ntotal * d * 2, (long long)n * idx->d,
                              idx->sm_count, idx->stream));
    }
    CU(cudaStreamSynchronize(idx->stream));
    idx->ntotal += n;
    return EVS_OK;
}

extern "C" int evs_index_reserve(evs_index* idx, int64_t nrows) {
    if (!idx) return fail(EVS_EINVAL, "idx is NULL");
    if (nrows < 0) return fail(EVS_EINVAL, "nrows must be >= 0");
    std::lock_guard<std::mutex> lk(idx->mu);
    int rc = use_device(idx->device);
    if (rc) return rc;
    return grow_locked(idx, nrows);
}

extern "C" int evs_index_add(evs_index* idx, int64_t n, const float* x_host) {
    if (!idx) return fail(EVS_EINVAL, "idx is NULL");
    if (n < 0) return fail(EVS_EINVAL, "n must be >= 0");
    if (n == 0) return EVS_OK;
    if (!x_host) return fail(EVS_EINVAL, "x is NULL");
    std::lock_guard<std::mutex> lk(idx->mu);
    int rc = use_device(idx->device);
    if (rc) return rc;
    if ((rc = grow_for_add_locked(idx, n))) return rc;
    const size_t d = (size_t)idx->d;
    CU(cudaMemcpyAsync(idx->xb32 + (size_t)idx->ntotal * d, x_host, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice,
                       idx->stream));
    return finish_add_locked(idx, n);
}

extern "C" int evs_index_add_dev(evs_index* idx, int64_t n, const void* x_dev, int dtype, void* stream) {
    if (!idx) return fail(EVS_EINVAL, "idx is NULL");
    if (n < 0) return fail(EVS_EINVAL, "n must be >= 0");
    if (n == 0) return EVS_OK;
    if (!x_dev) return fail(EVS_EINVAL, "x is NULL");
    if (dtype != EVS_F32 && dtype != EVS_F16 && dtype != EVS_BF16) return fail(EVS_EINVAL, "bad dtype %d", dtype);
    std::lock_guard<std::mutex> lk(idx->mu);
    int rc = use_device(idx->device);
    if (rc) return rc;
    if ((rc = grow_for_add_locked(idx, n))) return rc;
    // the producer of x_dev ran on `stream`: order our stream after it
    if ((cudaStream_t)stream != idx->stream) {
        cudaEvent_t ev;
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CU(cudaEventRecord(ev, (cudaStream_t)stream));
        CU(cudaStreamWaitEvent(idx->stream, ev, 0));
        CU(cudaEventDestroy(ev));
    }
    const size_t d = (size_t)idx->d;
    float* dst = idx->xb32 + (size_t)idx->ntotal * d;
    if (dtype == EVS_F32)
        CU(cudaMemcpyAsync(dst, x_dev, (size_t)n * d * sizeof(float), cudaMemcpyDeviceToDevice, idx->stream));
    else
        CU(launch_to_f32(x_dev, dtype, dst, (long long)n * idx->d, idx->sm_count, idx->stream));
    return finish_add_locked(idx, n);
}

extern "C" int evs_index_add_synth(evs_index* idx, int64_t n, uint64_t seed, int normalize) {
    if (!idx) return fail(EVS_EINVAL, "idx is NULL");
    if (n < 0) return fail(EVS_EINVAL, "n must be >= 0");
    if (n == 0) return EVS_OK;
    std::lock_guard<std::mutex> lk(idx->mu);
    int rc = use_device(idx->device);
    if (rc) return rc;
    if ((rc = grow_for_add_locked(idx, n))) return rc;
    float* dst = idx->xb32 + (size_t)idx->ntotal * idx->d;
    CU(launch_synth_fill(dst, n, idx->d, seed, idx->id_base + idx->ntotal, idx->sm_count, idx->stream));
    if (normalize) CU(launch_l2_normalize(dst, n, idx->d, EVS_F32, idx->sm_count, idx->stream));
    return finish_add_locked(idx, n);
}

extern "C" int evs_index_add_rows_from(evs_index* dst, const evs_index* src, int64_t n, const int64_t* rows_host) {
    if (!dst || !src) return fail(EVS_EINVAL, "NULL index");
    if (dst == src) return fail(EVS_EINVAL, "source and destination must be different indexes");
    if (n < 0) return fail(EVS_EINVAL, "n must be >= 0");
    if (n == 0) return EVS_OK;
    if (!rows_host) return fail(EVS_EINVAL, "rows is NULL");
    if (dst->d != src->d) return fail(EVS_EINVAL, "dimension mismatch: %d vs %d", dst->d, src->d);
    if (dst->device != src->device) return fail(EVS_EINVAL, "indexes live on different devices (%d, %d)", dst->device, src->device);
    for (int64_t i = 0; i < n; i++)
        if (rows_host[i] < 0 || rows_host[i] >= src->ntotal)
            return fail(EVS_EINVAL, "row %lld out of range [0, %lld)", (long long)rows_host[i], (long long)src->ntotal);
    // both handles: the source must not be grown (its row storage reallocated) while its rows are read
    std::unique_lock<std::mutex> lk(dst->mu, std::defer_lock);
    std::unique_lock<std::mutex> lks(const_cast<evs_index*>(src)->mu, std::defer_lock);
    std::lock(lk, lks);
    for (int64_t i = 0; i < n; i++)  // re-checked under the lock
        if (rows_host[i] >= src->ntotal) return fail(EVS_EINVAL, "row %lld out of range", (long long)rows_host[i]);
    int rc = use_device(dst->device);
    if (rc) return rc;
    if ((rc = grow_for_add_locked(dst, n))) return rc;
    long long* ids_dev = nullptr;
    CU(cudaMalloc(reinterpret_cast<void**>(&ids_dev), (size_t)n * sizeof(long long)));
    cudaError_t e = cudaMemcpyAsync(ids_dev, rows_host, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, dst->stream);
    if (e == cudaSuccess)
        e = launch_gather_rows(src->xb32, ids_dev, dst->xb32 + (size_t)dst->ntotal * dst->d, n, dst->d, dst->sm_count, dst->stream);
    if (e == cudaSuccess) rc = finish_add_locked(dst, n);  // synchronises the stream
    else cudaStreamSynchronize(dst->stream);
    cudaFree(ids_dev);
    if (e != cudaSuccess) return fail(EVS_ECUDA, "row gather failed: %s", cudaGetErrorString(e));
    return rc;
}

extern "C" int evs_index_get_rows(const evs_index* idx, int64_t row0, int64_t n, float* out_host) {
    if (!idx || (!out_host && n > 0)) return fail(EVS_EINVAL, "NULL argument");
    if (row0 < 0 || n < 0 || row0 + n > idx->ntotal) return fail(EVS_EINVAL, "rows [%lld,%lld) out of range", (long long)row0, (long long)(row0 + n));
    if (n == 0) return EVS_OK;
    int rc = use_device(idx->device);
    if (rc) return rc;
    CU(cudaMemcpy(out_host, idx->xb32 + (size_t)row0 * idx->d, (size_t)n * idx->d * sizeof(float), cudaMemcpyDeviceToHost));
    return EVS_OK;
}

// ---------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------
struct SearchOut {
    float* D = nullptr;        // final mode
    int64_t* I = nullptr;
    double* P_scores = nullptr;  // partial mode
    int64_t* P_ids = nullptr;
    const Exchange* x = nullptr;  // exchange mode: the finalise kernel stores the partial into every rank's slot
};

// Enqueue the search of `nq` device-resident queries on stream `st`; outputs are device pointers.
// idx->mu is held by the caller.
template <typename T>
static int ensure_pinned(T** ptr, size_t* cap, size_t need_elems);

static int search_enqueue_locked(evs_index* idx, int64_t nq, const float* q_dev, int64_t k, const SearchOut& out,
                                 cudaStream_t st, bool scan_only = false, bool allow_tc = true);

// Tensor-core path for a batch: one database pass per block of up to tc_max_queries queries.
// Queries whose candidate buffers overflowed are re-run through the GEMV path (needs one host sync).
static int search_tc_locked(evs_index* idx, int64_t nq, const float* q_dev, int64_t k, const SearchOut& out, cudaStream_t st,
                            bool scan_only, int profile, int pair_min_nq) {
    const int kp = pick_kp(k);
    const bool bf16 = idx->storage == EVS_STORE_BF16_F32;
    const void* scan_rows = bf16 ? idx->xb16 : (const void*)idx->xb32;
    // batches of pair_min_nq or more queries go through the CTA-pair kernel (N up to 256 per MMA, L2-shared slices)
    const bool can_pair = pair_min_nq > 0 && tc2_max_half(idx->d, bf16) > 0 && idx->sm_count >= 2;
    // ... and so do batches that would need more than one pass of the one-CTA kernel (fp32 rows: 64 queries per pass)
    const int one_cta_max = tc_max_queries(idx->d, bf16);
    auto use_pair = [&](int64_t cn) { return can_pair && (cn >= pair_min_nq || cn > one_cta_max); };
    const int64_t first = nq < kTcQueryChunk ? nq : kTcQueryChunk;
    size_t ws_need = 0;
    if (use_pair(first)) {
        Tc2Plan p2;
        CU(tc2_plan(idx->ntotal, idx->d, bf16, (int)first, kp, idx->sm_count, &p2));
        ws_need = tc2_workspace_bytes(p2);
    }
    const int64_t last = nq % kTcQueryChunk;  // a shorter final chunk may take the other kernel
    size_t lists_need = (size_t)first * kp;
    if (!use_pair(first) || (last && !use_pair(last))) {
        TcPlan pl;
        const int64_t cn1 = use_pair(first) ? last : first;
        CU(tc_plan(idx->ntotal, idx->d, bf16, (int)cn1, kp, idx->sm_count, &pl));
        if (tc_workspace_bytes(pl) > ws_need) ws_need = tc_workspace_bytes(pl);
        if (pl.heap && (size_t)cn1 * pl.grid * kp > lists_need) lists_need = (size_t)cn1 * pl.grid * kp;  // one list per CTA
    }
    int rc = ensure_dev(&idx->tc_ws, &idx->tc_ws_cap, ws_need);
    if (rc) return rc;
    if ((rc = ensure_dev(&idx->tc_overflow, &idx->tc_overflow_cap, (size_t)nq))) return rc;
    if ((rc = ensure_dev(reinterpret_cast<unsigned long long**>(&idx->lists), &idx->lists_cap, lists_need))) return rc;
    if ((rc = ensure_dev(&idx->margins_dev, &idx->margins_cap, (size_t)nq))) return rc;
    idx->last_nq = nq;
    bool may_overflow = false;  // MODE_HEAP chunks cannot: then nothing below needs the host
    for (int64_t c0 = 0; c0 < nq; c0 += kTcQueryChunk) {
        const int64_t cn = (nq - c0) < kTcQueryChunk ? (nq - c0) : kTcQueryChunk;
        std::pair<cudaEvent_t, cudaEvent_t>* pe = nullptr;
        if (profile && idx->prof_used < 65536) {
            if (idx->prof_used == idx->prof_events.size()) {
                cudaEvent_t a = nullptr, b = nullptr;
                CU(cudaEventCreate(&a));
                CU(cudaEventCreate(&b));
                idx->prof_events.emplace_back(a, b);
            }
            pe = &idx->prof_events[idx->prof_used++];
            CU(cudaEventRecord(pe->first, st));
        }
        int lists_per_query = 1;
        {
            TcArgs a;
            a.xb = scan_rows;
            a.is_bf16 = bf16;
            a.n = idx->ntotal;
            a.d = idx->d;
            a.xq = q_dev + (size_t)c0 * idx->d;
            a.nq = (int)cn;
            a.lists = idx->lists;
            a.overflow_out = idx->tc_overflow + c0;
            if (use_pair(cn)) {
                may_overflow = true;
                Tc2Plan plb;
                CU(tc2_plan(idx->ntotal, idx->d, bf16, (int)cn, kp, idx->sm_count, &plb));
                if (tc2_workspace_bytes(plb) > idx->tc_ws_cap) return fail(EVS_ECUDA, "internal: tensor-core workspace too small");
                CU(tc2_scan(a, plb, idx->tc_ws, st));
            } else {
                TcPlan plb;  // same workspace bound: cn <= the chunk the workspace was sized for
                CU(tc_plan(idx->ntotal, idx->d, bf16, (int)cn, kp, idx->sm_count, &plb));
                if (tc_workspace_bytes(plb) > idx->tc_ws_cap) return fail(EVS_ECUDA, "internal: tensor-core workspace too small");
                if (plb.heap) {
                    lists_per_query = plb.grid;
                    if ((size_t)cn * plb.grid * kp > idx->lists_cap) return fail(EVS_ECUDA, "internal: list workspace too small");
                } else {
                    may_overflow = true;
                }
                CU(tc_scan_block(a, plb, idx->tc_ws, st));
            }
        }
        if (pe) CU(cudaEventRecord(pe->second, st));
        if (scan_only) continue;
        FinalizeArgs f;
        f.lists = idx->lists;
        f.L = lists_per_query;
        f.kp = kp;
        f.xb = idx->xb32;
        f.xb_is_bf16 = 0;
        f.xq = q_dev + (size_t)c0 * idx->d;
        f.nq = cn;
        f.d = idx->d;
        f.k = (int)k;
        f.id_base = idx->id_base;
        f.D = out.D ? out.D + (size_t)c0 * k : nullptr;
        f.I = out.I ? out.I + (size_t)c0 * k : nullptr;
        f.P_scores = out.P_scores ? out.P_scores + (size_t)c0 * k : nullptr;
        f.P_ids = out.P_ids ? out.P_ids + (size_t)c0 * k : nullptr;
        f.margins = idx->margins_dev + c0;
        if (out.x) {  // exchange mode (only offered for batches that cannot need the host-side repair below)
            f.x = *out.x;
            f.x.nq_total = nq;
            f.x.q_off = c0;
        }
        CU(launch_finalize(f, st));
    }
    if (scan_only || !may_overflow) return EVS_OK;
    if (out.x) return fail(EVS_ECUDA, "internal: exchange-mode search took a path that may overflow");
    // exactness guard: re-run overflowed queries with the GEMV scan
    if ((rc = ensure_pinned(&idx->tc_overflow_pin, &idx->tc_overflow_pin_cap, (size_t)nq))) return rc;
    CU(cudaMemcpyAsync(idx->tc_overflow_pin, idx->tc_overflow, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int64_t q = 0; q < nq; q++) {
        if (!idx->tc_overflow_pin[q]) continue;
        g_tc_fallbacks.fetch_add(1);
        SearchOut o1;
        o1.D = out.D ? out.D + (size_t)q * k : nullptr;
        o1.I = out.I ? out.I + (size_t)q * k : nullptr;
        o1.P_scores = out.P_scores ? out.P_scores + (size_t)q * k : nullptr;
        o1.P_ids = out.P_ids ? out.P_ids + (size_t)q * k : nullptr;
        float* keep = idx->margins_dev;
        idx->margins_dev = keep + q;  // the re-run writes this query's margin in place
        rc = search_enqueue_locked(idx, 1, q_dev + (size_t)q * idx->d, k, o1, st, false, false);
        idx->margins_dev = keep;
        idx->last_nq = nq;
        if (rc) return rc;
    }
    return EVS_OK;
}

static bool takes_tc_path(const evs_index* idx, int64_t nq, const ScanTuning& tune) {
    // TMA row coordinates are int32: shards beyond 2^31 rows (4 TB of bf16 at d = 512: not on this hardware) keep the GEMV scan
    return tune.tc_min_nq > 0 && nq >= tune.tc_min_nq && idx->ntotal >= 65536 && idx->ntotal < ((int64_t)1 << 31) - 1024 &&
           tc_max_queries(idx->d, idx->storage == EVS_STORE_BF16_F32) > 0;
}

// true when the tensor-core path serves this batch entirely from on-chip heaps (MODE_HEAP): no overflow case, so no
// host synchronisation and the finalise kernel may write straight into the exchange slots
static bool tc_path_is_heap(const evs_index* idx, int64_t nq, int64_t k, const ScanTuning& tune) {
    if (nq > kTcQueryChunk) return false;
    const bool bf16 = idx->storage == EVS_STORE_BF16_F32;
    if (tune.tc_pair_min_nq > 0 && (nq >= tune.tc_pair_min_nq || nq > tc_max_queries(idx->d, bf16)) && tc2_max_half(idx->d, bf16) > 0)
        return false;
    TcPlan pl;
    if (tc_plan(idx->ntotal, idx->d, bf16, (int)nq, pick_kp(k), idx->sm_count, &pl) != cudaSuccess) return false;
    return pl.heap != 0;
}

static int search_enqueue_locked(evs_index* idx, int64_t nq, const float* q_dev, int64_t k, const SearchOut& out,
                                 cudaStream_t st, bool scan_only, bool allow_tc) {
    const int kp = pick_kp(k);
    const bool bf16 = idx->storage == EVS_STORE_BF16_F32;
    const void* scan_rows = bf16 ? idx->xb16 : (const void*)idx->xb32;
    ScanTuning tune;
    int profile;
    {
        std::lock_guard<std::mutex> lk(g_tune_mu);
        tune = g_tune;
        profile = g_profile_scans && !scan_only;
    }
    if (allow_tc && takes_tc_path(idx, nq, tune))
        return search_tc_locked(idx, nq, q_dev, k, out, st, scan_only, profile, tune.tc_pair_min_nq);
    const int qpp = max_queries_per_pass(idx->d, bf16);
    ScanPlan plan;
    CU(plan_scan(idx->ntotal, idx->d, bf16, kp, qpp, idx->sm_count, tune, &plan));
    const int64_t chunk_cap = nq < kQueryChunk ? nq : kQueryChunk;
    int rc = ensure_dev(reinterpret_cast<unsigned long long**>(&idx->lists), &idx->lists_cap, (size_t)chunk_cap * plan.grid * kp);
    if (rc) return rc;
    if (allow_tc && (rc = ensure_dev(&idx->margins_dev, &idx->margins_cap, (size_t)nq))) return rc;
    idx->last_nq = nq;

    for (int64_t c0 = 0; c0 < nq; c0 += kQueryChunk) {
        const int64_t cn = (nq - c0) < kQueryChunk ? (nq - c0) : kQueryChunk;
        const float* qc = q_dev + (size_t)c0 * idx->d;
        std::pair<cudaEvent_t, cudaEvent_t>* pe = nullptr;
        if (profile && idx->prof_used < 65536) {
            if (idx->prof_used == idx->prof_events.size()) {
                cudaEvent_t a = nullptr, b = nullptr;
                CU(cudaEventCreate(&a));
                CU(cudaEventCreate(&b));
                idx->prof_events.emplace_back(a, b);
            }
            pe = &idx->prof_events[idx->prof_used++];
            CU(cudaEventRecord(pe->first, st));
        }
        for (int64_t p0 = 0; p0 < cn; p0 += qpp) {
            ScanArgs a;
            a.xb = scan_rows;
            a.is_bf16 = bf16;
            a.n = idx->ntotal;
            a.d = idx->d;
            a.xq = qc;
            a.q0 = (int)p0;
            a.nq_pass = (int)((cn - p0) < qpp ? (cn - p0) : qpp);
            a.lists = idx->lists;
            a.kp = kp;
            // the plan's shared-memory size was computed for qpp queries per pass: large enough for fewer
            CU(launch_scan(a, &plan, st));
        }
        if (pe) CU(cudaEventRecord(pe->second, st));
        if (scan_only) continue;
        FinalizeArgs f;
        f.lists = idx->lists;
        f.L = plan.grid;
        f.kp = kp;
        f.xb = idx->xb32;
        f.xb_is_bf16 = 0;
        f.xq = qc;
        f.nq = cn;
        f.d = idx->d;
        f.k = (int)k;
        f.id_base = idx->id_base;
        f.D = out.D ? out.D + (size_t)c0 * k : nullptr;
        f.I = out.I ? out.I + (size_t)c0 * k : nullptr;
        f.P_scores = out.P_scores ? out.P_scores + (size_t)c0 * k : nullptr;
        f.P_ids = out.P_ids ? out.P_ids + (size_t)c0 * k : nullptr;
        f.margins = idx->margins_dev + c0;
        if (out.x) {
            f.x = *out.x;
            f.x.nq_total = nq;
            f.x.q_off = c0;
        }
        CU(launch_finalize(f, st));
    }
    return EVS_OK;
}

// ---- fill kernels for the empty-index case (faiss returns all padding) ----
__global__ void fill_final_kernel(float* D, long long* I, long long count) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        D[i] = -FLT_MAX;
        I[i] = -1;
    }
}
__global__ void fill_partial_kernel(double* S, long long* I, long long count) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        S[i] = -DBL_MAX;
        I[i] = -1;
    }
}

static int check_search_args(const evs_index* idx, int64_t nq, const void* q, int64_t k, const void* o1, const void* o2) {
    if (!idx) return fail(EVS_EINVAL, "idx is NULL");
    if (nq < 0) return fail(EVS_EINVAL, "nq must be >= 0");
    if (k <= 0) return fail(EVS_EINVAL, "k must be > 0 (got %lld)", (long long)k);  // FAISS_THROW_IF_NOT(k > 0)
    if (k > EVS_MAX_K) return fail(EVS_ELIMIT, "k = %lld exceeds EVS_MAX_K = %d", (long long)k, EVS_MAX_K);
    if (nq > 0 && (!q || !o1 || !o2)) return fail(EVS_EINVAL, "NULL query or output buffer");
    return EVS_OK;
}

// order stream `st` after the previous user of the handle's workspace, run, and mark the workspace busy until done
static int search_dev_common(evs_index* idx, int64_t nq, const float* q_dev, int64_t k, const SearchOut& out, cudaStream_t st) {
    CU(cudaStreamWaitEvent(st, idx->ws_free, 0));
    int rc;
    if (idx->ntotal == 0) {
        long long count = (long long)nq * k;
        int grid = (int)((count + 255) / 256 < 1024 ? (count + 255) / 256 : 1024);
        if (out.D) fill_final_kernel<<<grid, 256, 0, st>>>(out.D, reinterpret_cast<long long*>(out.I), count);
        else fill_partial_kernel<<<grid, 256, 0, st>>>(out.P_scores, reinterpret_cast<long long*>(out.P_ids), count);
        g_kernel_launches.fetch_add(1);
        CU(cudaGetLastError());
        idx->last_nq = 0;
        rc = EVS_OK;
    } else {
        rc = search_enqueue_locked(idx, nq, q_dev, k, out, st);
    }
    CU(cudaEventRecord(idx->ws_free, st));
    return rc;
}

extern "C" int evs_index_search_dev(evs_index* idx, int64_t nq, const float* q_dev, int64_t k, float* D_dev, int64_t* I_dev,
                                    void* stream) {
    int rc = check_search_args(idx, nq, q_dev, k, D_dev, I_dev);
    if (rc || nq == 0) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    if ((rc = use_device(idx->device))) return rc;
    SearchOut out;
    out.D = D_dev;
    out.I = I_dev;
    return search_dev_common(idx, nq, q_dev, k, out, (cudaStream_t)stream);
}

extern "C" int evs_index_search_partial_dev(evs_index* idx, int64_t nq, const float* q_dev, int64_t k, double* out_scores_dev,
                                            int64_t* out_ids_dev, void* stream) {
    int rc = check_search_args(idx, nq, q_dev, k, out_scores_dev, out_ids_dev);
    if (rc || nq == 0) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    if ((rc = use_device(idx->device))) return rc;
    SearchOut out;
    out.P_scores = out_scores_dev;
    out.P_ids = out_ids_dev;
    return search_dev_common(idx, nq, q_dev, k, out, (cudaStream_t)stream);
}

extern "C" int evs_merge_partials_dev(int device, int nparts, int64_t nq, int64_t k, const double* scores_dev,
                                      const int64_t* ids_dev, int64_t part_stride, float* D_dev, int64_t* I_dev, void* stream) {
    if (nparts <= 0 || nq < 0 || k <= 0) return fail(EVS_EINVAL, "bad nparts/nq/k");
    if (nq == 0) return EVS_OK;
    if (!scores_dev || !ids_dev || !D_dev || !I_dev) return fail(EVS_EINVAL, "NULL buffer");
    if ((size_t)nparts * k * 24 > 200 * 1024) return fail(EVS_ELIMIT, "nparts*k too large for one merge");
    int rc = use_device(device);
    if (rc) return rc;
    if (part_stride == 0) part_stride = nq * k;
    if (part_stride < nq * k) return fail(EVS_EINVAL, "part_stride smaller than nq*k");
    CU(launch_merge_partials(nparts, nq, (int)k, scores_dev, reinterpret_cast<const long long*>(ids_dev), part_stride, D_dev,
                             reinterpret_cast<long long*>(I_dev), (cudaStream_t)stream));
    return EVS_OK;
}

template <typename T>
static int ensure_pinned(T** ptr, size_t* cap, size_t need_elems) {
    if (*cap >= need_elems && *ptr) return EVS_OK;
    if (*ptr) CU(cudaFreeHost(*ptr));
    *ptr = nullptr;
    *cap = 0;
    CU(cudaMallocHost(reinterpret_cast<void**>(ptr), need_elems * sizeof(T)));
    *cap = need_elems;
    return EVS_OK;
}

// host staging shared by evs_index_search and evs_index_search_exchange: queries go caller memory -> pinned -> device in
// one async copy; (I, D) live in ONE device buffer ([I int64 nq*k][D float32 nq*k]) so that they come back in one copy.
static int host_stage_locked(evs_index* idx, int64_t nq, int64_t k) {
    const size_t d = (size_t)idx->d;
    int rc = ensure_pinned(&idx->q_pin, &idx->q_pin_cap, (size_t)nq * d);
    if (rc) return rc;
    if ((rc = ensure_dev(&idx->q_dev, &idx->q_cap, (size_t)nq * d))) return rc;
    const size_t out_need = (size_t)nq * k;
    if (idx->out_cap < out_need) {
        if (idx->I_dev) cudaFree(idx->I_dev);
        idx->I_dev = nullptr;
        idx->D_dev = nullptr;
        idx->out_cap = 0;
        CU(cudaMalloc(reinterpret_cast<void**>(&idx->I_dev), out_need * (sizeof(int64_t) + sizeof(float))));
        idx->D_dev = reinterpret_cast<float*>(idx->I_dev + out_need);
        idx->out_cap = out_need;
    } else {
        idx->D_dev = reinterpret_cast<float*>(idx->I_dev + out_need);  // packed for this nq*k
    }
    if (idx->out_pin_cap < out_need) {
        if (idx->I_pin) cudaFreeHost(idx->I_pin);
        idx->I_pin = nullptr;
        idx->D_pin = nullptr;
        idx->out_pin_cap = 0;
        CU(cudaMallocHost(reinterpret_cast<void**>(&idx->I_pin), out_need * (sizeof(int64_t) + sizeof(float))));
        idx->out_pin_cap = out_need;
    }
    idx->D_pin = reinterpret_cast<float*>(idx->I_pin + out_need);
    return EVS_OK;
}

static int host_fetch_locked(evs_index* idx, int64_t nq, int64_t k, float* D_host, int64_t* I_host, cudaStream_t st) {
    const size_t out_need = (size_t)nq * k;
    CU(cudaMemcpyAsync(idx->I_pin, idx->I_dev, out_need * (sizeof(int64_t) + sizeof(float)), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(D_host, idx->D_pin, out_need * sizeof(float));
    memcpy(I_host, idx->I_pin, out_need * sizeof(int64_t));
    return EVS_OK;
}

// fp32-storage indexes scan batches of 2+ queries in tf32 on the tensor cores.  The candidate set (k' = 64 or 128 rows by
// scan score) provably contains the exact top k when  margin = (canonical score of rank k) - (scan score of the worst retained
// candidate)  exceeds the scan's error: every row that was not retained scored below that candidate.  tf32 products carry
// ~2^-11 relative truncation per operand: a bias the margin already subtracts (finalize_kernel) plus ~4e-5 rms of scatter on
// unit vectors; queries whose margin is below tf32_guard_eps (1.5e-4: dense near-ties around rank k, e.g. bursts of
// near-duplicate images) are re-run with the fp32 GEMV scan.  Host API only: the device
// API cannot look at the margins without a synchronisation (evs_index_last_margins is there for its callers).
static bool tf32_guard_applies(const evs_index* idx, int64_t nq) {
    if (idx->storage != EVS_STORE_F32 || g_tf32_guard_eps_e6 <= 0 || idx->ntotal == 0) return false;
    ScanTuning tune;
    {
        std::lock_guard<std::mutex> lkt(g_tune_mu);
        tune = g_tune;
    }
    return takes_tc_path(idx, nq, tune);  // the GEMV scan accumulates in fp32: nothing to certify
}

// results (and, when the guard applies, the margins) come back in the same synchronisation; only if a query is not
// certified is there a second round
static int host_fetch_guarded_locked(evs_index* idx, int64_t nq, int64_t k, float* D_host, int64_t* I_host, cudaStream_t st) {
    if (!tf32_guard_applies(idx, nq)) return host_fetch_locked(idx, nq, k, D_host, I_host, st);
    int rc = ensure_pinned(&idx->m_pin, &idx->m_pin_cap, (size_t)nq);
    if (rc) return rc;
    CU(cudaMemcpyAsync(idx->m_pin, idx->margins_dev, (size_t)nq * sizeof(float), cudaMemcpyDeviceToHost, st));
    if ((rc = host_fetch_locked(idx, nq, k, D_host, I_host, st))) return rc;  // synchronises
    const float eps = (float)g_tf32_guard_eps_e6 * 1e-6f;
    int64_t reruns = 0;
    for (int64_t q = 0; q < nq; q++) {
        if (idx->m_pin[q] >= eps) continue;  // certified (also +inf: every row was a candidate); NaN margins re-run
        SearchOut o1;
        o1.D = idx->D_dev + (size_t)q * k;
        o1.I = idx->I_dev + (size_t)q * k;
        float* keep = idx->margins_dev;
        idx->margins_dev = keep + q;  // the re-run writes this query's margin in place
        rc = search_enqueue_locked(idx, 1, idx->q_dev + (size_t)q * idx->d, k, o1, st, false, false);
        idx->margins_dev = keep;
        idx->last_nq = nq;
        if (rc) return rc;
        reruns++;
    }
    if (!reruns) return EVS_OK;
    g_exact_reruns.fetch_add(reruns);
    CU(cudaEventRecord(idx->ws_free, st));
    return host_fetch_locked(idx, nq, k, D_host, I_host, st);
}

extern "C" int evs_index_search(evs_index* idx, int64_t nq, const float* q_host, int64_t k, float* D_host, int64_t* I_host) {
    int rc = check_search_args(idx, nq, q_host, k, D_host, I_host);
    if (rc || nq == 0) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    if ((rc = use_device(idx->device))) return rc;
    if ((rc = host_stage_locked(idx, nq, k))) return rc;
    memcpy(idx->q_pin, q_host, (size_t)nq * idx->d * sizeof(float));
    cudaStream_t st = idx->stream;
    CU(cudaStreamWaitEvent(st, idx->ws_free, 0));
    CU(cudaMemcpyAsync(idx->q_dev, idx->q_pin, (size_t)nq * idx->d * sizeof(float), cudaMemcpyHostToDevice, st));
    SearchOut out;
    out.D = idx->D_dev;
    out.I = idx->I_dev;
    if ((rc = search_dev_common(idx, nq, idx->q_dev, k, out, st))) return rc;
    return host_fetch_guarded_locked(idx, nq, k, D_host, I_host, st);
}

// ---------------------------------------------------------------------------------------------
// peer-store exchange of shard partials (row sharding, one process per GPU)
// ---------------------------------------------------------------------------------------------
struct evs_exchange {
    int device = 0, rank = 0, world = 0;
    int64_t max_nq = 0, max_k = 0;
    size_t slot_bytes = 0, total_bytes = 0;
    unsigned char* local = nullptr;     // [2][world][slot_bytes] slots, [2][world] u64 flags, done counter, timeout flag
    unsigned char* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool opened[8] = {false, false, false, false, false, false, false, false};
    bool connected = false;
    unsigned long long seq = 0;
    double* stage_scores = nullptr;     // local partial staging for the paths that cannot write the slots themselves
    int64_t* stage_ids = nullptr;
    std::mutex mu;
    size_t flags_off() const { return 2 * (size_t)world * slot_bytes; }
    unsigned* done() const { return reinterpret_cast<unsigned*>(local + flags_off() + 2 * (size_t)world * 8); }
    int* timed_out() const { return reinterpret_cast<int*>(local + flags_off() + 2 * (size_t)world * 8 + 8); }
};

extern "C" int evs_exchange_create(int device, int rank, int world, int64_t max_nq, int64_t max_k, evs_exchange** out) {
    if (!out) return fail(EVS_EINVAL, "out is NULL");
    *out = nullptr;
    if (world < 1 || world > 8) return fail(EVS_ELIMIT, "world must be in [1, 8] (one NVSwitch box), got %d", world);
    if (rank < 0 || rank >= world) return fail(EVS_EINVAL, "rank %d out of range for world %d", rank, world);
    if (max_nq <= 0 || max_k <= 0 || max_k > EVS_MAX_K) return fail(EVS_EINVAL, "bad max_nq/max_k");
    int ndev = 0;
    evs_device_count(&ndev);
    if (ndev <= 0) return fail(EVS_ENODEV, "no CUDA device: libevs has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(EVS_EINVAL, "device %d out of range (have %d)", device, ndev);
    int rc = use_device(device);
    if (rc) return rc;
    evs_exchange* ex = new (std::nothrow) evs_exchange();
    if (!ex) return fail(EVS_ENOMEM, "out of host memory");
    ex->device = device;
    ex->rank = rank;
    ex->world = world;
    ex->max_nq = max_nq;
    ex->max_k = max_k;
    ex->slot_bytes = ((size_t)max_nq * max_k * 16 + 255) & ~(size_t)255;
    ex->total_bytes = ex->flags_off() + 2 * (size_t)world * 8 + 16;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ex->local), ex->total_bytes);
    if (e == cudaSuccess) e = cudaMemset(ex->local, 0, ex->total_bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ex->stage_scores), (size_t)max_nq * max_k * 8);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ex->stage_ids), (size_t)max_nq * max_k * 8);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(ex->local);
        cudaFree(ex->stage_scores);
        cudaFree(ex->stage_ids);
        delete ex;
        return fail(e == cudaErrorMemoryAllocation ? EVS_ENOMEM : EVS_ECUDA, "exchange allocation failed: %s", cudaGetErrorString(e));
    }
    ex->peer[rank] = ex->local;
    ex->connected = (world == 1);
    *out = ex;
    return EVS_OK;
}

extern "C" int evs_exchange_handle(evs_exchange* ex, void* handle_out, int64_t handle_bytes) {
    if (!ex || !handle_out) return fail(EVS_EINVAL, "NULL argument");
    if (handle_bytes != EVS_IPC_HANDLE_BYTES) return fail(EVS_EINVAL, "handle buffer must be %d bytes", EVS_IPC_HANDLE_BYTES);
    static_assert(sizeof(cudaIpcMemHandle_t) == EVS_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    int rc = use_device(ex->device);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ex->local));
    memcpy(handle_out, &h, sizeof(h));
    return EVS_OK;
}

extern "C" int evs_exchange_connect(evs_exchange* ex, const void* handles, int64_t handles_bytes) {
    if (!ex || !handles) return fail(EVS_EINVAL, "NULL argument");
    if (handles_bytes != (int64_t)ex->world * EVS_IPC_HANDLE_BYTES)
        return fail(EVS_EINVAL, "expected %d handles of %d bytes", ex->world, EVS_IPC_HANDLE_BYTES);
    std::lock_guard<std::mutex> lk(ex->mu);
    int rc = use_device(ex->device);
    if (rc) return rc;
    for (int g = 0; g < ex->world; g++) {
        if (g == ex->rank || ex->opened[g]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, reinterpret_cast<const unsigned char*>(handles) + (size_t)g * EVS_IPC_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));  // maps the peer's buffer over NVLink
        ex->peer[g] = reinterpret_cast<unsigned char*>(p);
        ex->opened[g] = true;
    }
    ex->connected = true;
    return EVS_OK;
}

extern "C" int evs_exchange_status(evs_exchange* ex, int* timed_out, int64_t* searches) {
    if (!ex) return fail(EVS_EINVAL, "ex is NULL");
    std::lock_guard<std::mutex> lk(ex->mu);
    int rc = use_device(ex->device);
    if (rc) return rc;
    if (timed_out) CU(cudaMemcpy(timed_out, ex->timed_out(), sizeof(int), cudaMemcpyDeviceToHost));
    if (searches) *searches = (int64_t)ex->seq;
    return EVS_OK;
}

extern "C" int evs_exchange_free(evs_exchange* ex) {
    if (!ex) return EVS_OK;
    cudaSetDevice(ex->device);
    cudaDeviceSynchronize();
    for (int g = 0; g < ex->world; g++)
        if (ex->opened[g]) cudaIpcCloseMemHandle(ex->peer[g]);
    cudaFree(ex->local);
    cudaFree(ex->stage_scores);
    cudaFree(ex->stage_ids);
    delete ex;
    return EVS_OK;
}

// enqueue scan -> finalise (-> publish) -> merge for one exchange-mode search; idx->mu and ex->mu are held
static int search_exchange_enqueue_locked(evs_index* idx, evs_exchange* ex, int64_t nq, const float* q_dev, int64_t k, float* D_dev,
                                          int64_t* I_dev, cudaStream_t st) {
    Exchange x;
    for (int g = 0; g < ex->world; g++) x.peer[g] = ex->peer[g];
    x.rank = ex->rank;
    x.world = ex->world;
    x.seq = ++ex->seq;              // every rank calls in the same order: the sequence numbers agree
    x.parity = (int)(x.seq & 1ull);  // two generations of slots: a fast rank may start search s+1 while a slow one merges s
    x.slot_bytes = ex->slot_bytes;
    x.done = ex->done();
    x.nq_total = nq;
    ScanTuning tune;
    {
        std::lock_guard<std::mutex> lkt(g_tune_mu);
        tune = g_tune;
    }
    SearchOut out;
    int rc;
    if (idx->ntotal == 0 || (takes_tc_path(idx, nq, tune) && !tc_path_is_heap(idx, nq, k, tune))) {
        // empty shard / tensor-core scan with host-side overflow repair: partial into local staging, then publish
        out.P_scores = ex->stage_scores;
        out.P_ids = ex->stage_ids;
        if ((rc = search_dev_common(idx, nq, q_dev, k, out, st))) return rc;
        CU(launch_publish_partials(x, nq, (int)k, ex->stage_scores, reinterpret_cast<const long long*>(ex->stage_ids), st));
    } else {
        out.x = &x;  // the finalise kernel stores this shard's k best straight into every rank's slot
        if ((rc = search_dev_common(idx, nq, q_dev, k, out, st))) return rc;
    }
    CU(launch_merge_exchange(x, nq, (int)k, D_dev, reinterpret_cast<long long*>(I_dev), ex->timed_out(), st));
    return EVS_OK;
}

static int check_exchange_args(const evs_index* idx, const evs_exchange* ex, int64_t nq, int64_t k) {
    if (!ex) return fail(EVS_EINVAL, "ex is NULL");
    if (!ex->connected) return fail(EVS_EINVAL, "exchange is not connected (evs_exchange_connect)");
    if (ex->device != idx->device) return fail(EVS_EINVAL, "exchange lives on device %d, index on %d", ex->device, idx->device);
    if (nq > ex->max_nq || k > ex->max_k)
        return fail(EVS_ELIMIT, "nq=%lld k=%lld exceed the exchange's capacity (%lld, %lld)", (long long)nq, (long long)k,
                    (long long)ex->max_nq, (long long)ex->max_k);
    return EVS_OK;
}

extern "C" int evs_index_search_exchange_dev(evs_index* idx, evs_exchange* ex, int64_t nq, const float* q_dev, int64_t k,
                                             float* D_dev, int64_t* I_dev, void* stream) {
    int rc = check_search_args(idx, nq, q_dev, k, D_dev, I_dev);
    if (rc || nq == 0) return rc;
    if ((rc = check_exchange_args(idx, ex, nq, k))) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    std::lock_guard<std::mutex> lkx(ex->mu);
    if ((rc = use_device(idx->device))) return rc;
    return search_exchange_enqueue_locked(idx, ex, nq, q_dev, k, D_dev, I_dev, (cudaStream_t)stream);
}

extern "C" int evs_index_search_exchange(evs_index* idx, evs_exchange* ex, int64_t nq, const float* q_host, int64_t k, float* D_host,
                                         int64_t* I_host) {
    int rc = check_search_args(idx, nq, q_host, k, D_host, I_host);
    if (rc || nq == 0) return rc;
    if ((rc = check_exchange_args(idx, ex, nq, k))) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    std::lock_guard<std::mutex> lkx(ex->mu);
    if ((rc = use_device(idx->device))) return rc;
    if ((rc = host_stage_locked(idx, nq, k))) return rc;
    memcpy(idx->q_pin, q_host, (size_t)nq * idx->d * sizeof(float));
    cudaStream_t st = idx->stream;
    CU(cudaStreamWaitEvent(st, idx->ws_free, 0));
    CU(cudaMemcpyAsync(idx->q_dev, idx->q_pin, (size_t)nq * idx->d * sizeof(float), cudaMemcpyHostToDevice, st));
    if ((rc = search_exchange_enqueue_locked(idx, ex, nq, idx->q_dev, k, idx->D_dev, idx->I_dev, st))) return rc;
    return host_fetch_locked(idx, nq, k, D_host, I_host, st);
}

extern "C" int evs_index_last_margins(evs_index* idx, int64_t nq, float* margins_host) {
    if (!idx || !margins_host) return fail(EVS_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(idx->mu);
    if (nq != idx->last_nq) return fail(EVS_EINVAL, "last search had %lld queries, not %lld", (long long)idx->last_nq, (long long)nq);
    if (nq == 0) return EVS_OK;
    int rc = use_device(idx->device);
    if (rc) return rc;
    CU(cudaEventSynchronize(idx->ws_free));
    CU(cudaMemcpy(margins_host, idx->margins_dev, (size_t)nq * sizeof(float), cudaMemcpyDeviceToHost));
    return EVS_OK;
}

extern "C" int evs_index_time_scan(evs_index* idx, int64_t nq, const float* q_dev, int64_t k, int iters, float* mean_ms) {
    int rc = check_search_args(idx, nq, q_dev, k, mean_ms, mean_ms);
    if (rc) return rc;
    if (nq == 0 || iters <= 0 || idx->ntotal == 0) return fail(EVS_EINVAL, "nothing to time");
    std::lock_guard<std::mutex> lk(idx->mu);
    if ((rc = use_device(idx->device))) return rc;
    cudaStream_t st = idx->stream;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaStreamWaitEvent(st, idx->ws_free, 0));
    SearchOut none;
    rc = search_enqueue_locked(idx, nq, q_dev, k, none, st, true);  // warm-up
    if (!rc) {
        cudaEventRecord(e0, st);
        for (int i = 0; i < iters && !rc; i++) rc = search_enqueue_locked(idx, nq, q_dev, k, none, st, true);
        cudaEventRecord(e1, st);
    }
    cudaEventRecord(idx->ws_free, st);
    cudaError_t e = cudaStreamSynchronize(st);
    float ms = 0.f;
    if (!rc && e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(EVS_ECUDA, "scan failed: %s", cudaGetErrorString(e));
    *mean_ms = ms / iters;
    return EVS_OK;
}

extern "C" int evs_index_scan_profile(evs_index* idx, int64_t* count, double* total_ms) {
    if (!idx || !count || !total_ms) return fail(EVS_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(idx->mu);
    int rc = use_device(idx->device);
    if (rc) return rc;
    double sum = 0.0;
    for (size_t i = 0; i < idx->prof_used; i++) {
        CU(cudaEventSynchronize(idx->prof_events[i].second));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, idx->prof_events[i].first, idx->prof_events[i].second));
        sum += ms;
    }
    *count = (int64_t)idx->prof_used;
    *total_ms = sum;
    idx->prof_used = 0;
    return EVS_OK;
}

extern "C" int evs_index_tc_max_queries(const evs_index* idx, int* max_queries) {
    if (!idx || !max_queries) return fail(EVS_EINVAL, "NULL argument");
    *max_queries = tc_max_queries(idx->d, idx->storage == EVS_STORE_BF16_F32);
    return EVS_OK;
}

extern "C" int evs_index_tc_scores_dev(evs_index* idx, int64_t nq, const float* q_dev, float* out_dev, int* npad, void* stream) {
    if (!idx || !q_dev || !npad) return fail(EVS_EINVAL, "NULL argument");
    const bool bf16 = idx->storage == EVS_STORE_BF16_F32;
    const int nb_max = tc_max_queries(idx->d, bf16);
    if (nb_max == 0) return fail(EVS_ELIMIT, "d = %d is not supported by the tensor-core scan", idx->d);
    int pair_min;
    {
        std::lock_guard<std::mutex> lk(g_tune_mu);
        pair_min = g_tune.tc_pair_min_nq;
    }
    const bool pair = pair_min > 0 && (nq >= pair_min || nq > nb_max) && tc2_max_half(idx->d, bf16) > 0;
    if (nq <= 0 || (!pair && nq > nb_max) || nq > 4096) return fail(EVS_EINVAL, "nq must be in [1, %d]", pair ? 4096 : nb_max);
    if (idx->ntotal == 0) return fail(EVS_EINVAL, "empty index");
    std::lock_guard<std::mutex> lk(idx->mu);
    int rc = use_device(idx->device);
    if (rc) return rc;
    TcArgs a;
    a.xb = bf16 ? idx->xb16 : (const void*)idx->xb32;
    a.is_bf16 = bf16;
    a.n = idx->ntotal;
    a.d = idx->d;
    a.xq = q_dev;
    a.nq = (int)nq;
    a.lists = nullptr;
    a.overflow_out = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    if (pair) {
        Tc2Plan pl;
        CU(tc2_plan(idx->ntotal, idx->d, bf16, (int)nq, 64, idx->sm_count, &pl));
        *npad = pl.nqp;
        if (!out_dev) return EVS_OK;  // pitch query
        if ((rc = ensure_dev(&idx->tc_ws, &idx->tc_ws_cap, tc2_workspace_bytes(pl)))) return rc;
        CU(cudaStreamWaitEvent(st, idx->ws_free, 0));
        CU(tc2_dump_scores(a, pl, idx->tc_ws, out_dev, st));
        *npad = pl.nqp;
    } else {
        TcPlan pl;
        CU(tc_plan(idx->ntotal, idx->d, bf16, (int)nq, 64, idx->sm_count, &pl));
        *npad = pl.npad;
        if (!out_dev) return EVS_OK;  // pitch query
        if ((rc = ensure_dev(&idx->tc_ws, &idx->tc_ws_cap, tc_workspace_bytes(pl)))) return rc;
        CU(cudaStreamWaitEvent(st, idx->ws_free, 0));
        CU(tc_dump_scores(a, pl, idx->tc_ws, out_dev, st));
    }
    CU(cudaEventRecord(idx->ws_free, st));
    return EVS_OK;
}

// ---------------------------------------------------------------------------------------------
// stand-alone kernels
// ---------------------------------------------------------------------------------------------
static int device_sm_count(int device, int* sms) {
    CU(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, device));
    return EVS_OK;
}

extern "C" int evs_l2_normalize_dev(int device, void* x_dev, int64_t n, int d, int dtype, void* stream) {
    if (n < 0 || d <= 0) return fail(EVS_EINVAL, "bad n/d");
    if (dtype != EVS_F32 && dtype != EVS_F16 && dtype != EVS_BF16) return fail(EVS_EINVAL, "bad dtype %d", dtype);
    if (n == 0) return EVS_OK;
    if (!x_dev) return fail(EVS_EINVAL, "x is NULL");
    int ndev = 0;
    evs_device_count(&ndev);
    if (ndev <= 0) return fail(EVS_ENODEV, "no CUDA device: libevs has no CPU fallback");
    int rc = use_device(device), sms = 0;
    if (rc || (rc = device_sm_count(device, &sms))) return rc;
    if ((size_t)d * 4 * 8 > 200 * 1024) return fail(EVS_ELIMIT, "d = %d too large for the normalise kernel", d);
    CU(launch_l2_normalize(x_dev, n, d, dtype, sms, (cudaStream_t)stream));
    return EVS_OK;
}

extern "C" int evs_l2_normalize(int device, float* x_host, int64_t n, int d) {
    if (n < 0 || d <= 0) return fail(EVS_EINVAL, "bad n/d");
    if (n == 0) return EVS_OK;
    if (!x_host) return fail(EVS_EINVAL, "x is NULL");
    int ndev = 0;
    evs_device_count(&ndev);
    if (ndev <= 0) return fail(EVS_ENODEV, "no CUDA device: libevs has no CPU fallback");
    int rc = use_device(device);
    if (rc) return rc;
    float* dev = nullptr;
    size_t bytes = (size_t)n * d * sizeof(float);
    CU(cudaMalloc(reinterpret_cast<void**>(&dev), bytes));
    cudaError_t e = cudaMemcpy(dev, x_host, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = evs_l2_normalize_dev(device, dev, n, d, EVS_F32, nullptr);
        if (!rc) e = cudaMemcpy(x_host, dev, bytes, cudaMemcpyDeviceToHost);  // default-stream copy orders after the kernel
    }
    cudaFree(dev);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(EVS_ECUDA, "copy failed: %s", cudaGetErrorString(e));
    return EVS_OK;
}

extern "C" int evs_f32_to_bf16_dev(int device, const float* src_dev, void* dst_dev, int64_t count, void* stream) {
    if (count < 0) return fail(EVS_EINVAL, "bad count");
    if (count == 0) return EVS_OK;
    if (!src_dev || !dst_dev) return fail(EVS_EINVAL, "NULL buffer");
    int rc = use_device(device), sms = 0;
    if (rc || (rc = device_sm_count(device, &sms))) return rc;
    CU(launch_f32_to_bf16(src_dev, dst_dev, count, sms, (cudaStream_t)stream));
    return EVS_OK;
}

// ---------------------------------------------------------------------------------------------
// persistence: index.faiss (45-byte IndexFlat header + fp32 payload, little-endian)
// ---------------------------------------------------------------------------------------------
#pragma pack(push, 1)
struct FlatHeader {
    char fourcc[4];
    int32_t d;
    int64_t ntotal;
    int64_t dummy1, dummy2;
    uint8_t is_trained;
    int32_t metric_type;
    uint64_t count;  // number of float32 values
};
#pragma pack(pop)
static_assert(sizeof(FlatHeader) == 45, "index.faiss flat header is 45 bytes");

static const size_t kIoChunk = (size_t)64 << 20;

extern "C" int evs_index_write(const evs_index* idx, const char* path) {
    if (!idx || !path) return fail(EVS_EINVAL, "NULL argument");
    int rc = use_device(idx->device);
    if (rc) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(EVS_EIO, "cannot open '%s' for writing: %s", path, strerror(errno));
    FlatHeader h;
    memcpy(h.fourcc, "IxFI", 4);
    h.d = idx->d;
    h.ntotal = idx->ntotal;
    h.dummy1 = h.dummy2 = (int64_t)1 << 20;
    h.is_trained = 1;
    h.metric_type = 0;
    h.count = (uint64_t)idx->ntotal * (uint64_t)idx->d;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    size_t total = (size_t)h.count * sizeof(float);
    void* pin = nullptr;
    if (ok && total) {
        size_t chunk = total < kIoChunk ? total : kIoChunk;
        cudaError_t e = cudaMallocHost(&pin, chunk);
        if (e != cudaSuccess) {
            fclose(f);
            return fail(EVS_ENOMEM, "cudaMallocHost failed: %s", cudaGetErrorString(e));
        }
        for (size_t off = 0; ok && off < total; off += chunk) {
            size_t nb = total - off < chunk ? total - off : chunk;
            e = cudaMemcpy(pin, reinterpret_cast<const unsigned char*>(idx->xb32) + off, nb, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) {
                cudaFreeHost(pin);
                fclose(f);
                return fail(EVS_ECUDA, "device read failed: %s", cudaGetErrorString(e));
            }
            ok = fwrite(pin, 1, nb, f) == nb;
        }
        cudaFreeHost(pin);
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) return fail(EVS_EIO, "short write to '%s'", path);
    return EVS_OK;
}

extern "C" int evs_index_read(const char* path, int device, int storage, evs_index** out) {
    if (!path || !out) return fail(EVS_EINVAL, "NULL argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(EVS_EIO, "cannot open '%s': %s", path, strerror(errno));
    FlatHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1) {
        fclose(f);
        return fail(EVS_EFORMAT, "'%s': truncated header", path);
    }
    const bool ip = !memcmp(h.fourcc, "IxFI", 4);
    if (!ip && memcmp(h.fourcc, "IxF2", 4) && memcmp(h.fourcc, "IxFl", 4)) {
        fclose(f);
        return fail(EVS_EFORMAT, "'%s': not a flat index (fourcc %.4s)", path, h.fourcc);
    }
    if (h.metric_type != 0) {
        fclose(f);
        return fail(EVS_EFORMAT, "'%s': metric %d is not inner product", path, h.metric_type);
    }
    if (h.d <= 0 || h.ntotal < 0 || h.count >= ((uint64_t)1 << 40) || h.count != (uint64_t)h.ntotal * (uint64_t)h.d) {
        fclose(f);
        return fail(EVS_EFORMAT, "'%s': inconsistent header (d=%d ntotal=%lld count=%llu)", path, h.d, (long long)h.ntotal,
                    (unsigned long long)h.count);
    }
    struct stat sb;
    if (fstat(fileno(f), &sb) == 0 && (uint64_t)sb.st_size < sizeof(h) + h.count * 4) {
        fclose(f);
        return fail(EVS_EFORMAT, "'%s': truncated payload", path);
    }
    evs_index* idx = nullptr;
    int rc = evs_index_create(h.d, device, storage, &idx);
    if (rc) {
        fclose(f);
        return rc;
    }
    size_t total = (size_t)h.count * sizeof(float);
    if (total) {
        {
            std::lock_guard<std::mutex> lk(idx->mu);
            rc = grow_locked(idx, h.ntotal);
        }
        void* pin[2] = {nullptr, nullptr};
        size_t chunk = total < kIoChunk ? total : kIoChunk;
        if (!rc && (cudaMallocHost(&pin[0], chunk) != cudaSuccess || cudaMallocHost(&pin[1], chunk) != cudaSuccess))
            rc = fail(EVS_ENOMEM, "cudaMallocHost failed");
        // double-buffered: fread into one pinned buffer while the other is in flight to the device
        cudaEvent_t done[2] = {nullptr, nullptr};
        if (!rc) {
            cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming);
        }
        int b = 0;
        for (size_t off = 0; !rc && off < total; off += chunk, b ^= 1) {
            size_t nb = total - off < chunk ? total - off : chunk;
            cudaEventSynchronize(done[b]);
            if (fread(pin[b], 1, nb, f) != nb) {
                rc = fail(EVS_EFORMAT, "'%s': short read", path);
                break;
            }
            cudaError_t e = cudaMemcpyAsync(reinterpret_cast<unsigned char*>(idx->xb32) + off, pin[b], nb, cudaMemcpyHostToDevice,
                                            idx->stream);
            if (e == cudaSuccess) e = cudaEventRecord(done[b], idx->stream);
            if (e != cudaSuccess) rc = fail(EVS_ECUDA, "upload failed: %s", cudaGetErrorString(e));
        }
        if (!rc) {
            std::lock_guard<std::mutex> lk(idx->mu);
            rc = finish_add_locked(idx, h.ntotal);
        } else {
            cudaStreamSynchronize(idx->stream);
        }
        if (done[0]) cudaEventDestroy(done[0]);
        if (done[1]) cudaEventDestroy(done[1]);
        cudaFreeHost(pin[0]);
        cudaFreeHost(pin[1]);
    }
    fclose(f);
    if (rc) {
        evs_index_free(idx);
        return rc;
    }
    *out = idx;
    return EVS_OK;
}
