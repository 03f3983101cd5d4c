/*
 * oracle/flat_ip_oracle.c -- CPU restatement of the evo-ssearch similarity-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or the timed CPU baseline.  The product path
 * (evo-ssearch_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED.  The arithmetic of the reference's hot path does not live in
 * /root/reference: oldapp.py:87-88 (IndexFlatIP + add), oldapp.py:2005 and :2112 (search) call the
 * third-party wheel `faiss-cpu` (requirements.txt:6, ">=1.7.4", no lock file).  That wheel and its
 * source are absent from this image and there is no network, and the reference ships no tests,
 * golden vectors or fixtures.  What follows restates faiss's *published* flat inner-product
 * algorithm (IndexFlat::search -> knn_inner_product -> exhaustive_inner_product_seq with a
 * k-entry binary min-heap result handler, faiss/utils/distances.cpp, faiss/utils/Heap.h,
 * faiss/impl/ResultHandler.h as of the 1.7.4 line) from knowledge of that source, anchored on the
 * reference's own call sites.  It could not be checked against faiss itself here.
 *
 * Two rankings are provided:
 *   orc_faiss_seq_search   the faiss restatement: fp32 dot products, min-heap with strict-greater
 *                          replacement, (score,id)-lexicographic heap order, descending output,
 *                          (-FLT_MAX,-1) padding.
 *   orc_canon_search       the canonical ranking the CUDA path is held to bit-exactly: the inner
 *                          product of the fp32 inputs accumulated in fp64 in a FIXED order
 *                          (CANON-32, below), ranked by (score desc, id asc), score rounded to
 *                          fp32 on output.  It is independent of tile shape, query batch size and
 *                          shard count; it differs from any fp32 evaluation order (faiss's own is
 *                          ISA dependent) only where two neighbours are closer than the fp32
 *                          accumulation error.
 *
 * CANON-32 dot product of x,q in R^d (fp32 inputs):
 *   lane l in 0..31:  p[l] = sum over j = 0,1,.. while l+32j < d of (double)x[l+32j]*(double)q[l+32j],
 *                     added in increasing j starting from +0.0 (every product is exact in fp64);
 *   butterfly:        for off in 16,8,4,2,1:  p[l] = p[l] + p[l ^ off]   (all l at once);
 *   result:           p[0]  (all lanes hold the same bits; fp add is commutative).
 * This is exactly what one CUDA warp computes with 32 DFMA chains and five xor-shuffles.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * fp32 inner product, scalar definition (faiss utils/distances_simd.cpp fvec_inner_product:
 * "res += x[i] * y[i]" over i; the shipped wheel vectorises this loop, which changes the
 * summation order in an ISA-dependent way -- the reason parity is judged near-tie-aware).
 * Built with -ffp-contract=off so that this one is the strict sequential order.
 * ------------------------------------------------------------------------------------------ */
static float dot_f32_sequential(const float* x, const float* y, int64_t d) {
    float res = 0.0f;
    for (int64_t i = 0; i < d; i++) res += x[i] * y[i];
    return res;
}

/* fp32 inner product the way an auto-vectorised build evaluates it (lane-parallel partial sums).
 * Used for the timed CPU baseline; results differ from the sequential order in the last bits. */
static float dot_f32_simd(const float* x, const float* y, int64_t d) {
    float res = 0.0f;
#pragma omp simd reduction(+ : res)
    for (int64_t i = 0; i < d; i++) res += x[i] * y[i];
    return res;
}

/* CANON-32 (see header). */
static double dot_canon32(const float* x, const float* q, int64_t d) {
    double p[32];
    for (int l = 0; l < 32; l++) {
        double acc = 0.0;
        for (int64_t i = l; i < d; i += 32) acc += (double)x[i] * (double)q[i];
        p[l] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        double t[32];
        for (int l = 0; l < 32; l++) t[l] = p[l] + p[l ^ off];
        memcpy(p, t, sizeof(p));
    }
    return p[0];
}

ORC_API double orc_dot_canon32(const float* x, const float* q, int64_t d) { return dot_canon32(x, q, d); }
ORC_API float orc_dot_f32_sequential(const float* x, const float* q, int64_t d) {
    return dot_f32_sequential(x, q, d);
}

/* ------------------------------------------------------------------------------------------
 * k-entry binary MIN-heap over (score, id), 1-based sift like faiss utils/Heap.h with
 * CMin<float,int64>: the root is the smallest (score, id) in lexicographic order; a candidate
 * enters only if root_score < score (STRICT, ResultHandler.h add_result); reorder pops the root
 * to the back, giving descending output, then pads with (-FLT_MAX, -1).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t k;
    float* val; /* k entries */
    int64_t* ids;
} minheap_t;

static inline int pair_less(float a, int64_t ia, float b, int64_t ib) {
    return (a < b) || (a == b && ia < ib);
}

static void heap_fill_neutral(minheap_t* h) {
    for (int64_t i = 0; i < h->k; i++) {
        h->val[i] = -FLT_MAX;
        h->ids[i] = -1;
    }
}

/* replace the root by (v,id) and sift down over the first n entries */
static void heap_replace_root(int64_t n, float* val, int64_t* ids, float v, int64_t id) {
    float* hv = val - 1; /* 1-based */
    int64_t* hi = ids - 1;
    int64_t i = 1;
    for (;;) {
        int64_t c1 = i << 1, c2 = c1 + 1;
        if (c1 > n) break;
        int64_t c = c1;
        if (c2 <= n && !pair_less(hv[c1], hi[c1], hv[c2], hi[c2])) c = c2;
        if (pair_less(v, id, hv[c], hi[c])) break;
        hv[i] = hv[c];
        hi[i] = hi[c];
        i = c;
    }
    hv[i] = v;
    hi[i] = id;
}

/* remove the root of an n-entry heap (the last entry is re-inserted from the top) */
static void heap_pop_root(int64_t n, float* val, int64_t* ids) {
    float v = val[n - 1];
    int64_t id = ids[n - 1];
    heap_replace_root(n - 1, val, ids, v, id);
}

/* heap -> descending list, neutral padding at the tail; returns number of real entries */
static int64_t heap_to_sorted(minheap_t* h) {
    int64_t k = h->k, filled = 0;
    for (int64_t i = 0; i < k; i++) {
        float v = h->val[0];
        int64_t id = h->ids[0];
        heap_pop_root(k - i, h->val, h->ids);
        h->val[k - filled - 1] = v;
        h->ids[k - filled - 1] = id;
        if (id != -1) filled++;
    }
    memmove(h->val, h->val + k - filled, (size_t)filled * sizeof(float));
    memmove(h->ids, h->ids + k - filled, (size_t)filled * sizeof(int64_t));
    for (int64_t i = filled; i < k; i++) {
        h->val[i] = -FLT_MAX;
        h->ids[i] = -1;
    }
    return filled;
}

static inline void heap_offer(minheap_t* h, float score, int64_t id) {
    if (h->val[0] < score) heap_replace_root(h->k, h->val, h->ids, score, id);
}

/* ------------------------------------------------------------------------------------------
 * faiss restatement, seq path (nq < 20 in faiss; used here for any nq): OpenMP over QUERIES
 * only -- a single query scans the whole database on one thread, as faiss does.
 *   simd != 0 selects the vectorised dot (timed baseline), 0 the strict sequential one.
 * Returns 0, or -1 on bad arguments (k <= 0 mirrors FAISS_THROW_IF_NOT(k > 0)).
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_faiss_seq_search(const float* xq, const float* xb, int64_t d, int64_t nq, int64_t nb, int64_t k,
                                 float* D, int64_t* I, int simd, int nthreads) {
    if (k <= 0 || d <= 0 || nq < 0 || nb < 0) return -1;
#ifdef _OPENMP
    int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
    if (nt > nq) nt = (int)(nq > 0 ? nq : 1);
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (int64_t i = 0; i < nq; i++) {
        minheap_t h = {k, D + i * k, I + i * k};
        heap_fill_neutral(&h);
        const float* q = xq + i * d;
        if (simd) {
            for (int64_t j = 0; j < nb; j++) heap_offer(&h, dot_f32_simd(q, xb + j * d, d), j);
        } else {
            for (int64_t j = 0; j < nb; j++) heap_offer(&h, dot_f32_sequential(q, xb + j * d, d), j);
        }
        heap_to_sorted(&h);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * All-cores variant of the same scan (NOT what faiss does for one query): database rows are
 * split across threads, each keeps its own heap, the per-thread lists are merged through one
 * more heap in thread order.  Reported beside the faiss-like number so the CPU is not
 * handicapped by query-only parallelism.
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_allcores_search(const float* xq, const float* xb, int64_t d, int64_t nq, int64_t nb, int64_t k,
                                float* D, int64_t* I, int nthreads) {
    if (k <= 0 || d <= 0 || nq < 0 || nb < 0) return -1;
    int nt = 1;
#ifdef _OPENMP
    nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#endif
    float* tv = (float*)malloc((size_t)nt * (size_t)nq * (size_t)k * sizeof(float));
    int64_t* ti = (int64_t*)malloc((size_t)nt * (size_t)nq * (size_t)k * sizeof(int64_t));
    if (!tv || !ti) {
        free(tv);
        free(ti);
        return -2;
    }
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        int64_t lo = nb * t / nt, hi = nb * (t + 1) / nt;
        for (int64_t i = 0; i < nq; i++) {
            minheap_t h = {k, tv + ((size_t)t * nq + i) * k, ti + ((size_t)t * nq + i) * k};
            heap_fill_neutral(&h);
        }
        /* row-major walk so that every database row is read once per thread for all queries */
        for (int64_t j = lo; j < hi; j++) {
            const float* x = xb + j * d;
            for (int64_t i = 0; i < nq; i++) {
                minheap_t h = {k, tv + ((size_t)t * nq + i) * k, ti + ((size_t)t * nq + i) * k};
                heap_offer(&h, dot_f32_simd(xq + i * d, x, d), j);
            }
        }
    }
    for (int64_t i = 0; i < nq; i++) {
        minheap_t h = {k, D + i * k, I + i * k};
        heap_fill_neutral(&h);
        for (int t = 0; t < nt; t++) {
            const float* v = tv + ((size_t)t * nq + i) * k;
            const int64_t* id = ti + ((size_t)t * nq + i) * k;
            for (int64_t s = 0; s < k; s++)
                if (id[s] >= 0) heap_offer(&h, v[s], id[s]);
        }
        heap_to_sorted(&h);
    }
    free(tv);
    free(ti);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Canonical ranking: CANON-32 fp64 scores, order (score desc, id asc), fp32 score out.
 * D64 (optional, may be NULL) receives the fp64 scores.  Padding: (-FLT_MAX, -1).
 * Rows are split across threads; the per-thread k-lists are merged under the same total order,
 * so the result does not depend on the thread count.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    double s;
    int64_t id;
} cand_t;

static inline int cand_better(double s, int64_t id, const cand_t* b) {
    return (s > b->s) || (s == b->s && id < b->id);
}

/* keep list[0..n) sorted best-first, capacity k */
static inline void sorted_offer(cand_t* list, int64_t* n, int64_t k, double s, int64_t id) {
    if (*n == k && !cand_better(s, id, &list[k - 1])) return;
    int64_t pos = (*n < k) ? (*n)++ : k - 1;
    while (pos > 0 && cand_better(s, id, &list[pos - 1])) {
        list[pos] = list[pos - 1];
        pos--;
    }
    list[pos].s = s;
    list[pos].id = id;
}

ORC_API int orc_canon_search(const float* xq, const float* xb, int64_t d, int64_t nq, int64_t nb, int64_t k,
                             float* D, int64_t* I, double* D64, int64_t id_base, int nthreads) {
    if (k <= 0 || d <= 0 || nq < 0 || nb < 0) return -1;
    int nt = 1;
#ifdef _OPENMP
    nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#endif
    cand_t* lists = (cand_t*)malloc((size_t)nt * (size_t)nq * (size_t)k * sizeof(cand_t));
    int64_t* counts = (int64_t*)calloc((size_t)nt * (size_t)nq, sizeof(int64_t));
    if (!lists || !counts) {
        free(lists);
        free(counts);
        return -2;
    }
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        int64_t lo = nb * t / nt, hi = nb * (t + 1) / nt;
        for (int64_t j = lo; j < hi; j++) {
            const float* x = xb + j * d;
            for (int64_t i = 0; i < nq; i++) {
                double s = dot_canon32(x, xq + i * d, d);
                sorted_offer(lists + ((size_t)t * nq + i) * k, counts + (size_t)t * nq + i, k, s, j + id_base);
            }
        }
    }
    cand_t* fin = (cand_t*)malloc((size_t)k * sizeof(cand_t));
    for (int64_t i = 0; i < nq; i++) {
        int64_t n = 0;
        for (int t = 0; t < nt; t++) {
            const cand_t* l = lists + ((size_t)t * nq + i) * k;
            int64_t c = counts[(size_t)t * nq + i];
            for (int64_t s = 0; s < c; s++) sorted_offer(fin, &n, k, l[s].s, l[s].id);
        }
        for (int64_t s = 0; s < k; s++) {
            if (s < n) {
                D[i * k + s] = (float)fin[s].s;
                I[i * k + s] = fin[s].id;
                if (D64) D64[i * k + s] = fin[s].s;
            } else {
                D[i * k + s] = -FLT_MAX;
                I[i * k + s] = -1;
                if (D64) D64[i * k + s] = -(double)FLT_MAX;
            }
        }
    }
    free(fin);
    free(lists);
    free(counts);
    return 0;
}

/* every row's CANON-32 score for one query (used by near-tie-aware checks) */
ORC_API int orc_canon_scores(const float* q, const float* xb, int64_t d, int64_t nb, double* out, int nthreads) {
    if (d <= 0 || nb < 0) return -1;
#ifdef _OPENMP
    int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (int64_t j = 0; j < nb; j++) out[j] = dot_canon32(xb + j * d, q, d);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * L2 normalise (oldapp.py:35, :43, :51: x /= x.norm(dim=-1, keepdim=True), no epsilon).
 * Defined for parity with the CUDA kernel as: sum of squares accumulated in fp64 in CANON-32
 * order; norm = (float)sqrt(sum) (correctly rounded twice); out = x / norm in fp32 IEEE division.
 * A zero row gives 0/0 = NaN, as the reference does.
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_l2_normalize_f32(float* x, int64_t n, int64_t d, int nthreads) {
    if (d <= 0 || n < 0) return -1;
#ifdef _OPENMP
    int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (int64_t r = 0; r < n; r++) {
        float* row = x + r * d;
        float nrm = (float)sqrt(dot_canon32(row, row, d));
        for (int64_t i = 0; i < d; i++) row[i] = row[i] / nrm;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic embeddings: counter-based, a pure function of
 * (seed, global row, column) in integer arithmetic so that the CPU and the CUDA generator agree
 * bit for bit and shards are reproducible for any GPU count.  One splitmix64 draw gives four
 * 16-bit uniforms whose centred sum (Irwin-Hall n=4, an integer in [-131070, 131070], exact in
 * fp32) stands in for N(0,1); rows are L2-normalised afterwards.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

ORC_API float orc_synth_value(uint64_t seed, int64_t row, int64_t col) {
    uint64_t h = splitmix64(splitmix64(seed ^ 0xD1B54A32D192ED03ull) + (uint64_t)row * 0x2545F4914F6CDD1Dull);
    h = splitmix64(h + (uint64_t)col);
    int32_t s = (int32_t)(h & 0xFFFF) + (int32_t)((h >> 16) & 0xFFFF) + (int32_t)((h >> 32) & 0xFFFF) +
                (int32_t)((h >> 48) & 0xFFFF) - 131070;
    return (float)s;
}

ORC_API int orc_synth_fill(float* out, int64_t n, int64_t d, uint64_t seed, int64_t row_base, int normalize,
                           int nthreads) {
    if (d <= 0 || n < 0) return -1;
#ifdef _OPENMP
    int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (int64_t r = 0; r < n; r++) {
        float* row = out + r * d;
        uint64_t hr = splitmix64(splitmix64(seed ^ 0xD1B54A32D192ED03ull) +
                                 (uint64_t)(r + row_base) * 0x2545F4914F6CDD1Dull);
        for (int64_t c = 0; c < d; c++) {
            uint64_t h = splitmix64(hr + (uint64_t)c);
            int32_t s = (int32_t)(h & 0xFFFF) + (int32_t)((h >> 16) & 0xFFFF) + (int32_t)((h >> 32) & 0xFFFF) +
                        (int32_t)((h >> 48) & 0xFFFF) - 131070;
            row[c] = (float)s;
        }
        if (normalize) {
            float nrm = (float)sqrt(dot_canon32(row, row, d));
            for (int64_t i = 0; i < d; i++) row[i] = row[i] / nrm;
        }
    }
    return 0;
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
