/* CPU oracle, C restatement -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's
 * cpu_baseline / --impl reference legs).  The product library never links this file.
 *
 * Restates the exhaustive form of the distance functions the reference's vector
 * store evaluates: chroma-hnswlib 0.7.3 (pinned through chromadb==0.4.22,
 * /root/reference/requirements.txt:21; call sites /root/reference/app/utils/embedder.py:518,596,901)
 *   L2Sqr                     sum_i (a_i-b_i)^2
 *   InnerProductDistance      1 - sum_i a_i b_i
 *   cosine                    normalise (x * 1/(sqrt(sum x^2)+1e-30)), then InnerProductDistance
 * The packages are absent from /root/reference and this image: parity unpinned against
 * a real Chroma; pinned against the WAL fixture in tests/golden/ (see exact_oracle.py).
 *
 * Two accumulation modes:
 *   acc64=1  fp64 accumulation, order (distance, row) -- the parity oracle
 *   acc64=0  fp32 accumulation (what hnswlib's SIMD kernels do) -- the timed CPU baseline
 * Threads: OpenMP over row blocks, per-thread per-query bounded lists, merged at the end.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { SPACE_L2 = 0, SPACE_COSINE = 1, SPACE_IP = 2 };

typedef struct { double d; int64_t row; } cand_t;

static inline int cand_less(cand_t a, cand_t b) {          /* a ranks before b */
    return a.d < b.d || (a.d == b.d && a.row < b.row);
}

/* sorted bounded insertion list of length <= k (k is small: <= a few hundred) */
static inline void list_push(cand_t *l, int *len, int k, cand_t c) {
    if (*len == k && !cand_less(c, l[k - 1])) return;
    int i = (*len < k) ? (*len)++ : k - 1;
    while (i > 0 && cand_less(c, l[i - 1])) { l[i] = l[i - 1]; --i; }
    l[i] = c;
}

int b2r_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* n <= 0 restores the OpenMP default.  torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm of the bench
 * runs on rank 0 alone and asks for all the cores it may use. */
void b2r_oracle_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
    (void)n;
#endif
}

/* hnswlib bindings.cpp normalize_vector; sum of squares in fp64, rounded once to fp32 */
void b2r_oracle_normalize_f32(const float *x, int64_t n, int d, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        const float *xr = x + r * (int64_t)d;
        double s = 0.0;
        for (int i = 0; i < d; ++i) s += (double)xr[i] * (double)xr[i];
        float inv = 1.0f / (sqrtf((float)s) + 1e-30f);
        for (int i = 0; i < d; ++i) out[r * (int64_t)d + i] = xr[i] * inv;
    }
}

static inline double dist64(const float *q, const float *x, int d, int space) {
    double s = 0.0;
    if (space == SPACE_L2) {
        for (int i = 0; i < d; ++i) { double t = (double)q[i] - (double)x[i]; s += t * t; }
        return s;
    }
    for (int i = 0; i < d; ++i) s += (double)q[i] * (double)x[i];
    return 1.0 - s;
}

static inline float dist32(const float *q, const float *x, int d, int space) {
    float s = 0.0f;
    if (space == SPACE_L2) {
#pragma omp simd reduction(+ : s)
        for (int i = 0; i < d; ++i) { float t = q[i] - x[i]; s += t * t; }
        return s;
    }
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < d; ++i) s += q[i] * x[i];
    return 1.0f - s;
}

/* X [n,d] stored rows (already normalised for cosine), Q [nq,d] (already normalised for
 * cosine), allowed: NULL or n bytes (0 = skip row).  Outputs are [nq,k], padded with
 * row=-1 / dist=+inf; out_count[nq] = valid entries.  Returns 0, or -1 on bad args. */
int b2r_oracle_topk(const float *X, int64_t n, int d, const float *Q, int nq, int k, int space,
                    const uint8_t *allowed, int acc64, int64_t *out_rows, float *out_dist,
                    int32_t *out_count) {
    if (n < 0 || d <= 0 || nq <= 0 || k <= 0 || space < 0 || space > 2) return -1;
    int nt = b2r_oracle_threads();
    cand_t *lists = (cand_t *)malloc((size_t)nt * nq * k * sizeof(cand_t));
    int *lens = (int *)calloc((size_t)nt * nq, sizeof(int));
    if (!lists || !lens) { free(lists); free(lens); return -1; }
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        cand_t *my = lists + (size_t)t * nq * k;
        int *mylen = lens + (size_t)t * nq;
#pragma omp for schedule(dynamic, 1024)
        for (int64_t r = 0; r < n; ++r) {
            if (allowed && !allowed[r]) continue;
            const float *xr = X + r * (int64_t)d;
            for (int j = 0; j < nq; ++j) {               /* row stays in L1 across the batch */
                cand_t c;
                c.d = acc64 ? dist64(Q + (int64_t)j * d, xr, d, space)
                            : (double)dist32(Q + (int64_t)j * d, xr, d, space);
                c.row = r;
                list_push(my + (size_t)j * k, &mylen[j], k, c);
            }
        }
    }
    for (int j = 0; j < nq; ++j) {
        cand_t *dst = lists + (size_t)j * k;             /* thread 0's list is the merge target */
        int len = lens[j];
        for (int t = 1; t < nt; ++t) {
            cand_t *src = lists + ((size_t)t * nq + j) * k;
            for (int i = 0; i < lens[(size_t)t * nq + j]; ++i) list_push(dst, &len, k, src[i]);
        }
        out_count[j] = len;
        for (int i = 0; i < k; ++i) {
            out_rows[(int64_t)j * k + i] = i < len ? dst[i].row : -1;
            out_dist[(int64_t)j * k + i] = i < len ? (float)dst[i].d : INFINITY;
        }
    }
    free(lists); free(lens);
    return 0;
}
