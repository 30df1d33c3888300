/* CPU baseline, C restatement of the index the reference actually queries -- TEST / BENCH INFRASTRUCTURE ONLY
 * (bench.py's cpu_baseline and --impl reference legs, tests/).  The product library never links this file.
 *
 * The reference's collection.query (/root/reference/app/utils/embedder.py:595-601) ends in
 * chroma-hnswlib 0.7.3 `Index.knn_query` (pinned through chromadb==0.4.22, requirements.txt:21; not vendored, not
 * installable here).  This file restates the published HNSW algorithm as hnswlib implements it, with Chroma's
 * defaults, which the reference's committed index header also records (chroma_db/.../header.bin: M=16,
 * ef_construction=100, maxM0=32, mult=1/ln 16):
 *   - level = floor(-ln(U) / ln(M)); greedy descent through the upper layers; ef-bounded best-first search on a
 *     layer (searchBaseLayer); neighbour selection by the distance heuristic (getNeighborsByHeuristic2: keep a
 *     candidate only if it is closer to the query than to every neighbour already kept); back-links pruned with
 *     the same heuristic when a list overflows (mutuallyConnectNewElement).
 *   - query: ef = max(search_ef = 10, k) -- Chroma's default `hnsw:search_ef` -- results ascending by distance.
 *   - distances: `cosine` = 1 - dot on vectors normalised beforehand (hnswlib normalises at add/query time);
 *     `l2` = squared L2.  fp32 accumulation, as hnswlib's SIMD kernels.
 * It is approximate by construction: bench.py reports its recall@k against the exact oracle beside its speed.
 * Build and search are OpenMP-parallel like hnswlib's ParallelFor (per-node locks while linking).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int n, d, M, M0, efc, space;      /* space: 0 = l2, 1 = cosine / ip (1 - dot) */
    const float *X;                   /* [n, d], caller-owned                     */
    int *level;                       /* [n]                                      */
    int *link0;                       /* [n][M0 + 1]: count, ids                  */
    int **linkU;                      /* [n] -> [level][M + 1] or NULL            */
    int maxlevel, entry;
#ifdef _OPENMP
    omp_lock_t *locks;                /* [n]                                      */
    omp_lock_t glock;
#endif
} hnsw_t;

typedef struct { float d; int id; } cand_t;

static inline float dist_f(const hnsw_t *h, const float *a, const float *b) {
    float s = 0.f;
    const int d = h->d;
    if (h->space == 0) {
        for (int i = 0; i < d; ++i) { float t = a[i] - b[i]; s += t * t; }
        return s;
    }
    for (int i = 0; i < d; ++i) s += a[i] * b[i];
    return 1.0f - s;
}

/* ---- binary heaps on cand_t: max-heap (far = 1) keeps the worst on top, min-heap the best ---- */
typedef struct { cand_t *a; int n, cap, far; } heap_t;
static inline int hless(const heap_t *h, cand_t x, cand_t y) { return h->far ? x.d > y.d : x.d < y.d; }
static void hpush(heap_t *h, cand_t c) {
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 64; h->a = (cand_t *)realloc(h->a, sizeof(cand_t) * h->cap); }
    int i = h->n++;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!hless(h, c, h->a[p])) break;
        h->a[i] = h->a[p]; i = p;
    }
    h->a[i] = c;
}
static cand_t hpop(heap_t *h) {
    cand_t top = h->a[0], last = h->a[--h->n];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        cand_t best = last;
        if (l < h->n && hless(h, h->a[l], best)) { m = l; best = h->a[l]; }
        if (r < h->n && hless(h, h->a[r], best)) { m = r; best = h->a[r]; }
        if (m == i) break;
        h->a[i] = h->a[m]; i = m;
    }
    if (h->n) h->a[i] = last;
    return top;
}

static inline int *links_of(const hnsw_t *h, int id, int lev) {
    return lev == 0 ? h->link0 + (size_t)id * (h->M0 + 1) : h->linkU[id] + (size_t)(lev - 1) * (h->M + 1);
}

typedef struct { unsigned *stamp; unsigned epoch; heap_t cand, res; int *tmp; cand_t *sel; } scratch_t;

/* ef-bounded best-first search on one layer; results left in s->res (max-heap, <= ef entries) */
static void search_layer(hnsw_t *h, const float *q, int ep, float epd, int ef, int lev, scratch_t *s, int locking) {
    if (++s->epoch == 0) { memset(s->stamp, 0, sizeof(unsigned) * h->n); s->epoch = 1; }
    s->cand.n = 0; s->res.n = 0;
    cand_t c0 = {epd, ep};
    hpush(&s->cand, c0); hpush(&s->res, c0);
    s->stamp[ep] = s->epoch;
    while (s->cand.n) {
        cand_t c = s->cand.a[0];
        if (c.d > s->res.a[0].d && s->res.n >= ef) break;
        hpop(&s->cand);
        int *l = links_of(h, c.id, lev), cnt;
#ifdef _OPENMP
        if (locking) omp_set_lock(&h->locks[c.id]);
#endif
        cnt = l[0];
        memcpy(s->tmp, l + 1, sizeof(int) * cnt);
#ifdef _OPENMP
        if (locking) omp_unset_lock(&h->locks[c.id]);
#endif
        for (int i = 0; i < cnt; ++i) {
            int nb = s->tmp[i];
            if (s->stamp[nb] == s->epoch) continue;
            s->stamp[nb] = s->epoch;
            float d = dist_f(h, q, h->X + (size_t)nb * h->d);
            if (s->res.n < ef || d < s->res.a[0].d) {
                cand_t e = {d, nb};
                hpush(&s->cand, e); hpush(&s->res, e);
                if (s->res.n > ef) hpop(&s->res);
            }
        }
    }
}

static int cmp_cand(const void *a, const void *b) {
    float x = ((const cand_t *)a)->d, y = ((const cand_t *)b)->d;
    return x < y ? -1 : x > y ? 1 : 0;
}

/* hnswlib getNeighborsByHeuristic2: in[] ascending by distance to the base point; returns kept count */
static int select_heuristic(const hnsw_t *h, cand_t *in, int nin, int M, cand_t *out) {
    int nout = 0;
    for (int i = 0; i < nin && nout < M; ++i) {
        int good = 1;
        const float *xi = h->X + (size_t)in[i].id * h->d;
        for (int j = 0; j < nout; ++j)
            if (dist_f(h, xi, h->X + (size_t)out[j].id * h->d) < in[i].d) { good = 0; break; }
        if (good) out[nout++] = in[i];
    }
    return nout;
}

static void scratch_init(scratch_t *s, int n, int cap) {
    memset(s, 0, sizeof(*s));
    s->stamp = (unsigned *)calloc(n, sizeof(unsigned));
    s->cand.far = 0; s->res.far = 1;
    s->tmp = (int *)malloc(sizeof(int) * (cap + 1));
    s->sel = (cand_t *)malloc(sizeof(cand_t) * (cap + 2) * 2);
}
static void scratch_free(scratch_t *s) { free(s->stamp); free(s->cand.a); free(s->res.a); free(s->tmp); free(s->sel); }

static void insert(hnsw_t *h, int id, scratch_t *s) {
    const float *q = h->X + (size_t)id * h->d;
    const int lev = h->level[id];
#ifdef _OPENMP
    omp_set_lock(&h->glock);
#endif
    int ep = h->entry, maxl = h->maxlevel;
    if (ep < 0) { h->entry = id; h->maxlevel = lev; }
    const int hold_global = ep >= 0 && lev > maxl;     /* hnswlib keeps the global lock while raising the top level */
#ifdef _OPENMP
    if (!hold_global) omp_unset_lock(&h->glock);
#endif
    if (ep < 0) return;
    float epd = dist_f(h, q, h->X + (size_t)ep * h->d);
    for (int l = maxl; l > lev; --l) {                  /* greedy descent */
        int changed = 1;
        while (changed) {
            changed = 0;
            int *lk = links_of(h, ep, l), cnt;
#ifdef _OPENMP
            omp_set_lock(&h->locks[ep]);
#endif
            cnt = lk[0];
            memcpy(s->tmp, lk + 1, sizeof(int) * cnt);
#ifdef _OPENMP
            omp_unset_lock(&h->locks[ep]);
#endif
            for (int i = 0; i < cnt; ++i) {
                float d = dist_f(h, q, h->X + (size_t)s->tmp[i] * h->d);
                if (d < epd) { epd = d; ep = s->tmp[i]; changed = 1; }
            }
        }
    }
    for (int l = lev < maxl ? lev : maxl; l >= 0; --l) {
        search_layer(h, q, ep, epd, h->efc, l, s, 1);
        int nres = s->res.n;
        cand_t *sorted = s->sel, *kept = s->sel + nres + 1;
        memcpy(sorted, s->res.a, sizeof(cand_t) * nres);
        qsort(sorted, nres, sizeof(cand_t), cmp_cand);
        const int Mmax = l == 0 ? h->M0 : h->M;
        int nk = select_heuristic(h, sorted, nres, h->M, kept);
        ep = sorted[0].id; epd = sorted[0].d;            /* closest found: entry point for the next layer */
        int *mine = links_of(h, id, l);
#ifdef _OPENMP
        omp_set_lock(&h->locks[id]);
#endif
        mine[0] = nk;
        for (int i = 0; i < nk; ++i) mine[1 + i] = kept[i].id;
#ifdef _OPENMP
        omp_unset_lock(&h->locks[id]);
#endif
        for (int i = 0; i < nk; ++i) {                   /* back-links, pruned with the heuristic on overflow */
            const int nb = kept[i].id;
            int *lk = links_of(h, nb, l);
#ifdef _OPENMP
            omp_set_lock(&h->locks[nb]);
#endif
            if (lk[0] < Mmax) {
                lk[1 + lk[0]++] = id;
            } else {
                cand_t buf[66], out[66];
                const float *xn = h->X + (size_t)nb * h->d;
                int m = 0;
                buf[m].d = kept[i].d; buf[m++].id = id;
                for (int j = 0; j < lk[0]; ++j) { buf[m].id = lk[1 + j]; buf[m].d = dist_f(h, xn, h->X + (size_t)lk[1 + j] * h->d); ++m; }
                qsort(buf, m, sizeof(cand_t), cmp_cand);
                int no = select_heuristic(h, buf, m, Mmax, out);
                lk[0] = no;
                for (int j = 0; j < no; ++j) lk[1 + j] = out[j].id;
            }
#ifdef _OPENMP
            omp_unset_lock(&h->locks[nb]);
#endif
        }
    }
    if (hold_global) {
        h->entry = id; h->maxlevel = lev;
#ifdef _OPENMP
        omp_unset_lock(&h->glock);
#endif
    }
}

/* ---------------------------------------------------------------------------------------------- API */
int b2r_hnsw_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* n <= 0 restores the default; a single-threaded build is deterministic (tests) */
void b2r_hnsw_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
    (void)n;
#endif
}

void *b2r_hnsw_build(const float *X, int n, int d, int space, int M, int ef_construction, unsigned seed) {
    if (M > 32) return NULL;
    hnsw_t *h = (hnsw_t *)calloc(1, sizeof(hnsw_t));
    h->n = n; h->d = d; h->M = M; h->M0 = 2 * M; h->efc = ef_construction; h->space = space; h->X = X;
    h->entry = -1; h->maxlevel = -1;
    h->level = (int *)malloc(sizeof(int) * n);
    h->link0 = (int *)calloc((size_t)n * (h->M0 + 1), sizeof(int));
    h->linkU = (int **)calloc(n, sizeof(int *));
    const double mult = 1.0 / log((double)M);
    uint64_t st = seed ? seed : 100;                   /* hnswlib's default random_seed is 100 */
    for (int i = 0; i < n; ++i) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        double u = ((st >> 11) + 1.0) / 9007199254740993.0;
        int l = (int)(-log(u) * mult);
        h->level[i] = l;
        if (l > 0) h->linkU[i] = (int *)calloc((size_t)l * (M + 1), sizeof(int));
    }
#ifdef _OPENMP
    h->locks = (omp_lock_t *)malloc(sizeof(omp_lock_t) * n);
    for (int i = 0; i < n; ++i) omp_init_lock(&h->locks[i]);
    omp_init_lock(&h->glock);
#pragma omp parallel
#endif
    {
        scratch_t s;
        scratch_init(&s, n, h->M0 > h->efc ? h->M0 : h->efc);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (int i = 0; i < n; ++i) insert(h, i, &s);
        scratch_free(&s);
    }
    return h;
}

/* Q [nq, d] (normalised by the caller for cosine) -> out_rows [nq, k] (-1 pad), out_dist [nq, k] ascending */
void b2r_hnsw_search(void *hv, const float *Q, int nq, int k, int ef, int64_t *out_rows, float *out_dist) {
    hnsw_t *h = (hnsw_t *)hv;
    if (ef < k) ef = k;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        scratch_t s;
        scratch_init(&s, h->n, h->M0 > ef ? h->M0 : ef);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int qi = 0; qi < nq; ++qi) {
            const float *q = Q + (size_t)qi * h->d;
            int ep = h->entry;
            float epd = dist_f(h, q, h->X + (size_t)ep * h->d);
            for (int l = h->maxlevel; l > 0; --l) {
                int changed = 1;
                while (changed) {
                    changed = 0;
                    int *lk = links_of(h, ep, l);
                    for (int i = 0; i < lk[0]; ++i) {
                        float d = dist_f(h, q, h->X + (size_t)lk[1 + i] * h->d);
                        if (d < epd) { epd = d; ep = lk[1 + i]; changed = 1; }
                    }
                }
            }
            search_layer(h, q, ep, epd, ef, 0, &s, 0);
            while (s.res.n > k) hpop(&s.res);
            int m = s.res.n;
            for (int i = m; i < k; ++i) { out_rows[(size_t)qi * k + i] = -1; out_dist[(size_t)qi * k + i] = INFINITY; }
            for (int i = m - 1; i >= 0; --i) {
                cand_t c = hpop(&s.res);
                out_rows[(size_t)qi * k + i] = c.id; out_dist[(size_t)qi * k + i] = c.d;
            }
        }
        scratch_free(&s);
    }
}

void b2r_hnsw_free(void *hv) {
    hnsw_t *h = (hnsw_t *)hv;
    if (!h) return;
    for (int i = 0; i < h->n; ++i) free(h->linkU[i]);
#ifdef _OPENMP
    for (int i = 0; i < h->n; ++i) omp_destroy_lock(&h->locks[i]);
    omp_destroy_lock(&h->glock);
    free(h->locks);
#endif
    free(h->linkU); free(h->link0); free(h->level); free(h);
}
