// C ABI implementation (include/b2r.h): handle lifetime, HBM layout, host<->device
// staging and kernel dispatch.  No compute happens on the host.
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "engine.h"
#include "gemm.cuh"
#include "ingest.cuh"

using namespace b2r;

constexpr int GEMM_SEED_MIN_BATCH = 1;    // K3 in-kernel seeding pays from batch 1 on (smaller pools, no unseeded slow path)

// ---------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------
static thread_local std::string g_last_error;
void b2r::set_error(const std::string &msg) { g_last_error = msg; }

#define B2R_REQUIRE(cond, msg)                      \
    do {                                            \
        if (!(cond)) { set_error(msg); return B2R_EINVAL; } \
    } while (0)

extern "C" const char *b2r_last_error(void) { return g_last_error.c_str(); }
extern "C" int b2r_abi_version(void) { return B2R_ABI_VERSION; }

// ---------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------
namespace {

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int ensure(DevBuf &b, size_t bytes) {
    if (b.bytes >= bytes) return B2R_OK;
    if (b.p) { B2R_CUDA(cudaFree(b.p)); b.p = nullptr; b.bytes = 0; }
    size_t want = std::max(bytes, (size_t)256);
    want = (want + 255) & ~(size_t)255;
    B2R_CUDA(cudaMalloc(&b.p, want));
    b.bytes = want;
    return B2R_OK;
}
void release(DevBuf &b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.bytes = 0; }

int round_up(int v, int m) { return (v + m - 1) / m * m; }

int ingest_ctas_per_sm(ingest_fn fn) {
    static std::mutex mu;
    static std::map<const void *, int> cache;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find((const void *)fn);
    if (it != cache.end()) return it->second;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, INGEST_THREADS, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 4;
    }
    cache[(const void *)fn] = per_sm;
    return per_sm;
}

// entries-per-lane of the candidate lists for a fast-path query: KP = 32*EPL >= 2k
int epl_scored(int k) { return k <= 16 ? 1 : k <= 32 ? 2 : k <= 64 ? 4 : k <= 128 ? 8 : 0; }
// exact path: KP >= k
int epl_exact(int k) { return k <= 32 ? 1 : k <= 64 ? 2 : k <= 128 ? 4 : k <= 256 ? 8 : 0; }

size_t row_bytes_total(const b2r_index *h) {
    size_t b = (size_t)h->dp * 2 + 1;
    if (h->master) b += (size_t)h->dp * 4;
    if (h->space == B2R_SPACE_L2) b += 4;
    return b;
}

// (re)allocate the corpus arrays for `cap` rows, preserving the first h->rows rows
int grow(b2r_index *h, int64_t cap, cudaStream_t s) {
    if (cap <= h->capacity) return B2R_OK;
    cap = std::max<int64_t>(cap, 1024);
    uint4 *corpus = nullptr; float *master = nullptr, *bias = nullptr; uint8_t *tc = nullptr;
    const size_t cb = (size_t)cap * h->dp * 2;
    B2R_CUDA(cudaMalloc(&corpus, cb));
    if (!(h->flags & B2R_FLAG_NO_F32_MASTER)) {
        cudaError_t e = cudaMalloc(&master, (size_t)cap * h->dp * 4);
        if (e != cudaSuccess) { cudaFree(corpus); B2R_CUDA(e); }
    }
    if (h->space == B2R_SPACE_L2) {
        cudaError_t e = cudaMalloc(&bias, ((size_t)cap + 512) * 4);   // K3 reads whole tiles
        if (e != cudaSuccess) { cudaFree(corpus); cudaFree(master); B2R_CUDA(e); }
    }
    {
        cudaError_t e = cudaMalloc(&tc, (size_t)cap + 4);
        if (e != cudaSuccess) { cudaFree(corpus); cudaFree(master); cudaFree(bias); B2R_CUDA(e); }
    }
    if (h->rows > 0) {
        B2R_CUDA(cudaMemcpyAsync(corpus, h->corpus, (size_t)h->rows * h->dp * 2, cudaMemcpyDeviceToDevice, s));
        if (master) B2R_CUDA(cudaMemcpyAsync(master, h->master, (size_t)h->rows * h->dp * 4, cudaMemcpyDeviceToDevice, s));
        if (bias) B2R_CUDA(cudaMemcpyAsync(bias, h->bias, (size_t)h->rows * 4, cudaMemcpyDeviceToDevice, s));
        B2R_CUDA(cudaMemcpyAsync(tc, h->type_code, (size_t)h->rows, cudaMemcpyDeviceToDevice, s));
    }
    int32_t *cols[B2R_MAX_COLUMNS] = {};
    for (int c = 0; c < B2R_MAX_COLUMNS; ++c) {
        if (!h->cols[c]) continue;
        cudaError_t e = cudaMalloc(&cols[c], (size_t)cap * 4);
        if (e != cudaSuccess) {
            for (int j = 0; j < c; ++j) cudaFree(cols[j]);
            cudaFree(corpus); cudaFree(master); cudaFree(bias); cudaFree(tc);
            B2R_CUDA(e);
        }
        cudaMemsetAsync(cols[c], 0xff, (size_t)cap * 4, s);
        if (h->rows > 0) cudaMemcpyAsync(cols[c], h->cols[c], (size_t)h->rows * 4, cudaMemcpyDeviceToDevice, s);
    }
    B2R_CUDA(cudaStreamSynchronize(s));
    cudaFree(h->corpus); cudaFree(h->master); cudaFree(h->bias); cudaFree(h->type_code);
    for (int c = 0; c < B2R_MAX_COLUMNS; ++c) { cudaFree(h->cols[c]); h->cols[c] = cols[c]; }
    h->corpus = corpus; h->master = master; h->bias = bias; h->type_code = tc;
    h->capacity = cap;
    return B2R_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------
extern "C" int b2r_create(int dim, int space, int64_t capacity_rows, int device, uint32_t flags,
                          b2r_handle *out) {
    B2R_REQUIRE(out != nullptr, "b2r_create: out is NULL");
    *out = nullptr;
    B2R_REQUIRE(dim >= 1 && dim <= 8192, "b2r_create: dim must be in [1, 8192]");
    B2R_REQUIRE(space >= 0 && space <= 2, "b2r_create: unknown space");
    B2R_REQUIRE(capacity_rows >= 0 && capacity_rows < (1ll << 32) - 64, "b2r_create: capacity out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("b2r_create: no CUDA device (this engine has no CPU fallback)");
        return B2R_ECUDA;
    }
    B2R_REQUIRE(device >= 0 && device < ndev, "b2r_create: bad device ordinal");
    B2R_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2R_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("b2r_create: kernels are built for sm_100a only, device is sm_" + std::to_string(prop.major) +
                  std::to_string(prop.minor));
        return B2R_EUNSUPPORTED;
    }
    b2r_index *h = new (std::nothrow) b2r_index();
    if (!h) { set_error("b2r_create: host allocation failed"); return B2R_ENOMEM; }
    h->dim = dim; h->dp = round_up(dim, 64); h->space = space; h->device = device; h->flags = flags;
    h->sm_count = prop.multiProcessorCount;
    if (const char *e = getenv("B2R_TIME_STAGE")) h->timing_stage = atoi(e);
    if (const char *e = getenv("B2R_NO_SEED")) h->no_seed = atoi(e) != 0;      // development: K3 without in-kernel seeding
    h->seed_min_batch = GEMM_SEED_MIN_BATCH;
    if (const char *e = getenv("B2R_SEED_MIN_BATCH")) h->seed_min_batch = atoi(e);
    if (const char *e = getenv("B2R_SEED_TILES")) h->seed_tiles_override = atoi(e);   // development: seeding tiles per CTA
    if (const char *e = getenv("B2R_SEED_WAIT_NS")) h->seed_wait_ns = strtoull(e, nullptr, 10);
    if (const char *e = getenv("B2R_DELAY_US")) h->delay_us = atoi(e);
    if (const char *e = getenv("B2R_POOL_SAMPLE_DIV")) h->pool_sample_div = std::max(0, atoi(e));
    if (const char *e = getenv("B2R_TRACE")) { h->trace_on = atoi(e) != 0; h->trace_mode = atoi(e); }
    if (const char *e = getenv("B2R_NO_PAIR")) h->no_pair = atoi(e) != 0;
    if (const char *e = getenv("B2R_SEED_RANK_L")) h->seed_rank_l = atoi(e) != 0;
    if (const char *e = getenv("B2R_NO_DYN")) h->no_dyn = atoi(e) != 0;
    if (const char *e = getenv("B2R_NO_BM64")) h->no_bm64 = atoi(e) != 0;
    int rc = B2R_OK;
    do {
        if (cudaMalloc(&h->max_norm2, 256) != cudaSuccess || cudaMalloc(&h->counters, 256) != cudaSuccess ||
            cudaMalloc(&h->tickets, sizeof(unsigned) * (1 + EXACT_MAX_SLOTS)) != cudaSuccess ||
            cudaMalloc(&h->need_ctl, 16) != cudaSuccess) {
            set_error("b2r_create: cudaMalloc failed"); rc = B2R_ENOMEM; break;
        }
        cudaMemset(h->max_norm2, 0, 256);
        cudaMemset(h->counters, 0, 256);
        cudaMemset(h->tickets, 0, sizeof(unsigned) * (1 + EXACT_MAX_SLOTS));
        cudaMemset(h->need_ctl, 0, 16);
        rc = grow(h, std::max<int64_t>(capacity_rows, 1024), 0);
    } while (0);
    if (rc != B2R_OK) { b2r_destroy(h); return rc; }
    *out = h;
    return B2R_OK;
}

extern "C" int b2r_destroy(b2r_handle h) {
    if (!h) return B2R_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->corpus); cudaFree(h->master); cudaFree(h->bias); cudaFree(h->type_code);
    cudaFree(h->max_norm2); cudaFree(h->counters); cudaFree(h->tickets); cudaFree(h->need_ctl);
    for (int c = 0; c < B2R_MAX_COLUMNS; ++c) cudaFree(h->cols[c]);
    DevBuf *bufs[] = {&h->x_stage, &h->t_stage, &h->q_raw, &h->q_prep, &h->allow, &h->rows_stage, &h->gather_out,
                      &h->o_pack, &h->need_list, &h->scan_lists,
                      &h->exact_lists, &h->q_bf16, &h->q_err, &h->pass_bits, &h->gthr, &h->gemm_lists, &h->gemm_regions,
                      &h->gemm_samples, &h->q_eps, &h->where_lut, &h->where_bits, &h->col_stage, &h->trace};
    for (DevBuf *b : bufs) release(*b);
    if (h->o_host) cudaFreeHost(h->o_host);
    for (auto &sl : h->aslot) {
        release(sl.q_dev); release(sl.o_dev);
        if (sl.in_host) cudaFreeHost(sl.in_host);
        if (sl.out_host) cudaFreeHost(sl.out_host);
        if (sl.ev_h2d) { cudaEventDestroy(sl.ev_h2d); cudaEventDestroy(sl.ev_kernels); cudaEventDestroy(sl.ev_d2h); }
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->copy_stream_out) cudaStreamDestroy(h->copy_stream_out);
    for (auto &ev : h->ev_pending) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (auto &ev : h->ev_free) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    delete h;
    return B2R_OK;
}

extern "C" int b2r_clear(b2r_handle h) {
    B2R_REQUIRE(h, "b2r_clear: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    B2R_CUDA(cudaSetDevice(h->device));
    B2R_CUDA(cudaDeviceSynchronize());
    h->rows = 0; h->live = 0; h->mut_gen++;
    B2R_CUDA(cudaMemset(h->max_norm2, 0, 8));
    for (int c = 0; c < B2R_MAX_COLUMNS; ++c)
        if (h->cols[c]) B2R_CUDA(cudaMemset(h->cols[c], 0xff, (size_t)h->capacity * 4));
    return B2R_OK;
}

extern "C" int b2r_reserve(b2r_handle h, int64_t capacity_rows) {
    B2R_REQUIRE(h, "b2r_reserve: NULL handle");
    B2R_REQUIRE(capacity_rows < (1ll << 32) - 64, "b2r_reserve: capacity out of range");
    std::lock_guard<std::mutex> g(h->mu);
    B2R_CUDA(cudaSetDevice(h->device));
    return grow(h, capacity_rows, 0);
}

extern "C" int b2r_set_row_base(b2r_handle h, int64_t row_base) {
    B2R_REQUIRE(h, "b2r_set_row_base: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    h->row_base = row_base;
    return B2R_OK;
}

extern "C" int b2r_set_path(b2r_handle h, int path) {
    B2R_REQUIRE(h, "b2r_set_path: NULL handle");
    B2R_REQUIRE(path >= 0 && path <= 3, "b2r_set_path: path must be 0..3");
    std::lock_guard<std::mutex> g(h->mu);
    h->path = path;
    return B2R_OK;
}

extern "C" int64_t b2r_launch_count(b2r_handle h) { return h ? h->n_launches : 0; }

extern "C" int b2r_set_kernel_timing(b2r_handle h, int enable) {
    B2R_REQUIRE(h, "b2r_set_kernel_timing: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    h->timing = enable != 0;
    return B2R_OK;
}

extern "C" int b2r_kernel_time_ms(b2r_handle h, double *total_ms, int64_t *launches, int reset) {
    B2R_REQUIRE(h && total_ms && launches, "b2r_kernel_time_ms: NULL argument");
    std::lock_guard<std::mutex> g(h->mu);
    B2R_CUDA(cudaSetDevice(h->device));
    for (auto &ev : h->ev_pending) {
        B2R_CUDA(cudaEventSynchronize(ev.second));
        float ms = 0.f;
        B2R_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
        h->scoring_ms += ms; h->scoring_launches++;
        h->ev_free.push_back(ev);
    }
    h->ev_pending.clear();
    *total_ms = h->scoring_ms; *launches = h->scoring_launches;
    if (reset) { h->scoring_ms = 0.0; h->scoring_launches = 0; }
    return B2R_OK;
}
extern "C" int b2r_debug_trace(b2r_handle h, uint64_t *out, int max_ctas, int *n_ctas) {
    B2R_REQUIRE(h && out && n_ctas, "b2r_debug_trace: NULL argument");
    std::lock_guard<std::mutex> g(h->mu);
    *n_ctas = 0;
    if (!h->trace_on || !h->trace.p || h->trace_ctas <= 0) return B2R_OK;
    B2R_CUDA(cudaSetDevice(h->device));
    B2R_CUDA(cudaDeviceSynchronize());
    const int n = std::min(max_ctas, h->trace_ctas);
    B2R_CUDA(cudaMemcpy(out, h->trace.p, sizeof(uint64_t) * 8 * (size_t)n, cudaMemcpyDeviceToHost));
    *n_ctas = n;
    return B2R_OK;
}
extern "C" int64_t b2r_count(b2r_handle h) { return h ? h->live : -1; }

extern "C" int b2r_get_stats(b2r_handle h, b2r_stats *out) {
    B2R_REQUIRE(h && out, "b2r_get_stats: NULL argument");
    std::lock_guard<std::mutex> g(h->mu);
    B2R_CUDA(cudaSetDevice(h->device));
    unsigned long long c[4] = {0, 0, 0, 0};
    B2R_CUDA(cudaMemcpy(c, h->counters, sizeof(c), cudaMemcpyDeviceToHost));
    std::memset(out, 0, sizeof(*out));
    out->dim = h->dim; out->dim_padded = h->dp; out->space = h->space; out->flags = h->flags;
    out->rows = h->rows; out->live = h->live; out->capacity = h->capacity;
    out->bytes_device = (int64_t)(row_bytes_total(h) * (size_t)h->capacity);
    out->n_queries = h->n_queries; out->n_exact_fallbacks = (int64_t)c[1];
    out->sm_count = h->sm_count; out->device = h->device;
    out->n_pool_queries = (int64_t)c[2]; out->n_pool_entries = (int64_t)c[3];
    return B2R_OK;
}

// ---------------------------------------------------------------------------------
// ingest / tombstone / gather
// ---------------------------------------------------------------------------------
extern "C" int b2r_ingest_f32(b2r_handle h, const float *x, int64_t n, const uint8_t *type_code,
                              int64_t *first_row_out, void *stream) {
    B2R_REQUIRE(h, "b2r_ingest_f32: NULL handle");
    B2R_REQUIRE(n >= 0, "b2r_ingest_f32: negative row count");
    B2R_REQUIRE(n == 0 || x, "b2r_ingest_f32: x is NULL");
    std::lock_guard<std::mutex> g(h->mu);
    if (first_row_out) *first_row_out = h->rows;
    if (n == 0) return B2R_OK;
    B2R_REQUIRE(h->rows + n < (1ll << 32) - 64, "b2r_ingest_f32: shard row limit (2^32) exceeded");
    B2R_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (h->rows + n > h->capacity) {
        int rc = grow(h, std::max(h->rows + n, h->capacity * 2), s);
        if (rc != B2R_OK) return rc;
    }
    const float *xd = x;
    if (!is_device_ptr(x)) {
        int rc = ensure(h->x_stage, (size_t)n * h->dim * 4);
        if (rc != B2R_OK) return rc;
        B2R_CUDA(cudaMemcpyAsync(h->x_stage.p, x, (size_t)n * h->dim * 4, cudaMemcpyHostToDevice, s));
        xd = (const float *)h->x_stage.p;
    }
    const uint8_t *td = type_code;
    if (type_code && !is_device_ptr(type_code)) {
        for (int64_t i = 0; i < n; ++i) B2R_REQUIRE(type_code[i] < B2R_TYPE_DEAD, "b2r_ingest_f32: type_code must be 0..62");
        int rc = ensure(h->t_stage, (size_t)n);
        if (rc != B2R_OK) return rc;
        B2R_CUDA(cudaMemcpyAsync(h->t_stage.p, type_code, (size_t)n, cudaMemcpyHostToDevice, s));
        td = (const uint8_t *)h->t_stage.p;
    }
    IngestParams p;
    p.x = xd; p.n = n; p.d = h->dim; p.dp = h->dp; p.space = h->space;
    p.corpus = h->corpus + (size_t)h->rows * (h->dp / 8);
    p.master = h->master ? h->master + (size_t)h->rows * h->dp : nullptr;
    p.bias = h->bias ? h->bias + h->rows : nullptr;
    p.type_out = h->type_code + h->rows;
    p.type_in = td;
    p.max_norm2 = h->max_norm2; p.qerr = nullptr; p.q_eps = nullptr; p.norms = nullptr; p.eps_rel = 0.f;
    for (int z = 0; z < 3; ++z) { p.zero[z] = nullptr; p.zero_words[z] = 0; }
    p.wait_words = nullptr; p.wait_n = 0; p.wait_val = 0;
    std::memset(&p.flags, 0, sizeof p.flags);
    const int wpb = INGEST_THREADS / 32;
    const ingest_fn fn = ingest_lookup(h->dim, h->dp, xd);
    // one full wave of resident CTAs, every warp walks its share of the rows (no tail wave)
    int grid = (int)std::min<int64_t>((n + wpb - 1) / wpb, (int64_t)h->sm_count * ingest_ctas_per_sm(fn));
    fn<<<grid, INGEST_THREADS, 0, s>>>(p);
    B2R_CUDA(cudaGetLastError());
    h->n_launches++;
    if (xd != x || td != type_code) B2R_CUDA(cudaStreamSynchronize(s));   // staging buffers are reused
    h->rows += n; h->live += n; h->mut_gen++;
    return B2R_OK;
}

extern "C" int b2r_tombstone(b2r_handle h, const int64_t *rows, int64_t n, void *stream) {
    B2R_REQUIRE(h, "b2r_tombstone: NULL handle");
    B2R_REQUIRE(n >= 0 && (n == 0 || rows), "b2r_tombstone: bad arguments");
    if (n == 0) return B2R_OK;
    std::lock_guard<std::mutex> g(h->mu);
    B2R_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    for (int64_t i = 0; i < n; ++i)
        B2R_REQUIRE(rows[i] >= 0 && rows[i] < h->rows, "b2r_tombstone: row out of range");
    int rc = ensure(h->rows_stage, (size_t)n * 8);
    if (rc != B2R_OK) return rc;
    unsigned long long before = 0, after = 0;
    B2R_CUDA(cudaMemcpyAsync(&before, h->counters, 8, cudaMemcpyDeviceToHost, s));
    B2R_CUDA(cudaMemcpyAsync(h->rows_stage.p, rows, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    tombstone_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->type_code, (const long long *)h->rows_stage.p, n,
                                                                 h->rows, h->counters);
    B2R_CUDA(cudaGetLastError());
    h->n_launches++;
    B2R_CUDA(cudaMemcpyAsync(&after, h->counters, 8, cudaMemcpyDeviceToHost, s));
    B2R_CUDA(cudaStreamSynchronize(s));
    h->live -= (int64_t)(after - before);
    h->mut_gen++;
    return B2R_OK;
}

extern "C" int b2r_get_rows_f32(b2r_handle h, const int64_t *rows, int64_t n, float *out, void *stream) {
    B2R_REQUIRE(h, "b2r_get_rows_f32: NULL handle");
    B2R_REQUIRE(n >= 0 && (n == 0 || (rows && out)), "b2r_get_rows_f32: bad arguments");
    if (n == 0) return B2R_OK;
    std::lock_guard<std::mutex> g(h->mu);
    B2R_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    for (int64_t i = 0; i < n; ++i)
        B2R_REQUIRE(rows[i] >= 0 && rows[i] < h->rows, "b2r_get_rows_f32: row out of range");
    int rc = ensure(h->rows_stage, (size_t)n * 8);
    if (rc != B2R_OK) return rc;
    B2R_CUDA(cudaMemcpyAsync(h->rows_stage.p, rows, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    const bool dev_out = is_device_ptr(out);
    float *od = out;
    if (!dev_out) {
        rc = ensure(h->gather_out, (size_t)n * h->dim * 4);
        if (rc != B2R_OK) return rc;
        od = (float *)h->gather_out.p;
    }
    gather_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(h->master, h->corpus, h->dim, h->dp,
                                                               (const long long *)h->rows_stage.p, n, h->rows, od);
    B2R_CUDA(cudaGetLastError());
    h->n_launches++;
    if (!dev_out) B2R_CUDA(cudaMemcpyAsync(out, od, (size_t)n * h->dim * 4, cudaMemcpyDeviceToHost, s));
    B2R_CUDA(cudaStreamSynchronize(s));
    return B2R_OK;
}

// ---------------------------------------------------------------------------------
// metadata columns and compiled where clauses
// ---------------------------------------------------------------------------------
extern "C" int b2r_column_set(b2r_handle h, int column, int64_t first_row, int64_t n, const int32_t *codes, void *stream) {
    B2R_REQUIRE(h, "b2r_column_set: NULL handle");
    B2R_REQUIRE(column >= 0 && column < B2R_MAX_COLUMNS, "b2r_column_set: column must be 0..15");
    B2R_REQUIRE(n >= 0 && first_row >= 0 && (n == 0 || codes), "b2r_column_set: bad arguments");
    std::lock_guard<std::mutex> g(h->mu);
    B2R_REQUIRE(first_row + n <= h->rows, "b2r_column_set: rows must have been ingested first");
    if (n == 0) return B2R_OK;
    B2R_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (!h->cols[column]) {
        B2R_CUDA(cudaMalloc(&h->cols[column], (size_t)h->capacity * 4));
        B2R_CUDA(cudaMemsetAsync(h->cols[column], 0xff, (size_t)h->capacity * 4, s));
    }
    const bool dev = is_device_ptr(codes);
    B2R_CUDA(cudaMemcpyAsync(h->cols[column] + first_row, codes, (size_t)n * 4, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    if (!dev) B2R_CUDA(cudaStreamSynchronize(s));      // the caller's host array may go away
    h->mut_gen++;
    return B2R_OK;
}

namespace {
struct WhereProg {
    int n_nodes;
    b2r_where_node node[B2R_WHERE_MAX_NODES];
    const int32_t *cols[B2R_MAX_COLUMNS];     // nullptr: no row carries this key
};

// One warp = 128 consecutive rows per step (lane handles rows lane, lane+32, lane+64, lane+96 of the group: four coalesced
// loads per leaf in flight), one output word per 32 rows.  The clause is a postfix program over look-up-table leaves, the
// same for every row, so control flow is uniform; the operand stacks are bit fields in registers.
__global__ void __launch_bounds__(256)
where_bits_kernel(const WhereProg prog, const uint32_t *__restrict__ lut, const uint32_t *__restrict__ allow_in,
                  unsigned n, unsigned n_words, uint32_t *__restrict__ out) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long gw = (unsigned long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const unsigned long long nw = (unsigned long long)gridDim.x * (blockDim.x / 32);
    const unsigned long long total = (unsigned long long)n_words * 32;
    for (unsigned long long base = gw * 128; base < total; base += nw * 128) {
        unsigned stack[4] = {0u, 0u, 0u, 0u};
        int sp = 0;
        for (int i = 0; i < prog.n_nodes; ++i) {
            const b2r_where_node nd = prog.node[i];
            if (nd.op == B2R_WHERE_LEAF) {
                const int32_t *col = prog.cols[nd.column];
                int code[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned long long r = base + lane + 32u * j;
                    code[j] = (col && r < n) ? col[r] : -1;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    unsigned bit = 0;
                    if (code[j] >= 0 && (unsigned)code[j] < nd.lut_values)
                        bit = (lut[nd.lut_offset + ((unsigned)code[j] >> 5)] >> (code[j] & 31)) & 1u;
                    stack[j] |= bit << sp;
                }
                ++sp;
            } else {
                sp -= 2;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned a = (stack[j] >> (sp + 1)) & 1u, b = (stack[j] >> sp) & 1u;
                    const unsigned v = nd.op == B2R_WHERE_AND ? (a & b) : (a | b);
                    stack[j] = (stack[j] & ~(3u << sp)) | (v << sp);
                }
                ++sp;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned long long r = base + lane + 32u * j;
            bool ok = r < n && (stack[j] & 1u);
            if (ok && allow_in) ok = (allow_in[r >> 5] >> (r & 31)) & 1u;
            const unsigned w = __ballot_sync(FULL_MASK, ok);
            if (lane == 0 && r < total) out[r >> 5] = w;
        }
    }
}

// Validate `w`, stage its tables, and write the clause bitmap (ANDed with allow_dev when given) for the
// first n_words * 32 rows into h->where_bits.
int run_where(b2r_index *h, const b2r_where *w, const uint32_t *allow_dev, unsigned n_words, cudaStream_t s) {
    B2R_REQUIRE(w->n_nodes >= 1 && w->n_nodes <= B2R_WHERE_MAX_NODES, "where: clause has too many nodes (max 32)");
    B2R_REQUIRE(w->lut_words >= 0 && (w->lut_words == 0 || w->lut), "where: missing look-up tables");
    WhereProg prog;
    prog.n_nodes = w->n_nodes;
    int depth = 0;
    for (int i = 0; i < w->n_nodes; ++i) {
        const b2r_where_node &nd = w->nodes[i];
        prog.node[i] = nd;
        if (nd.op == B2R_WHERE_LEAF) {
            B2R_REQUIRE(nd.column >= 0 && nd.column < B2R_MAX_COLUMNS, "where: leaf column out of range");
            B2R_REQUIRE((uint64_t)nd.lut_offset + (nd.lut_values + 31) / 32 <= (uint64_t)w->lut_words, "where: leaf table out of range");
            ++depth;
        } else {
            B2R_REQUIRE(nd.op == B2R_WHERE_AND || nd.op == B2R_WHERE_OR, "where: unknown node");
            B2R_REQUIRE(depth >= 2, "where: malformed postfix clause");
            --depth;
        }
        B2R_REQUIRE(depth <= 32, "where: clause nests too deep");
    }
    B2R_REQUIRE(depth == 1, "where: malformed postfix clause");
    for (int c = 0; c < B2R_MAX_COLUMNS; ++c) prog.cols[c] = h->cols[c];
    int rc;
    const uint32_t *lut_dev = w->lut;
    if (w->lut_words > 0 && !is_device_ptr(w->lut)) {
        if ((rc = ensure(h->where_lut, (size_t)w->lut_words * 4)) != B2R_OK) return rc;
        B2R_CUDA(cudaMemcpyAsync(h->where_lut.p, w->lut, (size_t)w->lut_words * 4, cudaMemcpyHostToDevice, s));
        lut_dev = (const uint32_t *)h->where_lut.p;
    }
    if ((rc = ensure(h->where_bits, (size_t)std::max(n_words, 1u) * 4)) != B2R_OK) return rc;
    const unsigned blocks = std::max(1u, std::min((n_words * 32 + 1023) / 1024, (unsigned)h->sm_count * 8));   // 8 warps x 128 rows per block and step
    where_bits_kernel<<<blocks, 256, 0, s>>>(prog, lut_dev, allow_dev, (unsigned)h->rows, n_words, (uint32_t *)h->where_bits.p);
    B2R_CUDA(cudaGetLastError());
    h->n_launches++;
    return B2R_OK;
}

// identity of a compiled clause with host-resident tables: FNV-1a over nodes and tables (never 0)
uint64_t where_key(const b2r_where *w) {
    if (!w || w->n_nodes < 1 || w->n_nodes > B2R_WHERE_MAX_NODES || w->lut_words < 0 || (w->lut_words && is_device_ptr(w->lut))) return 0;
    uint64_t hsh = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t n) {
        const unsigned char *b = (const unsigned char *)p;
        for (size_t i = 0; i < n; ++i) { hsh ^= b[i]; hsh *= 1099511628211ull; }
    };
    mix(&w->n_nodes, sizeof(w->n_nodes));
    mix(w->nodes, sizeof(b2r_where_node) * (size_t)w->n_nodes);
    mix(&w->lut_words, sizeof(w->lut_words));
    if (w->lut_words) mix(w->lut, (size_t)w->lut_words * 4);
    return hsh ? hsh : 1;
}

// host allow bitmap -> h->allow (device)
int stage_allow(b2r_index *h, const b2r_filter &f, const uint32_t **allow_dev, cudaStream_t s) {
    *allow_dev = f.allow_bits;
    if (f.allow_bits && !is_device_ptr(f.allow_bits)) {
        const size_t words = (size_t)((h->rows + 31) / 32);
        int rc = ensure(h->allow, std::max<size_t>(words, 1) * 4);
        if (rc != B2R_OK) return rc;
        B2R_CUDA(cudaMemcpyAsync(h->allow.p, f.allow_bits, words * 4, cudaMemcpyHostToDevice, s));
        *allow_dev = (const uint32_t *)h->allow.p;
    }
    return B2R_OK;
}
}  // namespace

extern "C" int b2r_filter_eval(b2r_handle h, const b2r_filter *filter, uint32_t *out_bits, void *stream) {
    B2R_REQUIRE(h && out_bits, "b2r_filter_eval: NULL argument");
    std::lock_guard<std::mutex> g(h->mu);
    B2R_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    b2r_filter f; f.type_mask = ~0ull; f.allow_bits = nullptr; f.where = nullptr;
    if (filter) f = *filter;
    f.type_mask &= ~(1ull << B2R_TYPE_DEAD);
    const unsigned n_words = (unsigned)((h->rows + 31) / 32);
    if (n_words == 0) return B2R_OK;
    const uint32_t *allow_dev = nullptr;
    int rc;
    if ((rc = stage_allow(h, f, &allow_dev, s)) != B2R_OK) return rc;
    if (f.where) {
        if ((rc = run_where(h, f.where, allow_dev, n_words, s)) != B2R_OK) return rc;
        h->wb_key = 0;                           // the remembered clause bitmap lived here
        allow_dev = (const uint32_t *)h->where_bits.p;
    }
    const bool dev_out = is_device_ptr(out_bits);
    uint32_t *od = out_bits;
    if (!dev_out) {
        if ((rc = ensure(h->pass_bits, (size_t)n_words * 4 + 16)) != B2R_OK) return rc;
        od = (uint32_t *)h->pass_bits.p;
        h->pb_buf = nullptr;                     // the cached query bitmap lived here
    }
    B2R_CUDA(pass_bits_launch(h->type_code, f.type_mask, allow_dev, (unsigned)h->rows, n_words, od, h->sm_count, s));
    h->n_launches++;
    if (!dev_out) B2R_CUDA(cudaMemcpyAsync(out_bits, od, (size_t)n_words * 4, cudaMemcpyDeviceToHost, s));
    B2R_CUDA(cudaStreamSynchronize(s));
    return B2R_OK;
}

// ---------------------------------------------------------------------------------
// persistence: the shard as it sits in HBM <-> one file
// ---------------------------------------------------------------------------------
namespace {
struct ShardHeader {               // 128 bytes, little endian
    char magic[4];                 // "B2RS"
    uint32_t version;              // 1
    int32_t dim, dp, space;
    uint32_t flags;
    int64_t rows, live, row_base;
    float max_norm2[2];
    uint64_t checksum;             // sum of the payload as little-endian u64 words (tail bytes zero-extended) + byte count
    uint64_t payload_bytes;
    unsigned char reserved[56];
};
static_assert(sizeof(ShardHeader) == 128, "header layout");
constexpr size_t IO_CHUNK = 64u << 20;

uint64_t sum_words(const unsigned char *p, size_t n) {
    uint64_t s = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; std::memcpy(&w, p + i, 8); s += w; }
    if (i < n) { uint64_t w = 0; std::memcpy(&w, p + i, n - i); s += w; }
    return s;
}
struct PinnedBuf {
    void *p = nullptr;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};
struct FileCloser {
    FILE *f = nullptr;
    ~FileCloser() { if (f) fclose(f); }
};
}  // namespace

extern "C" int b2r_save(b2r_handle h, const char *path, void *stream) {
    B2R_REQUIRE(h && path, "b2r_save: NULL argument");
    std::lock_guard<std::mutex> g(h->mu);
    B2R_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    B2R_CUDA(cudaStreamSynchronize(s));
    ShardHeader hd;
    std::memset(&hd, 0, sizeof(hd));
    std::memcpy(hd.magic, "B2RS", 4);
    hd.version = 1; hd.dim = h->dim; hd.dp = h->dp; hd.space = h->space; hd.flags = h->flags;
    hd.rows = h->rows; hd.live = h->live; hd.row_base = h->row_base;
    B2R_CUDA(cudaMemcpy(hd.max_norm2, h->max_norm2, 8, cudaMemcpyDeviceToHost));
    const std::string tmp = std::string(path) + ".tmp";
    FileCloser fc;
    fc.f = fopen(tmp.c_str(), "wb");
    if (!fc.f) { set_error(std::string("b2r_save: cannot open ") + tmp); return B2R_EINVAL; }
    if (fwrite(&hd, sizeof(hd), 1, fc.f) != 1) { set_error("b2r_save: write failed"); return B2R_EINVAL; }
    PinnedBuf pb;
    B2R_CUDA(cudaHostAlloc(&pb.p, IO_CHUNK, cudaHostAllocDefault));
    const struct { const void *ptr; size_t bytes; } sect[4] = {
        {h->corpus, (size_t)h->rows * h->dp * 2},
        {h->master, h->master ? (size_t)h->rows * h->dp * 4 : 0},
        {h->bias, h->bias ? (size_t)h->rows * 4 : 0},
        {h->type_code, (size_t)h->rows}};
    uint64_t sum = 0, total = 0;
    for (const auto &sc : sect) {
        for (size_t off = 0; off < sc.bytes; off += IO_CHUNK) {
            const size_t n = std::min(IO_CHUNK, sc.bytes - off);
            B2R_CUDA(cudaMemcpyAsync(pb.p, (const char *)sc.ptr + off, n, cudaMemcpyDeviceToHost, s));
            B2R_CUDA(cudaStreamSynchronize(s));
            sum += sum_words((const unsigned char *)pb.p, n);      // chunks are multiples of 8 except a section's last
            if (fwrite(pb.p, 1, n, fc.f) != n) { set_error("b2r_save: write failed (disk full?)"); return B2R_EINVAL; }
            total += n;
        }
    }
    hd.checksum = sum + total; hd.payload_bytes = total;
    if (fseek(fc.f, 0, SEEK_SET) != 0 || fwrite(&hd, sizeof(hd), 1, fc.f) != 1 || fflush(fc.f) != 0) {
        set_error("b2r_save: write failed"); return B2R_EINVAL;
    }
    fclose(fc.f); fc.f = nullptr;
    if (rename(tmp.c_str(), path) != 0) { set_error(std::string("b2r_save: cannot rename to ") + path); return B2R_EINVAL; }
    return B2R_OK;
}

extern "C" int b2r_load(const char *path, int device, int64_t capacity_rows, b2r_handle *out) {
    B2R_REQUIRE(path && out, "b2r_load: NULL argument");
    *out = nullptr;
    FileCloser fc;
    fc.f = fopen(path, "rb");
    if (!fc.f) { set_error(std::string("b2r_load: cannot open ") + path); return B2R_EINVAL; }
    ShardHeader hd;
    if (fread(&hd, sizeof(hd), 1, fc.f) != 1 || std::memcmp(hd.magic, "B2RS", 4) != 0) {
        set_error(std::string("b2r_load: ") + path + " is not a b2r shard file"); return B2R_EINVAL;
    }
    B2R_REQUIRE(hd.version == 1, "b2r_load: unknown shard file version");
    B2R_REQUIRE(hd.dim >= 1 && hd.dim <= 8192 && hd.dp == round_up(hd.dim, 64) && hd.space >= 0 && hd.space <= 2 &&
                    hd.rows >= 0 && hd.live >= 0 && hd.live <= hd.rows, "b2r_load: corrupt header");
    const bool has_master = !(hd.flags & B2R_FLAG_NO_F32_MASTER), has_bias = hd.space == B2R_SPACE_L2;
    const size_t sect_bytes[4] = {(size_t)hd.rows * hd.dp * 2, has_master ? (size_t)hd.rows * hd.dp * 4 : 0,
                                  has_bias ? (size_t)hd.rows * 4 : 0, (size_t)hd.rows};
    B2R_REQUIRE(hd.payload_bytes == sect_bytes[0] + sect_bytes[1] + sect_bytes[2] + sect_bytes[3],
                "b2r_load: payload size does not match the header");
    b2r_handle h = nullptr;
    int rc = b2r_create(hd.dim, hd.space, std::max<int64_t>(capacity_rows, hd.rows), device, hd.flags, &h);
    if (rc != B2R_OK) return rc;
    PinnedBuf pb;
    cudaError_t ce = cudaHostAlloc(&pb.p, IO_CHUNK, cudaHostAllocDefault);
    if (ce != cudaSuccess) { b2r_destroy(h); set_error("b2r_load: cudaHostAlloc failed"); return B2R_ENOMEM; }
    void *dst[4] = {h->corpus, h->master, h->bias, h->type_code};
    uint64_t sum = 0, total = 0;
    for (int i = 0; i < 4; ++i) {
        for (size_t off = 0; off < sect_bytes[i]; off += IO_CHUNK) {
            const size_t n = std::min(IO_CHUNK, sect_bytes[i] - off);
            if (fread(pb.p, 1, n, fc.f) != n) { b2r_destroy(h); set_error("b2r_load: file is truncated"); return B2R_EINVAL; }
            sum += sum_words((const unsigned char *)pb.p, n);
            ce = cudaMemcpy((char *)dst[i] + off, pb.p, n, cudaMemcpyHostToDevice);
            if (ce != cudaSuccess) { b2r_destroy(h); set_error(std::string("b2r_load: ") + cudaGetErrorString(ce)); return B2R_ECUDA; }
            total += n;
        }
    }
    if (sum + total != hd.checksum) { b2r_destroy(h); set_error("b2r_load: checksum mismatch (file is corrupt)"); return B2R_EINVAL; }
    ce = cudaMemcpy(h->max_norm2, hd.max_norm2, 8, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { b2r_destroy(h); set_error(std::string("b2r_load: ") + cudaGetErrorString(ce)); return B2R_ECUDA; }
    h->rows = hd.rows; h->live = hd.live; h->row_base = hd.row_base; h->mut_gen++;
    *out = h;
    return B2R_OK;
}

// ---------------------------------------------------------------------------------
// query
// ---------------------------------------------------------------------------------
namespace {

// RAII-less pair of events around one scoring-kernel launch (only when h->timing)
struct KernelTimer {
    b2r_index *h; cudaStream_t s; std::pair<cudaEvent_t, cudaEvent_t> ev; bool on;
    KernelTimer(b2r_index *h_, cudaStream_t s_, int stage = 0) : h(h_), s(s_), on(h_->timing && h_->timing_stage == stage) {
        if (!on) return;
        if (!h->ev_free.empty()) { ev = h->ev_free.back(); h->ev_free.pop_back(); }
        else if (cudaEventCreate(&ev.first) != cudaSuccess || cudaEventCreate(&ev.second) != cudaSuccess) { on = false; return; }
        cudaEventRecord(ev.first, s);
    }
    void stop() {
        if (!on) return;
        cudaEventRecord(ev.second, s);
        h->ev_pending.push_back(ev);
    }
};

int launch_scan_batch(b2r_index *h, int nq, int epl, const ScanParams &base, cudaStream_t s) {
    int q = 0;
    while (q < nq) {
        int grp = (nq - q >= 4 && epl <= 2) ? 4 : (nq - q >= 2 && epl <= 2) ? 2 : 1;
        int max_grid = scan_max_grid(h->dp, grp, epl, h->sm_count);
        if (max_grid <= 0) { set_error("b2r_query: scan kernel cannot be resident (occupancy 0)"); return B2R_ECUDA; }
        const int tile = scan_tile_rows(h->dp);
        int64_t tiles = (h->rows + tile - 1) / tile;
        int grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, max_grid));
        int rc = ensure(h->scan_lists, sizeof(KeyS) * (size_t)max_grid * 4 * 32 * epl);
        if (rc != B2R_OK) return rc;
        ScanParams p = base;
        p.q0 = q;
        p.cta_lists = (KeyS *)h->scan_lists.p;
        p.stage_keys = scan_stage_keys(epl, grid);
        KernelTimer kt(h, s);
        B2R_CUDA(scan_launch(h->dp, grp, epl, p, grid, s));
        kt.stop();
        h->n_launches++;
        q += grp;
    }
    return B2R_OK;
}

// K5 over the whole batch (force_all) or over the work list of failed certificates.  A launch covers a bounded range of
// work items (one slot of per-CTA lists per group of G queries, no slot is ever reused inside a launch, so no CTA waits
// for another); the fix-up cannot know the length of its work list without a host sync, so it enqueues one launch per
// range the batch could fill -- one launch up to a few hundred queries, and a range that turns out empty costs ~2 us.
constexpr size_t EXACT_LISTS_BUDGET = 64u << 20;      // per-CTA list scratch per handle

int launch_exact_batch(b2r_index *h, int nq, int k, int force_all, const FinalizeParams &fin,
                       const b2r_filter &f, const uint32_t *allow_dev, cudaStream_t s, const XchgDev *rider = nullptr,
                       size_t rider_smem = 0) {
    const int epl = epl_exact(k);
    const int G = exact_group(epl, h->dp);
    int max_grid = exact_max_grid(epl, h->dp, h->sm_count);
    if (max_grid <= 0) { set_error("b2r_query: exact kernel cannot be resident"); return B2R_ECUDA; }
    int64_t warps_needed = std::max<int64_t>(1, (h->rows + 3) / 4);
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>((warps_needed + EXACT_WARPS - 1) / EXACT_WARPS, max_grid));
    // the certificate fix-up almost never has work: on small shards one CTA per SM keeps the empty launch short
    // (3.4 -> ~2 us); on large ones the empty launch is noise and a query that does need the scan gets the whole machine
    if (!force_all && h->rows < (4ll << 20)) grid = std::min(grid, h->sm_count);
    const size_t per_group = sizeof(KeyD) * (size_t)G * grid * 32 * epl;
    const int slots = (int)std::max<size_t>(1, std::min<size_t>(EXACT_MAX_SLOTS, EXACT_LISTS_BUDGET / per_group));
    const int items = slots * G;
    int rc = ensure(h->exact_lists, per_group * slots);
    if (rc != B2R_OK) return rc;
    ExactParams p;
    p.type_code = h->type_code; p.allow_bits = allow_dev; p.type_mask = f.type_mask;
    p.n = (unsigned)h->rows; p.nq = nq; p.force_all = force_all; p.items = items;
    p.cta_lists = (KeyD *)h->exact_lists.p; p.tickets = h->tickets + 1;
    p.n_fallbacks = (long long *)(h->counters + 1);
    p.fin = fin;
    for (int first = 0; first < nq; first += items) {
        p.first_item = first;
        const bool last = first + items >= nq;                            // the call's very last launch
        p.rider.nq = 0;
        if (rider && last) p.rider = *rider;
        KernelTimer kt(h, s, force_all ? 0 : 5);   // the certificate fix-up is not the scoring kernel
        B2R_CUDA(exact_launch(epl, p, grid, s, (rider && last) ? rider_smem : 0));
        kt.stop();
        h->n_launches++;
    }
    return B2R_OK;
}

// K3: pass bitmap -> tcgen05 scoring + per-(query, slice) lists -> per-query finalize
constexpr int GEMM_MIN_BATCH = 5;        // below this the scan reads the corpus at most twice anyway
constexpr int GEMM_MAX_QBLOCKS = 8;      // 128-query blocks per launch (1024 queries per corpus pass)
constexpr int GEMM_REGION_CAP = 512;       // pool mode: entries per private (query, slice, half) region; a full one keeps its best half
constexpr int GEMM_POOL_CAP = 16384;        // pool mode: compact pool entries per query (>= SMs*2*32 for the sampling pass)

int launch_gemm_batch(b2r_index *h, int nq, int k, int epl, const FinalizeParams &fin, const b2r_filter &f,
                      const uint32_t *allow_dev, uint64_t filter_key, cudaStream_t s) {
    const int L = gemm_list_len(k);                    // 8 / 16 / 32, or 0 = pool mode (32 < k <= 128)
    const bool pool_mode = L == 0;
    const int L_seed = pool_mode ? GEMM_POOL_SAMPLE_RANK : L;     // rank of the sampled score that becomes the bound
    const int BN = gemm_tile_rows(h->dp);
    const int tiles_total = (int)((h->rows + BN - 1) / BN);
    const unsigned n_words = (unsigned)tiles_total * (unsigned)(BN / 32);
    const int qblocks_total = (nq + GEMM_BM - 1) / GEMM_BM;
    // pool capacity per query: list mode = every list full; pool mode = 16x the expected 1024 entries (beyond that
    // the query is flagged and re-done by the exact scan)
    const int list_stride = pool_mode ? GEMM_POOL_CAP : std::max(2 * FIN_THREADS, h->sm_count * GEMM_HALVES * L);   // the finalize reads 512 slots unconditionally
    int rc;
    if ((rc = ensure(h->pass_bits, (size_t)n_words * 4 + 16)) != B2R_OK) return rc;
    // h->gthr (bounds, cursors, seed flags, arrival counters) was sized and cleared by the query preparation
    // list mode pools: [nq][SMs*2*L].  Pool mode: the sampling pass needs [nq][SMs*2*32] and the main pass
    // compacts at most [launch queries][slots*cap] -- both fit the same allocation.
    if ((rc = ensure(h->gemm_lists, sizeof(KeyS) * (size_t)nq * list_stride)) != B2R_OK) return rc;
    // private regions: [launch queries (padded: every lane of a block owns one)][n_slices * 2][cap]; n_qblocks * n_slices
    // <= SMs in every launch, so 128 * SMs * 2 regions cover the largest one
    if (pool_mode &&
        (rc = ensure(h->gemm_regions, sizeof(KeyS) * (size_t)GEMM_BM * h->sm_count * GEMM_HALVES * GEMM_REGION_CAP)) != B2R_OK)
        return rc;
    if (h->trace_on && (rc = ensure(h->trace, sizeof(unsigned long long) * 8 * (size_t)h->sm_count)) != B2R_OK) return rc;
    // cacheable: no bitmap at all (key 0), or a remembered clause's bitmap (its hash); a caller's own allow bitmap is not
    const bool cacheable = !allow_dev || filter_key != 0;
    const bool pb_hit = cacheable && h->pb_buf == h->pass_bits.p && h->pb_gen == h->mut_gen && h->pb_rows == h->rows &&
                        h->pb_mask == f.type_mask && h->pb_bn == BN && h->pb_key == filter_key;
    if (!pb_hit) {
        B2R_CUDA(pass_bits_launch(h->type_code, f.type_mask, allow_dev, (unsigned)h->rows, n_words,
                                  (uint32_t *)h->pass_bits.p, h->sm_count, s));
        h->n_launches++;
        h->pb_buf = cacheable ? h->pass_bits.p : nullptr;
        h->pb_gen = h->mut_gen; h->pb_rows = h->rows; h->pb_mask = f.type_mask; h->pb_bn = BN; h->pb_key = filter_key;
    }
    if (h->tm_corpus_base != h->corpus || h->tm_corpus_rows != h->capacity) {
        if ((rc = gemm_encode_map(&h->tm_corpus, h->corpus, h->dp, (uint64_t)h->capacity, BN)) != B2R_OK) return rc;
        if ((rc = gemm_encode_map(&h->tm_corpus_half, h->corpus, h->dp, (uint64_t)h->capacity, BN / 2)) != B2R_OK) return rc;
        h->tm_corpus_base = h->corpus; h->tm_corpus_rows = h->capacity;
    }
    // batches of at most 64 queries on rows of more than 512 dims run 64-query blocks (tcgen05.mma M = 64): half the tensor work
    // per corpus tile, and the query block becomes resident (at 128 queries it is streamed with every corpus K-block beyond 512
    // dims).  Measured at 10M x 768, batch 64: 2.21 vs 2.77 ms per batch.  Up to 512 dims the 128-query form is 1-2 % faster.
    const int bm = (nq <= 64 && h->dp > 512 && !h->no_bm64) ? 64 : GEMM_BM;
    if (h->tm_query_base != h->q_bf16.p || h->tm_query_rows != nq || h->tm_query_box != bm) {
        if ((rc = gemm_encode_map(&h->tm_query, h->q_bf16.p, h->dp, (uint64_t)nq, bm)) != B2R_OK) return rc;
        h->tm_query_base = h->q_bf16.p; h->tm_query_rows = nq; h->tm_query_box = bm;
    }
    // Threshold seeding happens inside K3 (GemmParams::seed_tiles): every CTA first scans a few tiles of its slice in
    // sampling mode and posts the step maxima, the epilogue warps fold the posts into gthr[q].  List mode: the L-th best
    // post.  Pool mode: the 32nd best, which must be in place before the first append -- the expected pool is then
    // 32 * (shard rows / sampled rows) entries per query whatever the shard size.
    if ((rc = ensure(h->gemm_samples, sizeof(unsigned) * (size_t)GEMM_BM * h->sm_count * GEMM_HALVES * L_seed)) != B2R_OK) return rc;
    for (int qb0 = 0; qb0 < qblocks_total; qb0 += GEMM_MAX_QBLOCKS) {
        GemmParams gp;
        gp.n = (unsigned)h->rows; gp.nq = nq; gp.qblock0 = qb0;
        gp.n_qblocks = std::min(GEMM_MAX_QBLOCKS, qblocks_total - qb0);
        gp.list_stride = list_stride;
        gp.pass_bits = (const uint32_t *)h->pass_bits.p; gp.bias = h->bias;
        gp.gthr = (unsigned *)h->gthr.p; gp.cnt = gp.gthr + (size_t)qblocks_total * GEMM_BM; gp.lists = (KeyS *)h->gemm_lists.p;
        gp.regions = (KeyS *)h->gemm_regions.p; gp.region_cap = GEMM_REGION_CAP;
        gp.samples = (unsigned *)h->gemm_samples.p; gp.seeded = gp.cnt + (size_t)qblocks_total * GEMM_BM; gp.arrive = gp.seeded + (size_t)qblocks_total * GEMM_BM;
        gp.tile_counter = nullptr;
        gp.seed_tiles = 0;
        // the bound is seeded from the k-th best sampled score where that was measured to pay (launches of >= 2 query blocks:
        // pools 144 -> 87 entries per query at k = 5, 286 -> 178 at k = 10, -1..4 % per step, -3..5 % at 512 dims / k = 10:
        // the finalize sorts less); single-block launches keep the L-th best (batch 1: no gain measured)
        gp.seed_rank = pool_mode ? GEMM_POOL_SAMPLE_RANK : ((h->seed_rank_l || k > L || gp.n_qblocks < 2) ? L : k);
        gp.seed_wait_ns = h->seed_wait_ns; gp.delay_us = h->delay_us;
        gp.trace = h->trace_on ? (unsigned long long *)h->trace.p : nullptr; gp.trace_mode = h->trace_mode;
        UnionParams un;
        un.lists = gp.lists; un.list_stride = list_stride; un.gthr = gp.gthr; un.cnt = gp.cnt;
        un.pool_stats = h->counters + 2;
        const int q0 = qb0 * GEMM_BM, nq_here = std::min(nq - q0, gp.n_qblocks * GEMM_BM);
        gp.tiles_total = tiles_total;
        gp.n_slices = std::max(1, std::min(h->sm_count / gp.n_qblocks, tiles_total));
        // an even number of query blocks runs as CTA pairs (cta_group::2): query blocks 2j and 2j+1 of a slice share every
        // corpus tile, each CTA stages half of it
        bool pair = false;
        if (!h->no_pair && gp.n_qblocks % 2 == 0) {
            const int max_pairs = gemm_max_pairs(h->dp, L, h->bias != nullptr);
            if (max_pairs >= gp.n_qblocks / 2) {
                pair = true;
                gp.n_slices = std::max(1, std::min(gp.n_slices, max_pairs / (gp.n_qblocks / 2)));
            }
        }
        {   // seeding tiles per CTA: >= 128 tiles (32k rows) over the block's slices, more on long slices (<= 1/64 extra work).
            // The pair form samples >= 192 tiles (3 per pair at 74 pairs): a tighter bound shortens the candidate pools the
            // finalize sorts (105 -> 70 entries per query at 1M rows) by more than the extra tile costs -- 2 / 3 / 4 / 6 tiles
            // per pair, A/B on one box, sustained: 191.9 / 189.4 / 192.5 / 199.5 us per batch-256 step (gpurun_out/seed_tiles2.log)
            const int tiles_per_cta = tiles_total / gp.n_slices;
            const int want128 = std::max((128 + gp.n_slices - 1) / gp.n_slices, std::min(tiles_per_cta / 64, 8));
            const int want192 = std::max((192 + gp.n_slices - 1) / gp.n_slices, std::min(tiles_per_cta / 64, 8));
            const int want = (pair && tiles_per_cta >= 8 * want192) ? want192 : want128;     // short slices keep the smaller sample
            gp.seed_stride = 1;
            if (pool_mode) {   // whenever there is anything to sample: max(4 tiles, 1/32 of the shard) spread over the slices, so
                               // that the 32nd best sample leaves ~1024 candidates per query -- far more than k, or the
                               // certificate (k-th exact candidate vs the bound) could not hold.  Without a seed (tiny
                               // shards, B2R_NO_SEED) every thread bounds itself from its own region.
                // (1/32 of the shard; 1/64 from 8M rows on: the sampling tiles are scored twice, and on a long scan the half-size
                // sample is worth more than the pools it doubles to ~2000 entries cost the finalize: -1.2 % at 12.5M and 50M rows)
                const int div = h->pool_sample_div > 0 ? h->pool_sample_div : (h->rows >= (8ll << 20) ? 64 : 32);
                const int sample_tiles = (tiles_total >= 8 && !h->no_seed) ? std::max(4, tiles_total / div) : 0;
                gp.seed_tiles = sample_tiles ? std::max(1, sample_tiles / gp.n_slices) : 0;
                if (sample_tiles && sample_tiles < gp.n_slices) gp.seed_stride = gp.n_slices / sample_tiles;
            } else {
                gp.seed_tiles = (!h->no_seed && nq >= h->seed_min_batch && tiles_per_cta >= 8 * want) ? want : 0;
            }
            if (gp.seed_tiles && h->seed_tiles_override > 0) gp.seed_tiles = std::min(h->seed_tiles_override, std::max(1, tiles_per_cta / 2));
        }
        // one query block, list mode, plenty of tiles: the CTAs take tiles dynamically instead of owning a static slice
        if (!pair && !pool_mode && !h->no_dyn && gp.n_qblocks == 1 && tiles_total >= 8 * gp.n_slices)
            gp.tile_counter = gp.arrive + (size_t)qblocks_total;
        un.max_entries = pool_mode ? GEMM_POOL_CAP : gp.n_slices * GEMM_HALVES * L;
        if (h->trace_on) B2R_CUDA(cudaMemsetAsync(h->trace.p, 0, sizeof(unsigned long long) * 8 * (size_t)h->sm_count, s));
        KernelTimer kt(h, s);
        B2R_CUDA(gemm_launch(h->dp, L, h->bias != nullptr, pair, pair ? GEMM_BM : bm, h->tm_query, pair ? h->tm_corpus_half : h->tm_corpus, gp, s));
        kt.stop();
        h->trace_ctas = gp.n_slices * gp.n_qblocks;
        h->n_launches++;
        KernelTimer kt4(h, s, 4);
        B2R_CUDA(finalize_union_launch(epl, fin, un, q0, nq_here, s));
        kt4.stop();
        h->n_launches++;
    }
    return B2R_OK;
}

}  // namespace

// the query itself; the caller holds h->mu
static int query_locked(b2r_handle h, const float *q, int nq, int k, const b2r_filter *filter,
                        int64_t *out_rows, float *out_dist, double *out_dist64, int32_t *out_count,
                        void *stream, b2r_xchg *push_to = nullptr, int64_t *merge_rows = nullptr, float *merge_dist = nullptr,
                        int32_t *merge_count = nullptr) {
    B2R_REQUIRE(nq >= 1 && q, "b2r_query: need at least one query");
    B2R_REQUIRE(k >= 1, "b2r_query: n_results must be a positive integer");
    B2R_REQUIRE(out_rows && out_dist && out_count, "b2r_query: NULL output");
    if (epl_exact(k) == 0) { set_error("b2r_query: k > 256 is not supported"); return B2R_EUNSUPPORTED; }
    B2R_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const bool dev_out = is_device_ptr(out_rows);
    B2R_REQUIRE(!push_to || dev_out, "b2r_query_push: outputs must be device pointers");
    B2R_REQUIRE(dev_out == is_device_ptr(out_dist) && dev_out == is_device_ptr(out_count) &&
                    (!out_dist64 || dev_out == is_device_ptr(out_dist64)),
                "b2r_query: outputs must all be host or all be device pointers");
    b2r_filter f; f.type_mask = ~0ull; f.allow_bits = nullptr; f.where = nullptr;
    if (filter) f = *filter;
    f.type_mask &= ~(1ull << B2R_TYPE_DEAD);

    int rc;
    // ---- stage inputs ----
    const float *q_raw = q;
    if (!is_device_ptr(q)) {
        if ((rc = ensure(h->q_raw, (size_t)nq * h->dim * 4)) != B2R_OK) return rc;
        B2R_CUDA(cudaMemcpyAsync(h->q_raw.p, q, (size_t)nq * h->dim * 4, cudaMemcpyHostToDevice, s));
        q_raw = (const float *)h->q_raw.p;
    }
    const uint32_t *allow_dev = nullptr;
    if ((rc = stage_allow(h, f, &allow_dev, s)) != B2R_OK) return rc;
    bool staged_host_filter = (f.allow_bits && allow_dev != f.allow_bits);
    // A clause on its own (no allow bitmap beside it) is remembered by its hash: a session that keeps asking with the same
    // filter re-uses the clause bitmap and the pass bitmap until rows / tombstones / columns change (mut_gen).
    uint64_t filter_key = (f.where && !f.allow_bits) ? where_key(f.where) : 0;
    if (f.where && h->rows > 0) {     // compiled clause -> device bitmap (ANDed with the allow bitmap), then it IS the allow bitmap
        const bool hit = filter_key && h->wb_key == filter_key && h->wb_gen == h->mut_gen && h->wb_rows == h->rows && h->where_bits.p;
        if (!hit) {
            if ((rc = run_where(h, f.where, allow_dev, (unsigned)((h->rows + 31) / 32), s)) != B2R_OK) return rc;
            h->wb_key = filter_key; h->wb_gen = h->mut_gen; h->wb_rows = h->rows;
            staged_host_filter = staged_host_filter || (f.where->lut_words > 0 && !is_device_ptr(f.where->lut));
        }
        allow_dev = (const uint32_t *)h->where_bits.p;
    }
    if ((rc = ensure(h->q_prep, (size_t)nq * h->dp * 4)) != B2R_OK) return rc;
    if ((rc = ensure(h->need_list, (size_t)nq * 4)) != B2R_OK) return rc;
    if ((rc = ensure(h->q_eps, (size_t)nq * 16)) != B2R_OK) return rc;
    long long *o_rows = (long long *)out_rows; float *o_dist = out_dist; double *o_dist64 = out_dist64; int *o_count = out_count;
    // host outputs: one packed device buffer [rows | dist64 | dist | count] -> ONE device-to-host copy into a pinned
    // mirror -> the caller's arrays (three or four small copies would each pay the PCIe round trip)
    const size_t off_d64 = (size_t)nq * k * 8, off_d = off_d64 + (out_dist64 ? (size_t)nq * k * 8 : 0),
                 off_c = off_d + (size_t)nq * k * 4, pack_bytes = off_c + (size_t)nq * 4;
    if (!dev_out) {
        if ((rc = ensure(h->o_pack, pack_bytes)) != B2R_OK) return rc;
        if (h->o_host_bytes < pack_bytes) {
            if (h->o_host) { cudaFreeHost(h->o_host); h->o_host = nullptr; h->o_host_bytes = 0; }
            B2R_CUDA(cudaHostAlloc(&h->o_host, std::max(pack_bytes, (size_t)4096), cudaHostAllocDefault));
            h->o_host_bytes = std::max(pack_bytes, (size_t)4096);
        }
        char *b = (char *)h->o_pack.p;
        o_rows = (long long *)b; o_dist = (float *)(b + off_d); o_count = (int *)(b + off_c);
        if (out_dist64) o_dist64 = (double *)(b + off_d64);
    }

    // ---- choose the scoring path ----
    //   1  warp-shuffle scan (K2): batches of <= 4 queries per corpus pass, every padded dim up to 1024, k <= 128
    //   2  tcgen05 GEMM (K3): one corpus pass per <= 1024 queries, k <= 128, dims 128/256/384/512/768/1024/1536
    //   3  exact fp64 scan (K5): always correct, used for shapes the fast kernels are not built for
    int path = h->path;
    const int epl_s = epl_scored(k);
    const bool scan_ok = scan_supported(h->dp) && epl_s != 0;
    const bool gemm_ok = gemm_supported(h->dp, k);
    // K3's TMA-fed stream is also the faster batch-1 scan up to 512 dims (6.2 vs 5.1 TB/s at 1M x 384); at 768
    // dims its shallower pipeline loses to K2 until the corpus pass is shared by a handful of queries
    // (small shards: K2's single launch has the lower fixed cost)
    if (path == 0) path = (gemm_ok && (nq >= GEMM_MIN_BATCH || (h->dp <= 512 && h->rows >= 262144))) ? 2 : scan_ok ? 1 : 3;
    if (path == 2 && !gemm_ok) path = scan_ok ? 1 : 3;
    if (path == 1 && !scan_ok) path = 3;
    if (h->rows == 0) path = 3;   // empty collection: Chroma returns empty lists; only the padding is written

    // ---- prepare queries (cosine: hnswlib normalisation; zero-pad to dp; bf16 copy for K3) ----
    FusedCall fc;
    if (path == 2 && (rc = ensure(h->q_bf16, (size_t)nq * h->dp * 2)) != B2R_OK) return rc;
    if (path == 2 && (rc = ensure(h->q_err, (size_t)nq * 4)) != B2R_OK) return rc;
    {
        IngestParams p;
        p.x = q_raw; p.n = nq; p.d = h->dim; p.dp = h->dp; p.space = h->space;
        p.corpus = path == 2 ? (uint4 *)h->q_bf16.p : nullptr; p.master = (float *)h->q_prep.p; p.bias = nullptr;
        p.type_out = nullptr; p.type_in = nullptr; p.max_norm2 = nullptr;
        p.qerr = path == 2 ? (float *)h->q_err.p : nullptr;
        // Error bound of the scan scores, per query, from the MEASURED rounding errors (max |x - bf16(x)| from ingest,
        // |q - bf16(q)| when K3 rounds the queries); eps_rel only carries the fp32 accumulation slop:
        // (dp+8) * 2^-24 for K2's FFMA chain, 4x that for the tensor core's accumulator.
        p.q_eps = (double *)h->q_eps.p; p.norms = h->max_norm2;
        p.eps_rel = (float)(h->dp + 8) * 5.9604645e-8f * (path == 2 ? 4.f : 1.f) * 1.01f;
        // per-call shared state, cleared by the preparation: the fix-up work list control, ...
        p.wait_words = nullptr; p.wait_n = 0; p.wait_val = 0;
        std::memset(&p.flags, 0, sizeof p.flags);
        std::memset(&fc, 0, sizeof fc);
        // fused exchange: take the next mailbox slot (and the rider) now -- nothing after this point fails before the kernels are
        // enqueued short of a CUDA error, which breaks the collective anyway
        if (push_to) {
            B2R_REQUIRE(!merge_rows || (merge_dist && merge_count && is_device_ptr(merge_rows) && is_device_ptr(merge_dist) &&
                                        is_device_ptr(merge_count)), "b2r_query_push: the merge outputs must be device pointers");
            if ((rc = xchg_begin_fused(push_to, h->device, nq, k, &fc, merge_rows, merge_dist, merge_count,
                                       exact_smem_limit(epl_exact(k), h->dp))) != B2R_OK) return rc;
            p.wait_words = fc.wait_words; p.wait_n = fc.wait_n; p.wait_val = fc.wait_val;
            p.flags = fc.flags;
        }
        p.zero[0] = (unsigned *)h->need_ctl; p.zero_words[0] = 2;
        p.zero[1] = nullptr; p.zero_words[1] = 0;
        p.zero[2] = nullptr; p.zero_words[2] = 0;
        if (path == 2) {   // ... and K3's [bounds | cursors | seed flags][q-blocks * 128], [arrivals][q-blocks], [tile counters][q-blocks]
            const int qblocks = (nq + GEMM_BM - 1) / GEMM_BM;
            p.zero_words[2] = qblocks * (GEMM_BM * 3 + 2);
            if ((rc = ensure(h->gthr, (size_t)p.zero_words[2] * 4)) != B2R_OK) return rc;
            p.zero[2] = (unsigned *)h->gthr.p;
        }
        const int wpb = INGEST_THREADS / 32;
        int grid = std::min((nq + wpb - 1) / wpb, h->sm_count * 8);
        KernelTimer kt1(h, s, 1);
        B2R_CUDA(launch_pdl(ingest_lookup(h->dim, h->dp, q_raw), dim3(grid), dim3(INGEST_THREADS), 0, s, p));
        kt1.stop();
        h->n_launches++;
    }

    FinalizeParams fin;
    fin.master = h->master; fin.corpus = h->corpus; fin.q = (const float *)h->q_prep.p;
    fin.max_norm2 = h->max_norm2; fin.dp = h->dp; fin.space = h->space; fin.k = k;
    fin.row_base = h->row_base; fin.out_rows = o_rows; fin.out_dist = o_dist; fin.out_dist64 = o_dist64;
    fin.out_count = o_count; fin.need_ctl = h->need_ctl; fin.need_list = (int *)h->need_list.p;
    fin.q_eps = (const double *)h->q_eps.p;
    fin.push = fc.push;
    // fused exchange: the merge of an earlier batch rides in this call's last kernel (the exact scan)
    const XchgDev *rider_p = (push_to && fc.rider.nq && !fc.merge_after) ? &fc.rider : nullptr;
    const size_t rider_smem = fc.rider_smem;
    const bool merge_after = push_to && fc.merge_after;

    if (path == 1) {
        ScanParams sp;
        sp.corpus = h->corpus; sp.bias = h->bias; sp.type_code = h->type_code; sp.allow_bits = allow_dev;
        sp.type_mask = f.type_mask; sp.n = (unsigned)h->rows; sp.q0 = 0; sp.cta_lists = nullptr;
        sp.ticket = h->tickets; sp.fin = fin;
        if ((rc = launch_scan_batch(h, nq, epl_s, sp, s)) != B2R_OK) return rc;
        if ((rc = launch_exact_batch(h, nq, k, 0, fin, f, allow_dev, s, rider_p, rider_smem)) != B2R_OK) return rc;
    } else if (path == 2) {
        // candidates kept by the finalize: KP = 32 / 64 / 128 / 128 / 256 for k <= 8 / 16 / 32 / 64 / 128 (>= 2k beyond 8)
        if ((rc = launch_gemm_batch(h, nq, k, k <= 8 ? 1 : k <= 16 ? 2 : k <= 64 ? 4 : 8, fin, f, allow_dev, filter_key, s)) != B2R_OK) return rc;
        if ((rc = launch_exact_batch(h, nq, k, 0, fin, f, allow_dev, s, rider_p, rider_smem)) != B2R_OK) return rc;
    } else {
        if ((rc = launch_exact_batch(h, nq, k, 1, fin, f, allow_dev, s, rider_p, rider_smem)) != B2R_OK) return rc;
    }
    h->n_queries += nq;
    if (merge_after && (rc = xchg_launch_merge(push_to, fc.rider, s)) != B2R_OK) return rc;

    if (!dev_out) {
        B2R_CUDA(cudaMemcpyAsync(h->o_host, h->o_pack.p, pack_bytes, cudaMemcpyDeviceToHost, s));
        B2R_CUDA(cudaStreamSynchronize(s));
        const char *b = (const char *)h->o_host;
        std::memcpy(out_rows, b, (size_t)nq * k * 8);
        if (out_dist64) std::memcpy(out_dist64, b + off_d64, (size_t)nq * k * 8);
        std::memcpy(out_dist, b + off_d, (size_t)nq * k * 4);
        std::memcpy(out_count, b + off_c, (size_t)nq * 4);
    } else if (q_raw != q || staged_host_filter) {
        B2R_CUDA(cudaStreamSynchronize(s));   // host inputs were staged through reusable buffers
    }
    return B2R_OK;
}

extern "C" int b2r_query_ex(b2r_handle h, const float *q, int nq, int k, const b2r_filter *filter,
                            int64_t *out_rows, float *out_dist, double *out_dist64, int32_t *out_count,
                            void *stream) {
    B2R_REQUIRE(h, "b2r_query: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    return query_locked(h, q, nq, k, filter, out_rows, out_dist, out_dist64, out_count, stream);
}

extern "C" int b2r_query_push(b2r_handle h, b2r_xchg_handle x, const float *q, int nq, int k, const b2r_filter *filter,
                              int64_t *out_rows, float *out_dist, int32_t *out_count, int64_t *merge_rows, float *merge_dist,
                              int32_t *merge_count, void *stream) {
    B2R_REQUIRE(h && x, "b2r_query_push: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    return query_locked(h, q, nq, k, filter, out_rows, out_dist, nullptr, out_count, stream, x, merge_rows, merge_dist, merge_count);
}

extern "C" int b2r_query(b2r_handle h, const float *q, int nq, int k, const b2r_filter *filter,
                         int64_t *out_rows, float *out_dist, int32_t *out_count, void *stream) {
    return b2r_query_ex(h, q, nq, k, filter, out_rows, out_dist, nullptr, out_count, stream);
}

// ---------------------------------------------------------------------------------
// pipelined host-buffer queries
// ---------------------------------------------------------------------------------
namespace {
// results of a finished slot -> the caller's arrays; the slot becomes free.  h->mu held, ev_d2h already reached.
void async_deliver(b2r_index::AsyncSlot &sl) {
    const size_t nk = (size_t)sl.nq * sl.k;
    const char *b = (const char *)sl.out_host;
    std::memcpy(sl.u_rows, b, nk * 8);
    std::memcpy(sl.u_dist, b + nk * 8, nk * 4);
    std::memcpy(sl.u_count, b + nk * 12, (size_t)sl.nq * 4);
    sl.ticket = 0;
}
int pinned_grow(void **p, size_t *have, size_t want) {
    if (*have >= want) return B2R_OK;
    if (*p) { cudaFreeHost(*p); *p = nullptr; *have = 0; }
    want = std::max(want, (size_t)4096);
    B2R_CUDA(cudaHostAlloc(p, want, cudaHostAllocDefault));
    *have = want;
    return B2R_OK;
}
bool is_pinned_host(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
}  // namespace

extern "C" int b2r_query_async(b2r_handle h, const float *q, int nq, int k, const b2r_filter *filter, int64_t *out_rows,
                               float *out_dist, int32_t *out_count, void *stream, uint64_t *ticket_out) {
    B2R_REQUIRE(h && ticket_out, "b2r_query_async: NULL argument");
    B2R_REQUIRE(nq >= 1 && q && k >= 1 && out_rows && out_dist && out_count, "b2r_query_async: bad arguments");
    B2R_REQUIRE(!is_device_ptr(q) && !is_device_ptr(out_rows) && !is_device_ptr(out_dist) && !is_device_ptr(out_count),
                "b2r_query_async: buffers must be host arrays (device buffers: b2r_query only enqueues anyway)");
    B2R_REQUIRE(!filter || ((!filter->allow_bits || is_device_ptr(filter->allow_bits)) && !filter->where),
                "b2r_query_async: only type-mask or device-resident filters");
    std::unique_lock<std::mutex> g(h->mu);
    B2R_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    // one stream per direction: on a single copy stream the next batch's upload would queue behind this batch's
    // download, which waits for this batch's kernels -- and nothing would overlap
    if (!h->copy_stream) B2R_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->copy_stream_out) B2R_CUDA(cudaStreamCreateWithFlags(&h->copy_stream_out, cudaStreamNonBlocking));
    // the slot: a free one, else the older of the two is completed first
    int si = h->aslot[0].ticket == 0 ? 0 : h->aslot[1].ticket == 0 ? 1 : (h->aslot[0].ticket < h->aslot[1].ticket ? 0 : 1);
    b2r_index::AsyncSlot &sl = h->aslot[si];
    if (sl.ticket != 0) {
        B2R_CUDA(cudaEventSynchronize(sl.ev_d2h));
        async_deliver(sl);
    }
    if (!sl.ev_h2d) {
        B2R_CUDA(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
        B2R_CUDA(cudaEventCreateWithFlags(&sl.ev_kernels, cudaEventDisableTiming));
        B2R_CUDA(cudaEventCreateWithFlags(&sl.ev_d2h, cudaEventDisableTiming));
    }
    const size_t q_bytes = (size_t)nq * h->dim * 4, nk = (size_t)nq * k, o_bytes = nk * 12 + (size_t)nq * 4;
    int rc;
    if ((rc = ensure(sl.q_dev, q_bytes)) != B2R_OK) return rc;
    if ((rc = ensure(sl.o_dev, o_bytes)) != B2R_OK) return rc;
    if ((rc = pinned_grow(&sl.out_host, &sl.out_bytes, o_bytes)) != B2R_OK) return rc;
    const void *src = q;
    if (!is_pinned_host(q)) {                 // pageable: through the pinned mirror (a host memcpy, hidden behind the GPU's work)
        if ((rc = pinned_grow(&sl.in_host, &sl.in_bytes, q_bytes)) != B2R_OK) return rc;
        std::memcpy(sl.in_host, q, q_bytes);
        src = sl.in_host;
    }
    B2R_CUDA(cudaMemcpyAsync(sl.q_dev.p, src, q_bytes, cudaMemcpyHostToDevice, h->copy_stream));
    B2R_CUDA(cudaEventRecord(sl.ev_h2d, h->copy_stream));
    B2R_CUDA(cudaStreamWaitEvent(s, sl.ev_h2d, 0));
    char *ob = (char *)sl.o_dev.p;
    rc = query_locked(h, (const float *)sl.q_dev.p, nq, k, filter, (int64_t *)ob, (float *)(ob + nk * 8), nullptr,
                      (int32_t *)(ob + nk * 12), stream);
    if (rc != B2R_OK) return rc;
    B2R_CUDA(cudaEventRecord(sl.ev_kernels, s));
    B2R_CUDA(cudaStreamWaitEvent(h->copy_stream_out, sl.ev_kernels, 0));
    B2R_CUDA(cudaMemcpyAsync(sl.out_host, sl.o_dev.p, o_bytes, cudaMemcpyDeviceToHost, h->copy_stream_out));
    B2R_CUDA(cudaEventRecord(sl.ev_d2h, h->copy_stream_out));
    sl.ticket = h->next_ticket++;
    sl.u_rows = out_rows; sl.u_dist = out_dist; sl.u_count = out_count; sl.nq = nq; sl.k = k;
    *ticket_out = sl.ticket;
    return B2R_OK;
}

extern "C" int b2r_wait(b2r_handle h, uint64_t ticket) {
    B2R_REQUIRE(h, "b2r_wait: NULL handle");
    std::unique_lock<std::mutex> g(h->mu);
    B2R_CUDA(cudaSetDevice(h->device));
    for (int si = 0; si < 2; ++si) {
        b2r_index::AsyncSlot &sl = h->aslot[si];
        if (sl.ticket != ticket || ticket == 0) continue;
        cudaEvent_t ev = sl.ev_d2h;
        g.unlock();                               // other threads may enqueue while this one waits
        B2R_CUDA(cudaEventSynchronize(ev));
        g.lock();
        if (sl.ticket == ticket) async_deliver(sl);   // (a third call may have completed it meanwhile)
        return B2R_OK;
    }
    return B2R_OK;                                // already delivered
}

// ---------------------------------------------------------------------------------
// cross-shard merge (K4)
// ---------------------------------------------------------------------------------
namespace {
constexpr int MERGE_THREADS = 256;
// one CTA per query: rank every gathered candidate by counting (<= nshards*k <= 2048)
__global__ void __launch_bounds__(MERGE_THREADS)
merge_shards_kernel(const char *in_rows_b, const char *in_dist_b, const char *in_count_b, long long stride_rows,
                    long long stride_dist, long long stride_count, int nshards, int nq, int k, long long *out_rows,
                    float *out_dist, int *out_count) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int qi = blockIdx.x;
    const int total = nshards * k;
    double *sd = reinterpret_cast<double *>(sm);
    long long *sr = reinterpret_cast<long long *>(sd + total);
    __shared__ int s_valid;
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    int my_valid = 0;
    for (int c = threadIdx.x; c < total; c += MERGE_THREADS) {
        int sh = c / k, i = c % k;
        const long long *in_rows = reinterpret_cast<const long long *>(in_rows_b + (size_t)sh * stride_rows);
        const double *in_dist = reinterpret_cast<const double *>(in_dist_b + (size_t)sh * stride_dist);
        const int *in_count = reinterpret_cast<const int *>(in_count_b + (size_t)sh * stride_count);
        bool ok = i < in_count[qi];
        size_t src = (size_t)qi * k + i;
        sd[c] = ok ? in_dist[src] : __longlong_as_double(0x7ff0000000000000ll);
        sr[c] = ok ? in_rows[src] : -1;
        my_valid += ok ? 1 : 0;
    }
    atomicAdd(&s_valid, my_valid);
    __syncthreads();
    const int cnt = min(s_valid, k);
    for (int c = threadIdx.x; c < total; c += MERGE_THREADS) {
        const long long r = sr[c];
        if (r < 0) continue;
        const double d = sd[c];
        int rank = 0;
        for (int j = 0; j < total; ++j) {
            const long long rj = sr[j];
            const double dj = sd[j];
            rank += (rj >= 0 && (dj < d || (dj == d && rj < r))) ? 1 : 0;
        }
        if (rank < k) {
            out_rows[(size_t)qi * k + rank] = r;
            out_dist[(size_t)qi * k + rank] = (float)d;
        }
    }
    for (int t = cnt + threadIdx.x; t < k; t += MERGE_THREADS) {
        out_rows[(size_t)qi * k + t] = -1;
        out_dist[(size_t)qi * k + t] = __int_as_float(0x7f800000);
    }
    if (threadIdx.x == 0) out_count[qi] = cnt;
}
}  // namespace

// shard s of each input starts s * (its stride) bytes after the base pointer
static int merge_launch(const char *rows_b, const char *dist_b, const char *count_b, long long stride_rows,
                        long long stride_dist, long long stride_count, int nshards, int nq, int k, int64_t *out_rows,
                        float *out_dist, int32_t *out_count, int device, void *stream) {
    B2R_REQUIRE(nshards >= 1 && nq >= 1 && k >= 1, "b2r_merge_shards: bad sizes");
    B2R_REQUIRE((size_t)nshards * k <= 8192, "b2r_merge_shards: nshards*k must be <= 8192");
    B2R_CUDA(cudaSetDevice(device));
    size_t smem = (size_t)nshards * k * 16;
    if (smem > 48 * 1024)
        B2R_CUDA(cudaFuncSetAttribute(merge_shards_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_shards_kernel<<<nq, MERGE_THREADS, smem, (cudaStream_t)stream>>>(rows_b, dist_b, count_b, stride_rows, stride_dist,
                                                                           stride_count, nshards, nq, k, (long long *)out_rows,
                                                                           out_dist, out_count);
    B2R_CUDA(cudaGetLastError());
    return B2R_OK;
}

extern "C" int b2r_merge_shards(const int64_t *in_rows, const double *in_dist64, const int32_t *in_count,
                                int nshards, int nq, int k, int64_t *out_rows, float *out_dist,
                                int32_t *out_count, int device, void *stream) {
    B2R_REQUIRE(in_rows && in_dist64 && in_count && out_rows && out_dist && out_count, "b2r_merge_shards: NULL argument");
    return merge_launch((const char *)in_rows, (const char *)in_dist64, (const char *)in_count, (long long)nq * k * 8,
                        (long long)nq * k * 8, (long long)nq * 4, nshards, nq, k, out_rows, out_dist, out_count, device, stream);
}

extern "C" int b2r_merge_shards_packed(const void *packed, int64_t shard_stride, int64_t off_rows, int64_t off_dist64,
                                       int64_t off_count, int nshards, int nq, int k, int64_t *out_rows, float *out_dist,
                                       int32_t *out_count, int device, void *stream) {
    B2R_REQUIRE(packed && out_rows && out_dist && out_count, "b2r_merge_shards_packed: NULL argument");
    B2R_REQUIRE(shard_stride > 0 && off_rows % 8 == 0 && off_dist64 % 8 == 0 && off_count % 4 == 0,
                "b2r_merge_shards_packed: bad layout");
    const char *b = (const char *)packed;
    return merge_launch(b + off_rows, b + off_dist64, b + off_count, shard_stride, shard_stride, shard_stride, nshards, nq, k,
                        out_rows, out_dist, out_count, device, stream);
}
