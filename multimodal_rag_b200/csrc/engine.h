// Host-side engine state behind the C ABI (include/b2r.h) and the launcher interfaces
// each kernel translation unit exports.  Not part of the ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b2r.h"
#include "common.cuh"
#include "finalize.cuh"
#include "scan.cuh"
#include "exact.cuh"

namespace b2r {

void set_error(const std::string &msg);

#define B2R_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            b2r::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));             \
            return _e == cudaErrorMemoryAllocation ? B2R_ENOMEM : B2R_ECUDA;                \
        }                                                                                   \
    } while (0)

// launch `k` allowing programmatic dependent launch after the previous kernel on the stream
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, std::forward<Args>(args)...);
}

// ---- launchers implemented in the kernel TUs ----
// returns false when no kernel is built for (dp, nq_group, epl)
bool scan_supported(int dp);
// max CTAs that can be co-resident for this instantiation (0 if unsupported)
int scan_max_grid(int dp, int nq_group, int epl, int sm_count);
cudaError_t scan_launch(int dp, int nq_group, int epl, const ScanParams &p, int grid, cudaStream_t s);
int scan_tile_rows(int dp);

int exact_max_grid(int epl, int dp, int sm_count);
int exact_group(int epl, int dp);   // queries K5 scores per corpus pass
cudaError_t exact_launch(int epl, const ExactParams &p, int grid, cudaStream_t s, size_t rider_smem = 0);
size_t exact_smem_limit(int epl, int dp);      // dynamic shared memory an exact-scan launch may ask for

// K3 (gemm_kernels.cu): tcgen05 batched scoring
struct GemmParams;
struct UnionParams;
bool gemm_supported(int dp, int k);
int gemm_list_len(int k);          // per-thread list length L for n_results = k; 0 = pool mode (32 < k <= 128); -1 = unsupported
int gemm_tile_rows(int dp);        // corpus rows per MMA tile (BN)
int gemm_encode_map(CUtensorMap *out, const void *base, int dp, uint64_t rows, int box_rows);
int gemm_max_pairs(int dp, int L, bool bias);   // co-resident CTA pairs of the cta_group::2 form (0 = unusable)
cudaError_t gemm_launch(int dp, int L, bool bias, bool pair, int bm, const CUtensorMap &tm_q, const CUtensorMap &tm_x,
                        const GemmParams &p, cudaStream_t s);
cudaError_t pass_bits_launch(const uint8_t *type_code, unsigned long long type_mask, const uint32_t *allow_bits,
                             unsigned n, unsigned n_words, uint32_t *out, int sm_count, cudaStream_t s);
cudaError_t finalize_union_launch(int epl, const FinalizeParams &fin, const UnionParams &u, int q0, int nq, cudaStream_t s);

// xchg.cu: reserves the next mailbox slot of `x` for a fused push of nq x k lists and fills what the kernels need; the acks the
// first kernel must see before anything is stored are returned as (words, count, value)
struct FusedCall {
    PushParams push;               // where this call's kernels store their lists
    XchgFlags flags;               // flag words the previous fused call left for this call's first kernel
    XchgDev rider;                 // the earlier batch merged inside this call's last kernel (nq = 0: none) ...
    size_t rider_smem;
    bool merge_after;              // ... or, when its lists do not fit there, by the stand-alone merge kernel behind the call
    const unsigned *wait_words; int wait_n; unsigned wait_val;     // acks the first kernel must see before anything is stored
};
int xchg_begin_fused(b2r_xchg *x, int device, int nq, int k, FusedCall *c, int64_t *merge_rows, float *merge_dist,
                     int32_t *merge_count, size_t rider_smem_limit);
int xchg_launch_merge(b2r_xchg *x, XchgDev job, cudaStream_t s);

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

}  // namespace b2r

struct b2r_index {
    int dim = 0, dp = 0, space = 0, device = 0;
    uint32_t flags = 0;
    int64_t rows = 0, live = 0, capacity = 0;
    int64_t row_base = 0;
    int sm_count = 0;
    int path = 0;
    int64_t n_queries = 0, n_launches = 0;

    // corpus (device)
    uint4 *corpus = nullptr;        // bf16 [capacity, dp]
    float *master = nullptr;        // fp32 [capacity, dp] unless B2R_FLAG_NO_F32_MASTER
    float *bias = nullptr;          // [capacity], l2 only
    uint8_t *type_code = nullptr;   // [capacity]
    float *max_norm2 = nullptr;     // [2]: max |x|^2, max |x - bf16(x)|^2 over the stored rows
    unsigned long long *counters = nullptr;   // [0] = rows killed by tombstone, [1] = exact fallbacks
    int32_t *cols[B2R_MAX_COLUMNS] = {};      // dictionary-encoded metadata columns [capacity], -1 = key absent; allocated on first use

    // scratch (device), grown on demand
    b2r::DevBuf x_stage, t_stage, q_raw, q_prep, allow, rows_stage, gather_out;
    b2r::DevBuf o_pack, need_list;  // packed outputs of a query with host result arrays
    void *o_host = nullptr; size_t o_host_bytes = 0;   // pinned mirror of o_pack
    int *need_ctl = nullptr;        // [4]: failed-certificate count, exit ticket (reset by K5 itself)
    b2r::DevBuf scan_lists, exact_lists;
    b2r::DevBuf q_bf16, q_err, pass_bits, gthr, gemm_lists, gemm_regions, gemm_samples;   // K3 scratch
    b2r::DevBuf where_lut, where_bits, col_stage;   // compiled-clause tables, clause bitmap, column upload staging
    b2r::DevBuf q_eps;              // [nq][2] fp64: per-query error bound and |q|^2 (query preparation)

    // TMA tensor maps (K3), re-encoded when the buffer they describe moves or grows
    CUtensorMap tm_corpus, tm_corpus_half, tm_query;     // corpus boxes of a whole tile / of half a tile (CTA pairs)
    const void *tm_corpus_base = nullptr; int64_t tm_corpus_rows = -1;
    const void *tm_query_base = nullptr; int64_t tm_query_rows = -1; int tm_query_box = 0;
    // the pass bitmap is reused while (rows, type mask, tombstones) are unchanged and no allow bitmap is given
    int64_t mut_gen = 0, pb_gen = -1, pb_rows = -1; unsigned long long pb_mask = 0; const void *pb_buf = nullptr; int pb_bn = 0;
    // ... and so is a compiled clause's bitmap: filter_key = hash of the clause (0 = no clause), kept with the bitmaps it produced
    uint64_t pb_key = 0, wb_key = 0; int64_t wb_gen = -1, wb_rows = -1;
    unsigned *tickets = nullptr;    // [1 + EXACT_MAX_SLOTS]: scan ticket, K5's per-group arrival tickets

    // optional per-kernel timing (bench.py roofline): CUDA events recorded around the scoring
    // kernel launches on the caller's stream, resolved lazily by b2r_kernel_time_ms
    bool timing = false;
    bool no_seed = false;
    int seed_min_batch = 0, seed_tiles_override = 0;
    // development knobs (environment, read by b2r_create): B2R_SEED_WAIT_NS overrides the wait budget of K3's seeding phase
    // (1 = never wait), B2R_DELAY_US makes every third slice post late, B2R_POOL_SAMPLE_DIV = pool mode samples 1/div of the
    // shard (default 32), B2R_TRACE=1 records per-CTA phase timestamps of the last K3 launch (b2r_debug_trace)
    unsigned long long seed_wait_ns = 0;
    int delay_us = 0, pool_sample_div = 0;     // 0 = by shard size (1/32, 1/64 from 8M rows on)
    bool no_pair = false;
    bool seed_rank_l = false;       // B2R_SEED_RANK_L=1: the L-th best sample seeds K3's bound (the form before the k-th best was used)
    bool no_dyn = false;
    bool no_bm64 = false;           // B2R_NO_BM64=1: 128-query blocks even for batches of at most 64            // B2R_NO_DYN=1: static slices only (no dynamic tile hand-out)           // B2R_NO_PAIR=1: never use the cta_group::2 form of K3
    b2r::DevBuf trace;
    bool trace_on = false; int trace_ctas = 0, trace_mode = 1;
    int timing_stage = 0;           // which launch the events bracket: 0 scoring (default); B2R_TIME_STAGE=1, 4, 5 (development):
                                    // 1 prepare, 4 finalize, 5 exact fix-up
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending, ev_free;
    double scoring_ms = 0.0;
    int64_t scoring_launches = 0;

    // pipelined queries with host buffers (b2r_query_async / b2r_wait): two slots, each with its own device staging
    // and pinned mirrors, copies on an internal stream so that the transfers of one call overlap the kernels of another
    struct AsyncSlot {
        b2r::DevBuf q_dev, o_dev;
        void *in_host = nullptr, *out_host = nullptr;       // pinned mirrors (in: pageable queries are staged through it)
        size_t in_bytes = 0, out_bytes = 0;
        cudaEvent_t ev_h2d = nullptr, ev_kernels = nullptr, ev_d2h = nullptr;
        uint64_t ticket = 0;                                 // 0 = free
        int64_t *u_rows = nullptr; float *u_dist = nullptr; int32_t *u_count = nullptr;   // the caller's host arrays
        int nq = 0, k = 0;
    } aslot[2];
    cudaStream_t copy_stream = nullptr, copy_stream_out = nullptr;   // uploads / downloads
    uint64_t next_ticket = 1;

    std::mutex mu;
};
