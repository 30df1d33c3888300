// Last-CTA epilogue shared by the scoring kernels: merge the per-CTA candidate lists,
// re-rank the KP survivors exactly (fp64 accumulate over the stored fp32 rows), prove
// that no row outside the candidate set can belong to the top-k (certificate), and
// write Chroma-ordered results.  Runs in ONE CTA of FIN_THREADS threads.
#pragma once
#include "common.cuh"

namespace b2r {

constexpr int FIN_THREADS = 256;
constexpr int FIN_WARPS = FIN_THREADS / 32;

// Which FIN_THREADS threads run the finalize code: a whole CTA of that size (FinCta<0>), or a 256-thread group
// of a larger CTA, threads OFFSET .. OFFSET+255, synchronised by named barrier 1.  (Running the finalize as the
// tail of K3 was measured and dropped: a finalize is a ~10 us latency chain per query, and a K3 CTA would walk
// its 1-8 queries one after the other, while the separate launch runs all queries side by side.)
template <int OFFSET>
struct FinCta {
    static __device__ __forceinline__ int tid() { return (int)threadIdx.x - OFFSET; }
    static __device__ __forceinline__ void sync() {
        if (OFFSET == 0) __syncthreads();
        else asm volatile("bar.sync 1, %0;" ::"n"(FIN_THREADS) : "memory");
    }
};

// Row-sharded corpus, fused exchange (b2r_query_push): whoever emits a query's final list also stores it -- global row,
// fp64 distance, count -- straight into every rank's mailbox over NVLink (xchg.cu describes the mailboxes), so the lists
// travel while the rest of the batch is still being finalized and no separate exchange kernel or collective runs.
struct PushParams {
    char *box[8];                       // every rank's mailbox as mapped here (box[rank] = the local one)
    unsigned long long lists_off;       // this call's slot, this rank's block of lists: slot * slot_bytes + rank * nq_max * entry_bytes
    unsigned entry_bytes, k_max;        // one query's entry: rows[k_max] i64 | dist[k_max] f64 | count i32 (+pad)
    int world;                          // 0 = no push.  (The arrival words are written by the NEXT kernel on the stream -- xchg.cuh,
                                        // XchgFlags -- once the grids that pushed have completed: no fence, no exit ticket here.)
};

struct FinalizeParams {
    const float *master;        // [rows, dp] fp32 stored rows, or nullptr (bf16-only corpus)
    const uint4 *corpus;        // [rows, dp] bf16 stored rows
    const float *q;             // [nq, dp] fp32 prepared queries (normalised for cosine)
    const float *max_norm2;     // device [2]: max |x|^2 over stored rows, max |x - bf16(x)|^2 (0 without an fp32 master)
    const double *q_eps;        // [nq][2]: error bound eps of this query's scan scores, |q|^2 (query preparation, ingest.cuh)
    int dp, space, k;
    long long row_base;         // added to local rows on output (shard offset)
    long long *out_rows;        // [nq, k]
    float *out_dist;            // [nq, k]
    double *out_dist64;         // [nq, k] or nullptr
    int *out_count;             // [nq]
    int *need_ctl;              // [0] = number of queries whose certificate failed (this call; the query preparation clears it)
    int *need_list;             // [nq]: those queries, in arrival order; the exact scan (K5) redoes them
    PushParams push;            // world = 0 unless the call is b2r_query_push
};



// exact distance between prepared query and stored row, fp64 accumulation, whole warp
__device__ __forceinline__ double exact_distance_warp(const FinalizeParams &p, const float *qv,
                                                      unsigned row, int lane) {
    double acc = 0.0;
    const int dp = p.dp;
    if (p.master) {
        const float4 *xr = reinterpret_cast<const float4 *>(p.master + (size_t)row * dp);
        const float4 *q4 = reinterpret_cast<const float4 *>(qv);
#pragma unroll 4
        for (int c = lane; c < dp / 4; c += 32) {      // unrolled so the row's loads are all in flight before the fp64 chain
            float4 x = __ldcg(xr + c);
            float4 q = q4[c];
            if (p.space == 0) {
                double a = (double)q.x - (double)x.x, b = (double)q.y - (double)x.y;
                double cc = (double)q.z - (double)x.z, d = (double)q.w - (double)x.w;
                acc = fma(a, a, acc); acc = fma(b, b, acc); acc = fma(cc, cc, acc); acc = fma(d, d, acc);
            } else {
                acc = fma((double)q.x, (double)x.x, acc); acc = fma((double)q.y, (double)x.y, acc);
                acc = fma((double)q.z, (double)x.z, acc); acc = fma((double)q.w, (double)x.w, acc);
            }
        }
    } else {
        const uint4 *xr = p.corpus + (size_t)row * (dp / 8);
#pragma unroll 4
        for (int c = lane; c < dp / 8; c += 32) {
            uint4 w = __ldcg(xr + c);
            const float *qq = qv + c * 8;
            unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double x0 = (double)bf16lo(ww[i]), x1 = (double)bf16hi(ww[i]);
                double q0 = (double)qq[2 * i], q1 = (double)qq[2 * i + 1];
                if (p.space == 0) {
                    double a = q0 - x0, b = q1 - x1;
                    acc = fma(a, a, acc); acc = fma(b, b, acc);
                } else {
                    acc = fma(q0, x0, acc); acc = fma(q1, x1, acc);
                }
            }
        }
    }
    acc = warp_sum(acc);
    return p.space == 0 ? acc : 1.0 - acc;
}

// The same with G lanes per row (G = 8 or 16): 32/G rows per warp in flight, for the short candidate prefixes of
// the re-rank.  `sub` = lane % G; every lane of the group returns the distance.
template <int G>
__device__ __forceinline__ double exact_distance_group(const FinalizeParams &p, const float *qv, unsigned row, int sub) {
    double a0 = 0.0, a1 = 0.0;
    const int dp = p.dp;
    if (p.master) {
        const float4 *xr = reinterpret_cast<const float4 *>(p.master + (size_t)row * dp);
        const float4 *q4 = reinterpret_cast<const float4 *>(qv);
#pragma unroll 6
        for (int c = sub; c < dp / 4; c += G) {
            float4 x = __ldcg(xr + c);
            float4 q = q4[c];
            if (p.space == 0) {
                double a = (double)q.x - (double)x.x, b = (double)q.y - (double)x.y;
                double cc = (double)q.z - (double)x.z, d = (double)q.w - (double)x.w;
                a0 = fma(a, a, a0); a1 = fma(b, b, a1); a0 = fma(cc, cc, a0); a1 = fma(d, d, a1);
            } else {
                a0 = fma((double)q.x, (double)x.x, a0); a1 = fma((double)q.y, (double)x.y, a1);
                a0 = fma((double)q.z, (double)x.z, a0); a1 = fma((double)q.w, (double)x.w, a1);
            }
        }
    } else {
        const uint4 *xr = p.corpus + (size_t)row * (dp / 8);
#pragma unroll 3
        for (int c = sub; c < dp / 8; c += G) {
            uint4 w = __ldcg(xr + c);
            const float *qq = qv + c * 8;
            unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double x0 = (double)bf16lo(ww[i]), x1 = (double)bf16hi(ww[i]);
                double q0 = (double)qq[2 * i], q1 = (double)qq[2 * i + 1];
                if (p.space == 0) {
                    double a = q0 - x0, b = q1 - x1;
                    a0 = fma(a, a, a0); a1 = fma(b, b, a1);
                } else {
                    a0 = fma(q0, x0, a0); a1 = fma(q1, x1, a1);
                }
            }
        }
    }
    double acc = a0 + a1;
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
    return p.space == 0 ? acc : 1.0 - acc;
}

// Sort `n` exact keys held in smem (ex[0..n)) by rank computation and emit results.
// Called by all FIN_THREADS threads.  Returns (via smem slot) nothing; writes outputs.
template <class C = FinCta<0>>
__device__ __forceinline__ void emit_sorted(const FinalizeParams &p, int qi, const KeyD *ex, int n,
                                            KeyD *kth_out /* smem, rank k-1 entry or worst */) {
    const int k = p.k;
    const int cnt = n < k ? n : k;
    const int tid = C::tid();
    for (int t = tid; t < n; t += FIN_THREADS) {
        KeyD me = ex[t];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += KeyD::better(ex[j], me) ? 1 : 0;
        if (rank < k) {
            p.out_rows[(size_t)qi * k + rank] = p.row_base + (long long)me.row;
            p.out_dist[(size_t)qi * k + rank] = (float)me.d;
            if (p.out_dist64) p.out_dist64[(size_t)qi * k + rank] = me.d;
            if (p.push.world) {                          // the same entry into every rank's mailbox (a redone query overwrites its entry)
                const size_t e_off = p.push.lists_off + (size_t)qi * p.push.entry_bytes;
                for (int w = 0; w < p.push.world; ++w) {
                    char *e = p.push.box[w] + e_off;
                    reinterpret_cast<long long *>(e)[rank] = p.row_base + (long long)me.row;
                    reinterpret_cast<double *>(e + (size_t)p.push.k_max * 8)[rank] = me.d;
                }
            }
        }
        if (rank == k - 1) *kth_out = me;
    }
    for (int t = cnt + tid; t < k; t += FIN_THREADS) {
        p.out_rows[(size_t)qi * k + t] = -1;
        p.out_dist[(size_t)qi * k + t] = __int_as_float(0x7f800000);
        if (p.out_dist64) p.out_dist64[(size_t)qi * k + t] = __longlong_as_double(0x7ff0000000000000ll);
    }
    if (tid == 0) p.out_count[qi] = cnt;
    if (p.push.world && tid < p.push.world)
        *reinterpret_cast<int *>(p.push.box[tid] + p.push.lists_off + (size_t)qi * p.push.entry_bytes + (size_t)p.push.k_max * 16) = cnt;
}

// Stages shared by every scoring path once the candidate set is known (sm_keys: nvalid candidates,
// best scan score first).  All FIN_THREADS threads call; sm_q must already hold the prepared query.
//
// Error bound.  scan score = q~.x~ (+bias) with x~ = bf16(x) and, for K3, q~ = bf16(q); the exact score is
// q.x (+bias).  q~.x~ - q.x = (q~-q).x~ + q.(x~-x), so
//     |scan - exact| <= |q~-q| (max|x| + max|x~-x|) + |q| max|x~-x| + accumulation slop =: eps
// with max|x~-x| and |q~-q| MEASURED (ingest / query preparation), not the worst case 2^-8.
//
// 1. A candidate whose scan score is below (k-th best scan score - 2 eps) has an exact score strictly
//    below the exact scores of k other candidates: it is skipped, only the prefix that can still reach
//    the top-k is re-ranked (fp64 accumulate over the stored rows, one warp per candidate).
// 2. Certificate: every row outside the candidate set has scan score <= T, hence exact score <= T + eps.
//    If the k-th exact candidate beats that, no outsider belongs to the top-k; otherwise the query goes
//    on the exact-scan work list.  T = -inf: there are no outsiders.
template <class C = FinCta<0>>
__device__ __forceinline__ void finalize_candidates(const FinalizeParams &p, int qi, const KeyS *sm_keys, int nvalid,
                                                    float T, KeyD *sm_ex, const float *sm_q, KeyD *sm_misc) {
    const int tid = C::tid();
    const int lane = tid & 31, warp = tid >> 5;
    // start pulling the likeliest rows towards L2 (the re-rank prefix is rarely longer than this): a cold row
    // costs a DRAM access and usually a TLB miss, the longest single wait of the whole finalize
    constexpr int G = 16, PER_WARP = 32 / G, EARLY = FIN_WARPS * PER_WARP;
    const int row_bytes = p.master ? p.dp * 4 : p.dp * 2;
    const int lines = (row_bytes + 127) / 128;
    auto prefetch_rows = [&](int first, int last) {
        for (int i = first * lines + tid; i < last * lines; i += FIN_THREADS) {
            const unsigned row = sm_keys[i / lines].row();
            const char *base = p.master ? reinterpret_cast<const char *>(p.master + (size_t)row * p.dp)
                                        : reinterpret_cast<const char *>(p.corpus + (size_t)row * (p.dp / 8));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)(i % lines) * 128));
        }
    };
    prefetch_rows(0, min(nvalid, EARLY));
    // eps and |q|^2 come from the query preparation; every warp derives the prefix length on its own (same
    // inputs, same arithmetic -> same value), so no barrier stands between the ranking and the first row fetch
    const double eps = __ldg(p.q_eps + 2 * qi), qn2 = __ldg(p.q_eps + 2 * qi + 1);
    if (tid == 0) sm_misc[0] = KeyD::worst();
    // ---- prefix that can still reach the top-k (the list is sorted, so the survivors are a prefix) ----
    int m = nvalid;
    if (nvalid > p.k) {
        const double cut = (double)sm_keys[p.k - 1].score() - 2.0 * eps;
        m = 0;
        for (int i0 = 0; i0 < nvalid; i0 += 32) {
            const int i = i0 + lane;
            m += __popc(__ballot_sync(FULL_MASK, i < nvalid && (double)sm_keys[i].score() >= cut));
        }
    }
    if (m > EARLY) prefetch_rows(EARLY, m);     // several rounds per warp: the rest of the prefix too
    for (int c0 = warp * PER_WARP; c0 < m; c0 += FIN_WARPS * PER_WARP) {     // warp-uniform trip count
        const int c = c0 + lane / G;
        const unsigned row = sm_keys[c < m ? c : m - 1].row();
        const double d = exact_distance_group<G>(p, sm_q, row, lane % G);
        if (c < m && lane % G == 0) sm_ex[c] = KeyD::make(d, row);
    }
    C::sync();
    emit_sorted<C>(p, qi, sm_ex, m, &sm_misc[0]);
    C::sync();
    if (tid == 0 && T > -INFINITY) {
        const KeyD kth = sm_misc[0];
        double s_k;
        if (!kth.valid()) s_k = -1e300;                       // fewer than k candidates although rows were rejected
        else if (p.space == 0) s_k = 0.5 * (qn2 - kth.d);
        else s_k = 1.0 - kth.d;
        if (!(s_k - eps > (double)T)) p.need_list[atomicAdd(&p.need_ctl[0], 1)] = qi;
    }
    C::sync();
}

// Tree-merge the per-warp lists of a CTA: after the call warp 0 holds the CTA's best KP and
// has stored them rank-ordered at sm[0, KP).  sm must hold FIN_WARPS*KP keys.  All threads call.
template <class K, int EPL, class C = FinCta<0>>
__device__ __forceinline__ void cta_tree_merge(WarpList<K, EPL> &wl, K *sm, int warp, int lane) {
    constexpr int KP = 32 * EPL;
#pragma unroll
    for (int half = FIN_WARPS / 2; half >= 1; half >>= 1) {
        if (warp >= half && warp < 2 * half) wl.store(sm + warp * KP, lane);
        C::sync();
        if (warp < half) wl.template merge_bitonic<false>(sm + (warp + half) * KP, lane);
        C::sync();
    }
    if (warp == 0) wl.store(sm, lane);
    C::sync();
}

// Merge nlists rank-ordered KeyS lists of length KP (global memory, written by other
// CTAs -> read with ld.cg), re-rank, certify, emit.  The lists are staged through shared
// memory with coalesced loads (one DRAM/L2 latency per chunk instead of one per list).
// smem: stage[stage_keys] KeyS (stage_keys a multiple of KP, >= FIN_WARPS*KP), sm_ex[KP] KeyD,
// sm_q[dp] float, sm_misc[4] KeyD.
template <int EPL>
__device__ void finalize_scored_query(const FinalizeParams &p, int qi, const KeyS *lists, int nlists,
                                      size_t list_stride /* keys between consecutive lists */,
                                      KeyS *stage, int stage_keys, KeyD *sm_ex, float *sm_q, KeyD *sm_misc) {
    constexpr int KP = 32 * EPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    KeyS *sm_keys = stage;

    for (int i = threadIdx.x; i < p.dp; i += FIN_THREADS) sm_q[i] = p.q[(size_t)qi * p.dp + i];

    // ---- stage 1: every warp folds its share of the lists, chunk by chunk ----
    WarpList<KeyS, EPL> wl; wl.init();
    const int lists_per_chunk = stage_keys / KP;
    for (int c0 = 0; c0 < nlists; c0 += lists_per_chunk) {
        const int nl = min(lists_per_chunk, nlists - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nl * KP; i += FIN_THREADS) {
            const int li = i / KP, r = i % KP;
            stage[i] = KeyS::load_cg(lists + (size_t)(c0 + li) * list_stride + r);
        }
        __syncthreads();
        for (int li = warp; li < nl; li += FIN_WARPS) {
            const KeyS head = stage[li * KP];
            if (head.valid() && wl.accepts(head)) wl.template merge_bitonic<false>(stage + li * KP, lane);
        }
    }
    __syncthreads();
    // ---- stage 2: fold the warps ----
    cta_tree_merge<KeyS, EPL>(wl, stage, warp, lane);
    int nvalid = 0;
    for (int i = 0; i < KP; ++i) nvalid += sm_keys[i].valid() ? 1 : 0;   // tiny, uniform

    // every row outside the list scored <= T in the scan (T = -inf: the list holds every passing row)
    const float T = nvalid == KP ? sm_keys[KP - 1].score() : -INFINITY;
    finalize_candidates(p, qi, sm_keys, nvalid, T, sm_ex, sm_q, sm_misc);
}

}  // namespace b2r
