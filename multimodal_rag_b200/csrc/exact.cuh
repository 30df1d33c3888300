// K5: exact fp64 scan.  Reads the stored fp32 rows (or the bf16 rows when the handle
// keeps no fp32 master), accumulates every distance in fp64 and selects on
// (distance, row) directly, so its answer needs no certificate.  It is the always-correct
// fallback for queries whose bf16 candidate certificate failed (fin.need_list), and the forced
// path for shapes the fast kernels are not built for.
//
// Work items (queries) are taken G at a time: a warp loads a stored row ONCE and scores it against the G
// queries of the group (fp64 copies of the queries sit in shared memory), so a corpus pass costs the same HBM
// traffic for G queries as for one.  One launch covers at most `items` work items starting at `first_item`,
// every group of the launch has its own slot of per-CTA lists, and the last CTA to finish a group merges it:
// no CTA ever waits for another one, so the grid need not be co-resident.  A launch whose range of the work
// list is empty costs one L2 read per CTA.
#pragma once
#include "common.cuh"
#include "finalize.cuh"
#include "xchg.cuh"

namespace b2r {

constexpr int EXACT_THREADS = FIN_THREADS;
constexpr int EXACT_WARPS = EXACT_THREADS / 32;
constexpr int EXACT_MAX_SLOTS = 64;       // groups per launch (tickets)
constexpr int EXACT_MAX_G = 4;            // queries per corpus pass

struct ExactParams {
    const uint8_t *type_code;
    const uint32_t *allow_bits;
    unsigned long long type_mask;
    unsigned n;
    int nq;                       // queries in the batch
    int force_all;                // 1: redo every query of the batch; 0: only fin.need_list[0 .. need_ctl[0])
    int first_item, items;        // this launch covers work items [first_item, first_item + items) (items <= EXACT_MAX_SLOTS * G)
    KeyD *cta_lists;              // [slots][G][gridDim.x][KP]
    unsigned *tickets;            // [EXACT_MAX_SLOTS] arrival tickets (0 between launches: the last CTA of a group resets its own)
    long long *n_fallbacks;       // device counter (may be nullptr)
    XchgDev rider;                // b2r_query_push, last launch of the call: the grid merges the batch pushed BEFORE this one on its way
                                  // out (rider.nq = 0: none)
    FinalizeParams fin;
};

// fp64 distances of one stored row against the G prepared queries in shared memory; whole warp.  The fp64 query copies are
// laid out [G][elements per chunk][chunks] -- element e of chunk c at e * chunks + c -- so that the 32 lanes of a warp, which
// own consecutive chunks, read consecutive doubles (conflict-free); [G][dp] in natural order had every lane 32 bytes apart: an
// 8-way bank conflict on each of the 48 loads per row, which bound a corpus pass at ~480 GB/s.
template <int G>
__device__ __forceinline__ void exact_distance_warp_multi(const FinalizeParams &p, const double *qd, unsigned row, int lane,
                                                          double (&out)[G]) {
    const int dp = p.dp;
    double acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = 0.0;
    if (p.master) {
        const float4 *xr = reinterpret_cast<const float4 *>(p.master + (size_t)row * dp);
        const int nc = dp / 4;
#pragma unroll 3
        for (int c = lane; c < nc; c += 32) {
            const float4 x = __ldcg(xr + c);
            const double x0 = (double)x.x, x1 = (double)x.y, x2 = (double)x.z, x3 = (double)x.w;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const double *q = qd + (size_t)g * dp + c;
                const double q0 = q[0], q1 = q[nc], q2 = q[2 * nc], q3 = q[3 * nc];
                if (p.space == 0) {
                    const double a = q0 - x0, b = q1 - x1, cc = q2 - x2, d = q3 - x3;
                    acc[g] = fma(a, a, acc[g]); acc[g] = fma(b, b, acc[g]); acc[g] = fma(cc, cc, acc[g]); acc[g] = fma(d, d, acc[g]);
                } else {
                    acc[g] = fma(q0, x0, acc[g]); acc[g] = fma(q1, x1, acc[g]);
                    acc[g] = fma(q2, x2, acc[g]); acc[g] = fma(q3, x3, acc[g]);
                }
            }
        }
    } else {
        const uint4 *xr = p.corpus + (size_t)row * (dp / 8);
        const int nc = dp / 8;
#pragma unroll 2
        for (int c = lane; c < nc; c += 32) {
            const uint4 w = __ldcg(xr + c);
            const unsigned ww[4] = {w.x, w.y, w.z, w.w};
            double x[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) { x[2 * i] = (double)bf16lo(ww[i]); x[2 * i + 1] = (double)bf16hi(ww[i]); }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const double *q = qd + (size_t)g * dp + c;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const double qi = q[i * nc];
                    if (p.space == 0) { const double a = qi - x[i]; acc[g] = fma(a, a, acc[g]); }
                    else acc[g] = fma(qi, x[i], acc[g]);
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const double s = warp_sum(acc[g]);
        out[g] = p.space == 0 ? s : 1.0 - s;
    }
}

// The accumulation order differs from exact_distance_warp / exact_distance_group only in how the per-lane partial sums
// are formed; every path accumulates in fp64 over fp32 inputs, whose rounding (2^-53 relative per operation) is far below the
// spacing of distinct fp32-derived distances the tests compare (ids bit-exact, distances to 1e-5 relative).
template <int EPL, int G>
__global__ void __launch_bounds__(EXACT_THREADS) exact_topk_kernel(const ExactParams p) {
    constexpr int KP = 32 * EPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dp = p.fin.dp;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KeyD *sm_keys = reinterpret_cast<KeyD *>(smem_raw);                 // [EXACT_WARPS][KP]
    KeyD *sm_misc = sm_keys + EXACT_WARPS * KP;                         // [4]
    double *sm_q = reinterpret_cast<double *>(sm_misc + 4);            // [G][dp] fp64 copies of the group's queries
    __shared__ unsigned s_ticket;
    __shared__ int s_qi[G];

    pdl_wait();
    pdl_trigger();
    const int total = p.force_all ? p.nq : min(__ldcg(&p.fin.need_ctl[0]), p.nq);
    const int first = p.first_item, last = min(total, p.first_item + p.items);
    for (int it0 = first, slot = 0; it0 < last; it0 += G, ++slot) {
        const int ng = min(G, last - it0);
        KeyD *my_lists = p.cta_lists + (size_t)slot * G * gridDim.x * KP;
        __syncthreads();
        if (threadIdx.x < G) s_qi[threadIdx.x] = threadIdx.x < ng ? (p.force_all ? it0 + (int)threadIdx.x : __ldcg(&p.fin.need_list[it0 + threadIdx.x])) : -1;
        __syncthreads();
        const int epc = p.fin.master ? 4 : 8, nch = dp / epc;        // elements per chunk (one 16-byte load of a stored row), chunks
        for (int g = 0; g < G; ++g) {
            const int qi = s_qi[g];
            for (int i = threadIdx.x; i < dp; i += EXACT_THREADS)
                sm_q[(size_t)g * dp + (size_t)(i % epc) * nch + i / epc] = qi >= 0 ? (double)p.fin.q[(size_t)qi * dp + i] : 0.0;
        }
        __syncthreads();

        WarpList<KeyD, EPL> wl[G];
#pragma unroll
        for (int g = 0; g < G; ++g) wl[g].init();
        // every CTA streams one contiguous block of rows (a grid-wide stride made every warp touch a new 2 MB page on each
        // row: TLB-bound at 25M+ rows)
        const unsigned per_cta = (p.n + gridDim.x - 1) / gridDim.x;
        const unsigned r_begin = blockIdx.x * per_cta, r_end = min(p.n, r_begin + per_cta);
        for (unsigned row = r_begin + warp; row < r_end; row += EXACT_WARPS) {
            if (!row_passes(row, p.type_code, p.type_mask, p.allow_bits)) continue;   // warp-uniform
            double d[G];
            exact_distance_warp_multi<G>(p.fin, sm_q, row, lane, d);
#pragma unroll
            for (int g = 0; g < G; ++g)
                if (g < ng) wl[g].offer(KeyD::make(d[g], row), lane);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (g >= ng) break;
            __syncthreads();
            cta_tree_merge<KeyD, EPL>(wl[g], sm_keys, warp, lane);
            if (warp == 0) wl[g].store(my_lists + ((size_t)g * gridDim.x + blockIdx.x) * KP, lane);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = atomicAdd(&p.tickets[slot], 1u);
        __syncthreads();
        if (s_ticket == gridDim.x - 1) {
            // ---- last CTA of this group: fold all CTA lists of each of its queries, emit ----
            __threadfence();
            for (int g = 0; g < ng; ++g) {
                WarpList<KeyD, EPL> m; m.init();
                for (unsigned li = warp; li < gridDim.x; li += EXACT_WARPS) {
                    const KeyD *src = my_lists + ((size_t)g * gridDim.x + li) * KP;
                    const KeyD head = KeyD::load_cg(src);
                    if (head.valid() && m.accepts(head)) m.template merge_bitonic<true>(src, lane);
                }
                __syncthreads();
                cta_tree_merge<KeyD, EPL>(m, sm_keys, warp, lane);
                if (threadIdx.x == 0) sm_misc[0] = KeyD::worst();
                __syncthreads();
                int nvalid = 0;
                for (int i = 0; i < KP; ++i) nvalid += sm_keys[i].valid() ? 1 : 0;
                emit_sorted(p.fin, s_qi[g], sm_keys, nvalid, &sm_misc[0]);
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                p.tickets[slot] = 0u;                        // ready for the next launch on this stream
                if (!p.force_all && p.n_fallbacks) atomicAdd((unsigned long long *)p.n_fallbacks, (unsigned long long)ng);
            }
        }
    }
    // Rider: the cross-shard merge of the previous fused batch.  Its lists arrived a whole scan ago and this kernel is on the
    // stream anyway (as the -- normally empty -- certificate fix-up): the exchange adds no launch to a step.  Neither this
    // call's arrival nor the rider's "slot has been read" is published from here: the next kernel on the stream does both
    // (XchgFlags), after this grid has completed, so no fence or exit ticket is needed.
    if (p.rider.nq) {
        __shared__ int s_valid;
        xchg_wait_arrivals(p.rider);
        for (int q = blockIdx.x; q < p.rider.nq; q += gridDim.x) {
            __syncthreads();
            xchg_merge_query(p.rider, q, smem_raw, &s_valid);
        }
    }
}

inline size_t exact_smem_bytes(int EPL, int G, int dp) {
    const int KP = 32 * EPL;
    return sizeof(KeyD) * (EXACT_WARPS * KP + 4) + sizeof(double) * (size_t)G * dp;
}

}  // namespace b2r
