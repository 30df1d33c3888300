// K5: exact fp64 scan.  Reads the stored fp32 rows (or the bf16 rows when the handle
// keeps no fp32 master), accumulates every distance in fp64 and selects on
// (distance, row) directly, so its answer needs no certificate.  It is the always-correct
// fallback for queries whose bf16 candidate certificate failed (fin.need_list), and the forced
// path for shapes the fast kernels are not built for.  One query at a time, whole grid; a launch
// with an empty work list costs one L2 read per CTA.
#pragma once
#include "common.cuh"
#include "finalize.cuh"

namespace b2r {

constexpr int EXACT_THREADS = FIN_THREADS;
constexpr int EXACT_WARPS = EXACT_THREADS / 32;
constexpr int EXACT_MAX_BATCH = 64;

struct ExactParams {
    const uint8_t *type_code;
    const uint32_t *allow_bits;
    unsigned long long type_mask;
    unsigned n;
    int nq;                       // queries in the batch
    int force_all;                // 1: redo every query of the batch; 0: only fin.need_list[0 .. need_ctl[0])
    KeyD *cta_lists;              // [EXACT_MAX_BATCH][gridDim.x][KP]: slot = work item % EXACT_MAX_BATCH
    unsigned *tickets;            // [EXACT_MAX_BATCH] arrival tickets, then [EXACT_MAX_BATCH] slot generations
    long long *n_fallbacks;       // device counter (may be nullptr)
    FinalizeParams fin;
};

template <int EPL>
__global__ void __launch_bounds__(EXACT_THREADS) exact_topk_kernel(const ExactParams p) {
    constexpr int KP = 32 * EPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dp = p.fin.dp;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KeyD *sm_keys = reinterpret_cast<KeyD *>(smem_raw);                 // [EXACT_WARPS][KP]
    KeyD *sm_misc = sm_keys + EXACT_WARPS * KP;                         // [4]
    float *sm_q = reinterpret_cast<float *>(sm_misc + 4);              // [dp]
    __shared__ unsigned s_ticket;

    // Work list: every query (forced) or the compacted list of failed certificates.  The grid is
    // sized to be fully resident, so waiting on another CTA's progress below cannot deadlock.
    pdl_wait();
    pdl_trigger();
    int count = p.force_all ? p.nq : min(__ldcg(&p.fin.need_ctl[0]), p.nq);
    unsigned *slot_gen = p.tickets + EXACT_MAX_BATCH;
    for (int it = 0; it < count; ++it) {
        const int qi = p.force_all ? it : __ldcg(&p.fin.need_list[it]);
        const int slot = it % EXACT_MAX_BATCH;
        const unsigned gen = (unsigned)(it / EXACT_MAX_BATCH);
        KeyD *my_lists = p.cta_lists + (size_t)slot * gridDim.x * KP;
        if (gen > 0) {      // the slot's previous user must have been merged before its lists are overwritten
            if (threadIdx.x == 0)
                while (*reinterpret_cast<volatile unsigned *>(&slot_gen[slot]) < gen) __nanosleep(200);
            __threadfence();
        }
        __syncthreads();
        for (int i = threadIdx.x; i < dp; i += EXACT_THREADS) sm_q[i] = p.fin.q[(size_t)qi * dp + i];
        __syncthreads();

        WarpList<KeyD, EPL> wl; wl.init();
        const unsigned gw = blockIdx.x * EXACT_WARPS + warp, nw = gridDim.x * EXACT_WARPS;
        for (unsigned row = gw; row < p.n; row += nw) {
            if (!row_passes(row, p.type_code, p.type_mask, p.allow_bits)) continue;   // warp-uniform
            double d = exact_distance_warp(p.fin, sm_q, row, lane);
            wl.offer(KeyD::make(d, row), lane);
        }
        __syncthreads();
        cta_tree_merge<KeyD, EPL>(wl, sm_keys, warp, lane);
        if (warp == 0) wl.store(my_lists + (size_t)blockIdx.x * KP, lane);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = atomicAdd(&p.tickets[slot], 1u);
        __syncthreads();
        if (s_ticket == gridDim.x - 1) {
            // ---- last CTA for this query: fold all CTA lists, emit ----
            __threadfence();
            WarpList<KeyD, EPL> m; m.init();
            for (unsigned li = warp; li < gridDim.x; li += EXACT_WARPS) {
                const KeyD *src = my_lists + (size_t)li * KP;
                const KeyD head = KeyD::load_cg(src);
                if (head.valid() && m.accepts(head)) m.template merge_bitonic<true>(src, lane);
            }
            __syncthreads();
            cta_tree_merge<KeyD, EPL>(m, sm_keys, warp, lane);
            if (threadIdx.x == 0) sm_misc[0] = KeyD::worst();
            __syncthreads();
            int nvalid = 0;
            for (int i = 0; i < KP; ++i) nvalid += sm_keys[i].valid() ? 1 : 0;
            emit_sorted(p.fin, qi, sm_keys, nvalid, &sm_misc[0]);
            __syncthreads();
            if (threadIdx.x == 0) {
                p.tickets[slot] = 0u;
                if (!p.force_all && p.n_fallbacks) atomicAdd((unsigned long long *)p.n_fallbacks, 1ull);
                __threadfence();
                *reinterpret_cast<volatile unsigned *>(&slot_gen[slot]) = gen + 1;
            }
        }
        __syncthreads();
    }
    // the work list control and the slot generations are cleared by the next call's query preparation
}

inline size_t exact_smem_bytes(int EPL, int dp) {
    const int KP = 32 * EPL;
    return sizeof(KeyD) * (EXACT_WARPS * KP + 4) + sizeof(float) * dp;
}

}  // namespace b2r
