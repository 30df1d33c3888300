// K2: bandwidth-bound warp-shuffle scan for small query batches (nq <= 4).
//
// The packed bf16 corpus is streamed from HBM exactly once per launch with 16-byte
// read-only loads; every group of LPR lanes owns one row at a time (RPS = 32/LPR rows
// per warp step, U steps in flight), the fp32 queries live in registers, partial dot
// products are reduced with xor-shuffles, and each warp keeps its best KP scores in a
// register-resident sorted list.  CTA lists go to global memory; the last CTA to finish
// (atomic ticket) runs the finalize epilogue (merge, fp64 re-rank, certificate, output),
// so a whole query is one launch.
#pragma once
#include "common.cuh"
#include "finalize.cuh"

namespace b2r {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
static_assert(SCAN_THREADS == FIN_THREADS, "finalize runs in the scan CTA");

struct ScanParams {
    const uint4 *corpus;          // bf16 rows, dp/8 uint4 per row
    const float *bias;            // [n] (-|x|^2/2, l2 space) or nullptr
    const uint8_t *type_code;     // [n]
    const uint32_t *allow_bits;   // nullptr or bitmap
    unsigned long long type_mask;
    unsigned n;                   // rows in this shard
    int q0;                       // first query of this launch (index into fin.q / outputs)
    KeyS *cta_lists;              // [gridDim.x][NQ][KP]
    unsigned *ticket;
    int stage_keys;               // KeyS slots of shared staging for the finalize (multiple of KP, >= 8*KP)
    FinalizeParams fin;
};

template <int LPR, int CPL, int NQ, int EPL, int U>
__global__ void __launch_bounds__(SCAN_THREADS) scan_topk_kernel(const ScanParams p) {
    constexpr int RPS = 32 / LPR;
    constexpr int KP = 32 * EPL;
    constexpr unsigned TILE_ROWS = SCAN_WARPS * RPS * U;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, sl = lane % LPR;
    const int dp = p.fin.dp;
    const int row_chunks = dp / 8;

    extern __shared__ __align__(16) unsigned char smem_raw[];

    pdl_wait();
    pdl_trigger();
    // queries -> registers (this lane only ever touches chunks j*LPR+sl of a row)
    float qr[NQ][CPL][8];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi)
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const float4 *src = reinterpret_cast<const float4 *>(p.fin.q + (size_t)(p.q0 + qi) * dp + (j * LPR + sl) * 8);
            float4 a = src[0], b = src[1];
            qr[qi][j][0] = a.x; qr[qi][j][1] = a.y; qr[qi][j][2] = a.z; qr[qi][j][3] = a.w;
            qr[qi][j][4] = b.x; qr[qi][j][5] = b.y; qr[qi][j][6] = b.z; qr[qi][j][7] = b.w;
        }

    WarpList<KeyS, EPL> wl[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) wl[qi].init();

    const unsigned n = p.n;
    const bool has_bias = p.bias != nullptr;
    for (unsigned long long base = (unsigned long long)blockIdx.x * TILE_ROWS; base < n;
         base += (unsigned long long)gridDim.x * TILE_ROWS) {
        uint4 v[U][CPL];
        float bia[U];
        bool ok[U];
        // ---- issue every load of the tile before touching any of them ----
#pragma unroll
        for (int u = 0; u < U; ++u) {
            unsigned long long row = base + (unsigned)((u * SCAN_WARPS + warp) * RPS + sub);
            unsigned rc = row < n ? (unsigned)row : n - 1;
            const uint4 *src = p.corpus + (size_t)rc * row_chunks + sl;
#pragma unroll
            for (int j = 0; j < CPL; ++j) v[u][j] = ldg_stream(src + j * LPR);
            ok[u] = row < n && row_passes(rc, p.type_code, p.type_mask, p.allow_bits);
            bia[u] = has_bias ? p.bias[rc] : 0.f;
        }
        // ---- dot products ----
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float acc[NQ];
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) acc[qi] = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const unsigned w[4] = {v[u][j].x, v[u][j].y, v[u][j].z, v[u][j].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float lo = bf16lo(w[i]), hi = bf16hi(w[i]);
#pragma unroll
                    for (int qi = 0; qi < NQ; ++qi) {
                        acc[qi] = fmaf(lo, qr[qi][j][2 * i], acc[qi]);
                        acc[qi] = fmaf(hi, qr[qi][j][2 * i + 1], acc[qi]);
                    }
                }
            }
            const unsigned row = (unsigned)(base + (unsigned)((u * SCAN_WARPS + warp) * RPS + sub));
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
                float s = acc[qi];
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
                s += bia[u];
                KeyS mine = KeyS::make(s, row);
                bool want = ok[u] && wl[qi].accepts(mine);
                unsigned hits = __ballot_sync(FULL_MASK, want && sl == 0);
                while (hits) {                       // rare after the first few tiles
                    int src = __ffs(hits) - 1;
                    hits &= hits - 1;
                    KeyS kk = KeyS::shfl(mine, src);
                    wl[qi].offer(kk, lane);
                }
            }
        }
    }

    // ---- CTA merge: tree-fold the warps' lists of each query, publish the CTA list ----
    KeyS *stage = reinterpret_cast<KeyS *>(smem_raw);                // [stage_keys]
    __shared__ unsigned s_ticket;
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        cta_tree_merge<KeyS, EPL>(wl[qi], stage, warp, lane);
        if (warp == 0) wl[qi].store(p.cta_lists + ((size_t)blockIdx.x * NQ + qi) * KP, lane);
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;

    // ---- last CTA: finalize every query of this launch ----
    __threadfence();
    KeyD *sm_ex = reinterpret_cast<KeyD *>(smem_raw + sizeof(KeyS) * (size_t)p.stage_keys);
    KeyD *sm_misc = sm_ex + KP;
    float *sm_q = reinterpret_cast<float *>(sm_misc + 4);
    for (int qi = 0; qi < NQ; ++qi)
        finalize_scored_query<EPL>(p.fin, p.q0 + qi, p.cta_lists + (size_t)qi * KP, gridDim.x,
                                   (size_t)NQ * KP, stage, p.stage_keys, sm_ex, sm_q, sm_misc);
    if (threadIdx.x == 0) *p.ticket = 0u;
}

// shared staging for the finalize: all CTA lists of one query if they fit in 64 KB
inline int scan_stage_keys(int EPL, int grid) {
    const int KP = 32 * EPL;
    long long want = (long long)grid * KP;
    if (want < FIN_WARPS * KP) want = FIN_WARPS * KP;
    if (want > 8192) want = 8192;
    return (int)want;
}
inline size_t scan_smem_bytes(int EPL, int dp, int stage_keys) {
    const int KP = 32 * EPL;
    return sizeof(KeyS) * (size_t)stage_keys + sizeof(KeyD) * (KP + 4) + sizeof(float) * dp;
}

}  // namespace b2r
