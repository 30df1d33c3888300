// Instantiations + launch table for K2 (scan.cuh).  Split from the API translation unit
// so the build can compile kernel families in parallel.
#include <map>
#include <tuple>

#include "engine.h"

namespace b2r {
namespace {

struct ScanShape { int lpr, cpl, u; };

// dp/8 chunks per row -> lanes per row, chunks per lane, steps in flight (0 = no kernel)
constexpr ScanShape shape_of(int dp) {
    switch (dp) {
        case 64:   return {8, 1, 8};
        case 128:  return {16, 1, 8};
        case 256:  return {32, 1, 8};
        case 384:  return {16, 3, 4};   // all-MiniLM-L6-v2
        case 512:  return {32, 2, 6};   // CLIP ViT-B/32 shape
        case 768:  return {32, 3, 4};   // mpnet / BERT-base shape
        case 1024: return {32, 4, 3};
        default:   return {0, 0, 0};
    }
}

typedef void (*scan_fn)(const ScanParams);

}  // namespace

#ifdef B2R_DP
// ---- per-dimension translation unit: nvcc -DB2R_DP=384 ----
#define B2R_CAT2(a, b) a##b
#define B2R_CAT(a, b) B2R_CAT2(a, b)
scan_fn B2R_CAT(scan_lookup_, B2R_DP)(int nq, int epl) {
    constexpr ScanShape S = shape_of(B2R_DP);
    constexpr int LPR = S.lpr, CPL = S.cpl, U = S.u;
    if (nq == 1) {
        if (epl == 1) return scan_topk_kernel<LPR, CPL, 1, 1, U>;
        if (epl == 2) return scan_topk_kernel<LPR, CPL, 1, 2, U>;
        if (epl == 4) return scan_topk_kernel<LPR, CPL, 1, 4, U>;
        if (epl == 8) return scan_topk_kernel<LPR, CPL, 1, 8, U>;
    } else if (nq == 2) {
        if (epl == 1) return scan_topk_kernel<LPR, CPL, 2, 1, U>;
        if (epl == 2) return scan_topk_kernel<LPR, CPL, 2, 2, U>;
    } else if (nq == 4) {
        if (epl == 1) return scan_topk_kernel<LPR, CPL, 4, 1, U>;
        if (epl == 2) return scan_topk_kernel<LPR, CPL, 4, 2, U>;
    }
    return nullptr;
}
}  // namespace b2r
#else
// ---- dispatcher translation unit ----
scan_fn scan_lookup_64(int, int);
scan_fn scan_lookup_128(int, int);
scan_fn scan_lookup_256(int, int);
scan_fn scan_lookup_384(int, int);
scan_fn scan_lookup_512(int, int);
scan_fn scan_lookup_768(int, int);
scan_fn scan_lookup_1024(int, int);

namespace {
scan_fn lookup(int dp, int nq, int epl) {
    switch (dp) {
        case 64:   return scan_lookup_64(nq, epl);
        case 128:  return scan_lookup_128(nq, epl);
        case 256:  return scan_lookup_256(nq, epl);
        case 384:  return scan_lookup_384(nq, epl);
        case 512:  return scan_lookup_512(nq, epl);
        case 768:  return scan_lookup_768(nq, epl);
        case 1024: return scan_lookup_1024(nq, epl);
        default: return nullptr;
    }
}
}  // namespace

bool scan_supported(int dp) { return shape_of(dp).lpr != 0; }

int scan_max_grid(int dp, int nq, int epl, int sm_count) {
    // sized for the largest staging area (8192 keys) so the grid never exceeds residency
    // occupancy is a property of the instantiation: query the runtime once per (dp, nq, epl)
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    auto key = std::make_tuple(dev, dp, nq, epl);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second * sm_count;
    scan_fn f = lookup(dp, nq, epl);
    if (!f) return 0;
    size_t smem = scan_smem_bytes(epl, dp, 8192);
    if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f, SCAN_THREADS, smem) != cudaSuccess) return 0;
    cache[key] = per_sm;
    return per_sm * sm_count;
}

cudaError_t scan_launch(int dp, int nq, int epl, const ScanParams &p, int grid, cudaStream_t s) {
    scan_fn f = lookup(dp, nq, epl);
    if (!f) return cudaErrorInvalidValue;
    size_t smem = scan_smem_bytes(epl, dp, p.stage_keys);
    return launch_pdl(f, dim3(grid), dim3(SCAN_THREADS), smem, s, p);
}

// rows one CTA consumes per tile (grid sizing in the API layer)
int scan_tile_rows(int dp) {
    const ScanShape s = shape_of(dp);
    if (!s.lpr) return 0;
    return SCAN_WARPS * (32 / s.lpr) * s.u;
}

}  // namespace b2r
#endif  // B2R_DP
