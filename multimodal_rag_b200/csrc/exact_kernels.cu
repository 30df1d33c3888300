// Instantiations + launch table for K5 (exact.cuh).
#include <algorithm>
#include <map>
#include <tuple>

#include "engine.h"

namespace b2r {
namespace {
typedef void (*exact_fn)(const ExactParams);
exact_fn lookup(int epl, int g) {
    if (g == 4) {
        switch (epl) {
            case 1: return exact_topk_kernel<1, 4>;
            case 2: return exact_topk_kernel<2, 4>;
            case 4: return exact_topk_kernel<4, 4>;
            default: return nullptr;
        }
    }
    if (g == 2) return epl == 8 ? exact_topk_kernel<8, 2> : nullptr;
    if (g == 1) {
        switch (epl) {
            case 1: return exact_topk_kernel<1, 1>;
            case 2: return exact_topk_kernel<2, 1>;
            case 4: return exact_topk_kernel<4, 1>;
            case 8: return exact_topk_kernel<8, 1>;
            default: return nullptr;
        }
    }
    return nullptr;
}
}  // namespace

// queries scored per corpus pass: 4 (2 for the 256-entry lists) while their fp64 copies fit 64 KB of shared memory
int exact_group(int epl, int dp) {
    const int g = epl == 8 ? 2 : EXACT_MAX_G;
    return (size_t)g * dp * 8 <= 65536 ? g : 1;
}

int exact_max_grid(int epl, int dp, int sm_count) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    auto key = std::make_tuple(dev, epl, dp);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second * sm_count;
    const int grp = exact_group(epl, dp);
    exact_fn f = lookup(epl, grp);
    if (!f) return 0;
    const size_t smem = exact_smem_bytes(epl, grp, dp);
    // the attribute is per function, not per dp: raise it to the largest this instantiation is launched with
    // (G > 1: the fp64 query copies stay under 64 KB; G = 1: the largest dp the ABI accepts)
    const size_t smem_max = grp > 1 ? exact_smem_bytes(epl, grp, 65536 / 8 / grp) : exact_smem_bytes(epl, 1, 8192);
    if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) != cudaSuccess) return 0;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f, EXACT_THREADS, smem) != cudaSuccess) return 0;
    cache[key] = per_sm;
    return per_sm * sm_count;
}

cudaError_t exact_launch(int epl, const ExactParams &p, int grid, cudaStream_t s, size_t rider_smem) {
    const int grp = exact_group(epl, p.fin.dp);
    exact_fn f = lookup(epl, grp);
    if (!f) return cudaErrorInvalidValue;
    return launch_pdl(f, dim3(grid), dim3(EXACT_THREADS), std::max(exact_smem_bytes(epl, grp, p.fin.dp), rider_smem), s, p);
}

// what exact_max_grid raised the function attribute to
size_t exact_smem_limit(int epl, int dp) {
    const int grp = exact_group(epl, dp);
    return grp > 1 ? exact_smem_bytes(epl, grp, 65536 / 8 / grp) : exact_smem_bytes(epl, 1, 8192);
}

}  // namespace b2r
