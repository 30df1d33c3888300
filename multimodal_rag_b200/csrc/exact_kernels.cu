// Instantiations + launch table for K5 (exact.cuh).
#include <map>
#include <tuple>

#include "engine.h"

namespace b2r {
namespace {
typedef void (*exact_fn)(const ExactParams);
exact_fn lookup(int epl) {
    switch (epl) {
        case 1: return exact_topk_kernel<1>;
        case 2: return exact_topk_kernel<2>;
        case 4: return exact_topk_kernel<4>;
        case 8: return exact_topk_kernel<8>;
        default: return nullptr;
    }
}
}  // namespace

int exact_max_grid(int epl, int dp, int sm_count) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    auto key = std::make_tuple(dev, epl, dp);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second * sm_count;
    exact_fn f = lookup(epl);
    if (!f) return 0;
    size_t smem = exact_smem_bytes(epl, dp);
    // the attribute is per function (per epl), not per dp: raise it to the largest dp the ABI accepts
    if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)exact_smem_bytes(epl, 8192)) != cudaSuccess) return 0;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f, EXACT_THREADS, smem) != cudaSuccess) return 0;
    cache[key] = per_sm;
    return per_sm * sm_count;
}

cudaError_t exact_launch(int epl, const ExactParams &p, int grid, cudaStream_t s) {
    exact_fn f = lookup(epl);
    if (!f) return cudaErrorInvalidValue;
    return launch_pdl(f, dim3(grid), dim3(EXACT_THREADS), exact_smem_bytes(epl, p.fin.dp), s, p);
}

}  // namespace b2r
