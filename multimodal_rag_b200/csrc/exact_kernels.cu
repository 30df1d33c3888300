// Instantiations + launch table for K5 (exact.cuh).
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
    exact_fn f = lookup(epl);
    if (!f) return 0;
    size_t smem = exact_smem_bytes(epl, dp);
    if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f, EXACT_THREADS, smem) != cudaSuccess) return 0;
    return per_sm * sm_count;
}

cudaError_t exact_launch(int epl, const ExactParams &p, int grid, cudaStream_t s) {
    exact_fn f = lookup(epl);
    if (!f) return cudaErrorInvalidValue;
    f<<<grid, EXACT_THREADS, exact_smem_bytes(epl, p.fin.dp), s>>>(p);
    return cudaGetLastError();
}

}  // namespace b2r
