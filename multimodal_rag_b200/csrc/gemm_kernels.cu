// Instantiations + launch table for K3 (gemm.cuh) and the TMA tensor-map encoding.
// Compiled once per supported K-block count (-DB2R_KB=...) plus a dispatcher unit.
#include <map>
#include <tuple>

#include "engine.h"
#include "gemm.cuh"

namespace b2r {

typedef void (*gemm_fn)(const CUtensorMap, const CUtensorMap, const GemmParams);

#ifdef B2R_KB
#define B2R_CAT2(a, b) a##b
#define B2R_CAT(a, b) B2R_CAT2(a, b)
template <bool PAIR, int BM>
static gemm_fn pick(int L, bool bias) {
    if (L == 8) return bias ? gemm_topk_kernel<B2R_KB, 8, true, PAIR, BM> : gemm_topk_kernel<B2R_KB, 8, false, PAIR, BM>;
    if (L == 16) return bias ? gemm_topk_kernel<B2R_KB, 16, true, PAIR, BM> : gemm_topk_kernel<B2R_KB, 16, false, PAIR, BM>;
    if (L == 32) return bias ? gemm_topk_kernel<B2R_KB, 32, true, PAIR, BM> : gemm_topk_kernel<B2R_KB, 32, false, PAIR, BM>;
    if (L == 0) return bias ? gemm_topk_kernel<B2R_KB, 0, true, PAIR, BM> : gemm_topk_kernel<B2R_KB, 0, false, PAIR, BM>;   // pool mode
    return nullptr;
}
gemm_fn B2R_CAT(gemm_lookup_, B2R_KB)(int L, bool bias, bool pair, int bm) {
#if B2R_KB > 8          // 64-query blocks are only used beyond 512 dims
    if (!pair && bm == 64) return pick<false, 64>(L, bias);
#endif
    if (bm != 128) return nullptr;
    return pair ? pick<true, 128>(L, bias) : pick<false, 128>(L, bias);
}
}  // namespace b2r
#else
gemm_fn gemm_lookup_2(int, bool, bool, int);
gemm_fn gemm_lookup_4(int, bool, bool, int);
gemm_fn gemm_lookup_6(int, bool, bool, int);
gemm_fn gemm_lookup_8(int, bool, bool, int);
gemm_fn gemm_lookup_12(int, bool, bool, int);
gemm_fn gemm_lookup_16(int, bool, bool, int);
gemm_fn gemm_lookup_24(int, bool, bool, int);

// ---------------------------------------------------------------------------------
// pass bitmap: bit r of word r>>5 = row r is live, passes the type mask and the allow bitmap.
// Words cover [0, n_words*32); rows >= n get 0, so tile padding never scores.
// ---------------------------------------------------------------------------------
static __global__ void pass_bits_kernel(const uint8_t *__restrict__ type_code, unsigned long long type_mask,
                                 const uint32_t *__restrict__ allow_bits, unsigned n, unsigned n_words,
                                 uint32_t *__restrict__ out) {
    const unsigned gt = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned nthreads = gridDim.x * blockDim.x;
    for (unsigned long long r = gt; r < (unsigned long long)n_words * 32; r += nthreads) {
        bool ok = r < n && row_passes((unsigned)r, type_code, type_mask, allow_bits);
        unsigned w = __ballot_sync(FULL_MASK, ok);
        if ((threadIdx.x & 31) == 0) out[r >> 5] = w;
    }
}


namespace {
gemm_fn lookup(int kb, int L, bool bias, bool pair, int bm = 128) {
    switch (kb) {
        case 2:  return gemm_lookup_2(L, bias, pair, bm);
        case 4:  return gemm_lookup_4(L, bias, pair, bm);
        case 6:  return gemm_lookup_6(L, bias, pair, bm);     // all-MiniLM-L6-v2 (384)
        case 8:  return gemm_lookup_8(L, bias, pair, bm);     // CLIP ViT-B/32 shape (512)
        case 12: return gemm_lookup_12(L, bias, pair, bm);    // 768
        case 16: return gemm_lookup_16(L, bias, pair, bm);    // 1024
        case 24: return gemm_lookup_24(L, bias, pair, bm);    // 1536
        default: return nullptr;
    }
}
size_t smem_of(int kb, bool pair, int bm = 128) {
    switch (kb) {
        case 2:  return gemm_smem_bytes(2, pair, bm);
        case 4:  return gemm_smem_bytes(4, pair, bm);
        case 6:  return gemm_smem_bytes(6, pair, bm);
        case 8:  return gemm_smem_bytes(8, pair, bm);
        case 12: return gemm_smem_bytes(12, pair, bm);
        case 16: return gemm_smem_bytes(16, pair, bm);
        case 24: return gemm_smem_bytes(24, pair, bm);
        default: return 0;
    }
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_fn get_encode() {
    static encode_fn fn = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    if (fn) return fn;
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !sym) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<encode_fn>(sym);
    return fn;
}
}  // namespace

// per-thread list length for n_results = k; 0 = pool mode (no lists, 32 < k <= 128); -1 = unsupported
int gemm_list_len(int k) { return k <= 8 ? 8 : k <= 16 ? 16 : k <= 32 ? 32 : k <= 128 ? 0 : -1; }
int gemm_tile_rows(int dp) { return gemm_bn(dp / 64); }
bool gemm_supported(int dp, int k) { return dp % 64 == 0 && lookup(dp / 64, 8, false, false) != nullptr && gemm_list_len(k) >= 0; }

// [rows, dp] bf16 row-major -> 2-D tensor map, box = 64 elements (128 B, one swizzle row) x box_rows
int gemm_encode_map(CUtensorMap *out, const void *base, int dp, uint64_t rows, int box_rows) {
    encode_fn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return B2R_ECUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)dp, rows ? rows : 1};
    cuuint64_t strides[1] = {(cuuint64_t)dp * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r)); return B2R_ECUDA; }
    return B2R_OK;
}

static cudaError_t gemm_prepare(gemm_fn f, size_t smem) {
    static std::mutex mu;
    static std::map<std::tuple<int, const void *>, bool> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    auto key = std::make_tuple(dev, (const void *)f);
    if (!done.count(key)) {
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        done[key] = true;
    }
    return cudaSuccess;
}

// CTA pairs (clusters of 2) of the pair kernel that can be co-resident on this device: 0 = the pair form is not usable
int gemm_max_pairs(int dp, int L, bool bias) {
    const int kb = dp / 64;
    gemm_fn f = lookup(kb, L, bias, true);
    if (!f) return 0;
    static std::mutex mu;
    static std::map<std::tuple<int, const void *>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(std::make_tuple(dev, (const void *)f));
        if (it != cache.end()) return it->second;
    }
    const size_t smem = smem_of(kb, true);
    int n = 0;
    if (gemm_prepare(f, smem) == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, f, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    }
    std::lock_guard<std::mutex> g(mu);
    cache[std::make_tuple(dev, (const void *)f)] = n;
    return n;
}

// pair = launch clusters of two CTAs (query blocks 2j, 2j+1 of a slice); tm_x must then describe half-tile boxes.
// bm = queries per CTA (128, or 64 for one block of at most 64 queries, never with pair); tm_q's box must have bm rows
cudaError_t gemm_launch(int dp, int L, bool bias, bool pair, int bm, const CUtensorMap &tm_q, const CUtensorMap &tm_x,
                        const GemmParams &p, cudaStream_t s) {
    const int kb = dp / 64;
    gemm_fn f = lookup(kb, L, bias, pair, bm);
    if (!f) return cudaErrorInvalidValue;
    const size_t smem = smem_of(kb, pair, bm);
    cudaError_t e = gemm_prepare(f, smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.n_slices * p.n_qblocks); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (pair) {
        at[1].id = cudaLaunchAttributeClusterDimension;
        at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    }
    return cudaLaunchKernelEx(&cfg, f, tm_q, tm_x, p);
}

cudaError_t pass_bits_launch(const uint8_t *type_code, unsigned long long type_mask, const uint32_t *allow_bits,
                             unsigned n, unsigned n_words, uint32_t *out, int sm_count, cudaStream_t s) {
    unsigned blocks = (unsigned)std::min<unsigned long long>(((unsigned long long)n_words * 32 + 255) / 256,
                                                             (unsigned long long)sm_count * 8);
    if (blocks == 0) blocks = 1;
    pass_bits_kernel<<<blocks, 256, 0, s>>>(type_code, type_mask, allow_bits, n, n_words, out);
    return cudaGetLastError();
}

cudaError_t finalize_union_launch(int epl, const FinalizeParams &fin, const UnionParams &u, int q0, int nq, cudaStream_t s) {
    const size_t smem = finalize_union_smem(epl, fin.dp);
    if (smem > 48 * 1024) {      // opt in to large dynamic shared memory (per function; cheap, idempotent)
        cudaError_t e = cudaSuccess;
        if (epl == 1) e = cudaFuncSetAttribute(finalize_union_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (epl == 2) e = cudaFuncSetAttribute(finalize_union_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (epl == 4) e = cudaFuncSetAttribute(finalize_union_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (epl == 8) e = cudaFuncSetAttribute(finalize_union_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    switch (epl) {
        case 1: return launch_pdl(finalize_union_kernel<1>, dim3(nq), dim3(FIN_THREADS), smem, s, fin, u, q0);
        case 2: return launch_pdl(finalize_union_kernel<2>, dim3(nq), dim3(FIN_THREADS), smem, s, fin, u, q0);
        case 4: return launch_pdl(finalize_union_kernel<4>, dim3(nq), dim3(FIN_THREADS), smem, s, fin, u, q0);
        case 8: return launch_pdl(finalize_union_kernel<8>, dim3(nq), dim3(FIN_THREADS), smem, s, fin, u, q0);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace b2r
#endif  // B2R_KB
