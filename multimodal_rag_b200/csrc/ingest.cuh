// K1: fused ingest.  One warp per row: (cosine) fp32 L2-normalise with hnswlib's
// formula, round-to-nearest bf16 pack with 128-bit coalesced stores, optional fp32
// master store, per-row -|x|^2/2 (l2 space) and a running max |x|^2 for the
// certificate's error bound.  Also prepares query batches (same normalisation, no pack).
#pragma once
#include "common.cuh"

namespace b2r {

constexpr int INGEST_THREADS = 256;

struct IngestParams {
    const float *x;            // [n, d] fp32 (device)
    long long n;
    int d, dp, space;
    uint4 *corpus;             // destination of row 0 of this batch: bf16 [n, dp] (nullptr: queries)
    float *master;             // fp32 [n, dp] or nullptr
    float *bias;               // [n] or nullptr (only l2)
    uint8_t *type_out;         // [n] or nullptr
    const uint8_t *type_in;    // [n] device or nullptr (= 0)
    float *max_norm2;          // device [2] or nullptr: [0] max |x|^2 of the stored rows, [1] max |x - bf16(x)|^2
    float *qerr;               // [n] or nullptr (query batches for K3): |q - bf16(q)|, rounded up
    double *q_eps;             // query preparation: [n][2] = { error bound eps of the scan scores, |q|^2 } (finalize.cuh) ...
    const float *norms;        // ... from the stored rows' max |x|^2, max |x - bf16(x)|^2 ...
    float eps_rel;             // ... and the scoring kernel's accumulation slop; qerr == nullptr: queries are not rounded
    unsigned *zero[3];         // query preparation: per-call shared state (K3's bounds, cursors, flags, counters; the
    int zero_words[3];         // fix-up work list control; K5's slot generations) is cleared here, after the previous
                               // query's kernels have finished with it
};

// 1/(sqrt(s)+1e-30) exactly as hnswlib's normalize_vector does it in fp32
__device__ __forceinline__ float hnsw_inv_norm(double sumsq) {
    float s = __double2float_rn(sumsq);
    return __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(s), 1e-30f));
}

__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);     // .x = lo (low 16 bits), .y = hi
    return *reinterpret_cast<unsigned *>(&b);
}

template <bool VEC>   // VEC: d % 8 == 0 and x 16-byte aligned -> float4 loads
__global__ void __launch_bounds__(INGEST_THREADS) ingest_kernel(const IngestParams p) {
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (INGEST_THREADS / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (INGEST_THREADS / 32);
    const int d = p.d, dp = p.dp, chunks = dp / 8;
    pdl_wait();          // query preparation overwrites buffers the previous query's kernels may still read
    pdl_trigger();
#pragma unroll
    for (int z = 0; z < 3; ++z)
        for (int i = blockIdx.x * INGEST_THREADS + threadIdx.x; i < p.zero_words[z]; i += gridDim.x * INGEST_THREADS) p.zero[z][i] = 0u;
    for (long long r = gw; r < p.n; r += nw) {
        const float *xr = p.x + r * d;
        float inv = 1.0f;
        if (p.space == 1) {
            double s = 0.0;
            if (VEC) {
                const float4 *x4 = reinterpret_cast<const float4 *>(xr);
                for (int c = lane; c < d / 4; c += 32) {
                    float4 a = __ldg(x4 + c);
                    s = fma((double)a.x, (double)a.x, s); s = fma((double)a.y, (double)a.y, s);
                    s = fma((double)a.z, (double)a.z, s); s = fma((double)a.w, (double)a.w, s);
                }
            } else {
                for (int i = lane; i < d; i += 32) { double a = (double)__ldg(xr + i); s = fma(a, a, s); }
            }
            inv = hnsw_inv_norm(warp_sum(s));
        }
        double s2 = 0.0, e2 = 0.0;
        for (int c = lane; c < chunks; c += 32) {
            float y[8];
            if (VEC && c * 8 < d) {
                const float4 *x4 = reinterpret_cast<const float4 *>(xr + c * 8);
                float4 a = __ldg(x4), b = __ldg(x4 + 1);
                y[0] = a.x; y[1] = a.y; y[2] = a.z; y[3] = a.w; y[4] = b.x; y[5] = b.y; y[6] = b.z; y[7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = (c * 8 + i < d) ? __ldg(xr + c * 8 + i) : 0.f;
            }
            if (p.space == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = __fmul_rn(y[i], inv);
            }
            uint4 w;
            w.x = pack_bf16x2(y[0], y[1]); w.y = pack_bf16x2(y[2], y[3]);
            w.z = pack_bf16x2(y[4], y[5]); w.w = pack_bf16x2(y[6], y[7]);
            if (p.corpus) p.corpus[(size_t)r * chunks + c] = w;
            if (p.master) {
                float4 *m4 = reinterpret_cast<float4 *>(p.master + (size_t)r * dp + c * 8);
                m4[0] = make_float4(y[0], y[1], y[2], y[3]);
                m4[1] = make_float4(y[4], y[5], y[6], y[7]);
#pragma unroll
                for (int i = 0; i < 8; ++i) s2 = fma((double)y[i], (double)y[i], s2);
                const unsigned ww[4] = {w.x, w.y, w.z, w.w};   // rounding error of the packed copy (exact in fp32)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double a = (double)(y[2 * i] - bf16lo(ww[i])), b = (double)(y[2 * i + 1] - bf16hi(ww[i]));
                    e2 = fma(a, a, e2); e2 = fma(b, b, e2);
                }
            } else {
                const unsigned ww[4] = {w.x, w.y, w.z, w.w};   // the stored corpus is the bf16 rounding
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double a = (double)bf16lo(ww[i]), b = (double)bf16hi(ww[i]);
                    s2 = fma(a, a, s2); s2 = fma(b, b, s2);
                }
            }
        }
        s2 = warp_sum(s2);
        e2 = warp_sum(e2);
        if (lane == 0) {
            if (p.q_eps) {
                // |scan - exact| <= |q~-q| (max|x| + max|x~-x|) + |q| max|x~-x| + accumulation slop (see finalize.cuh)
                const double dq = p.qerr ? (double)__double2float_ru(sqrt(e2) * (1.0 + 1e-6)) : 0.0;
                const double xn = sqrt((double)p.norms[0]), dxn = sqrt((double)p.norms[1]), qn = sqrt(s2);
                p.q_eps[2 * r] = dq * (xn + dxn) + qn * dxn + (double)p.eps_rel * (qn + dq) * (xn + dxn) +
                                 1e-6 * (0.5 * xn * xn + qn * xn) + 1e-30;
                p.q_eps[2 * r + 1] = s2;
            }
            if (p.qerr) p.qerr[r] = __double2float_ru(sqrt(e2) * (1.0 + 1e-6));
            if (p.bias) p.bias[r] = __double2float_rn(-0.5 * s2);
            if (p.type_out) p.type_out[r] = p.type_in ? p.type_in[r] : (uint8_t)0;
            if (p.max_norm2) {
                atomicMax(reinterpret_cast<unsigned *>(p.max_norm2), __float_as_uint(__double2float_ru(s2)));
                atomicMax(reinterpret_cast<unsigned *>(p.max_norm2) + 1, __float_as_uint(__double2float_ru(e2)));
            }
        }
    }
}

// tombstone: type_code[row] = DEAD; counts rows that were alive
__global__ void tombstone_kernel(uint8_t *type_code, const long long *rows, long long n, long long limit,
                                 unsigned long long *n_killed) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long r = rows[i];
    if (r < 0 || r >= limit) return;
    // duplicates inside one call: only the first writer counts (byte CAS via 32-bit word)
    unsigned *word = reinterpret_cast<unsigned *>(type_code + (r & ~3ll));
    const unsigned shift = (unsigned)(r & 3) * 8;
    unsigned old = *word, assumed;
    do {
        assumed = old;
        if (((assumed >> shift) & 0xffu) == B2R_TYPE_DEAD) return;
        unsigned nv = (assumed & ~(0xffu << shift)) | ((unsigned)B2R_TYPE_DEAD << shift);
        old = atomicCAS(word, assumed, nv);
    } while (old != assumed);
    atomicAdd(n_killed, 1ull);
}

// gather stored rows (fp32 master, or widened bf16) into a dense [n, d] fp32 buffer
__global__ void gather_rows_kernel(const float *master, const uint4 *corpus, int d, int dp,
                                   const long long *rows, long long n, long long limit, float *out) {
    const int lane = threadIdx.x & 31;
    long long w = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (w >= n) return;
    long long r = rows[w];
    for (int i = lane; i < d; i += 32) {
        float v = 0.f;
        if (r >= 0 && r < limit) {
            if (master) v = master[(size_t)r * dp + i];
            else {
                const unsigned short *b = reinterpret_cast<const unsigned short *>(corpus + (size_t)r * (dp / 8));
                v = __uint_as_float((unsigned)b[i] << 16);
            }
        }
        out[(size_t)w * d + i] = v;
    }
}

}  // namespace b2r
