// K1: fused ingest.  One warp per row: (cosine) fp32 L2-normalise with hnswlib's
// formula, round-to-nearest bf16 pack with 128-bit coalesced stores, optional fp32
// master store, per-row -|x|^2/2 (l2 space) and a running max |x|^2 for the
// certificate's error bound.  Also prepares query batches (same normalisation, no pack).
#pragma once
#include "common.cuh"
#include "xchg.cuh"

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
    const unsigned *wait_words;    // b2r_query_push: the first kernel of the call holds the stream until every rank has read what
    int wait_n;                    // the mailbox slot held before (wait_words[0 .. wait_n) >= wait_val, written by the peers);
    unsigned wait_val;             // wait_n = 0 otherwise.  Normally true long before: the slot was used four calls ago.
    XchgFlags flags;               // ... and writes the flag words the previous fused call left behind (world = 0: none)
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

// Sum of squares in double-float (hi + lo, ~48 significant bits) with fp32 instructions only.  ncu showed the first
// version of this kernel bound by the XU pipe at 76 % (1.3 TB/s): every element was widened to fp64 (F2F runs at
// 16 lanes/clk/SM) three times.  x*x splits exactly into p + e with one FMA, TwoSum keeps the accumulation error;
// only the per-lane totals are widened.  The _rn intrinsics keep the compiler from contracting the steps into FMAs.
struct DFloat {
    float hi = 0.f, lo = 0.f;
    __device__ __forceinline__ void add_sq(float x) {
        const float p = __fmul_rn(x, x);
        const float e = __fmaf_rn(x, x, -p);                 // p + e == x * x exactly
        const float s = __fadd_rn(hi, p);
        const float bb = __fsub_rn(s, hi);
        const float err = __fadd_rn(__fsub_rn(hi, __fsub_rn(s, bb)), __fsub_rn(p, bb));   // hi + p == s + err exactly
        hi = s;
        lo = __fadd_rn(lo, __fadd_rn(err, e));
    }
    __device__ __forceinline__ double value() const { return (double)hi + (double)lo; }
};

// per-row epilogue shared by both builds: error bound of a prepared query, bias, type code, running maxima
__device__ __forceinline__ void ingest_row_tail(const IngestParams &p, long long r, double s2, double e2, float &max_s2,
                                                float &max_e2) {
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
    // running maxima of this warp's rows; published once per warp (one atomic per ROW on the same two words capped
    // the first version at 0.36 G rows/s whatever the row length -- same-address atomics serialise in the L2)
    max_s2 = fmaxf(max_s2, __double2float_ru(s2 * (1.0 + 1e-12)));
    max_e2 = fmaxf(max_e2, __double2float_ru(e2));
}

// scale 8 values, pack, store the bf16 chunk and the fp32 master chunk, accumulate |y|^2 and the rounding error
__device__ __forceinline__ void ingest_chunk(const IngestParams &p, long long r, int c, int chunks, int dp, float (&y)[8],
                                             float inv, DFloat &s2d, float &e2f) {
    if (p.space == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = __fmul_rn(y[i], inv);
    }
    uint4 w;
    w.x = pack_bf16x2(y[0], y[1]); w.y = pack_bf16x2(y[2], y[3]);
    w.z = pack_bf16x2(y[4], y[5]); w.w = pack_bf16x2(y[6], y[7]);
    if (p.corpus) p.corpus[(size_t)r * chunks + c] = w;
    const unsigned ww[4] = {w.x, w.y, w.z, w.w};
    if (p.master) {
        float4 *m4 = reinterpret_cast<float4 *>(p.master + (size_t)r * dp + c * 8);
        m4[0] = make_float4(y[0], y[1], y[2], y[3]);
        m4[1] = make_float4(y[4], y[5], y[6], y[7]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s2d.add_sq(y[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {          // rounding error of the packed copy (exact in fp32)
            const float a = y[2 * i] - bf16lo(ww[i]), b = y[2 * i + 1] - bf16hi(ww[i]);
            e2f = fmaf(a, a, e2f); e2f = fmaf(b, b, e2f);
        }
    } else {                                    // the stored corpus is the bf16 rounding
#pragma unroll
        for (int i = 0; i < 4; ++i) { s2d.add_sq(bf16lo(ww[i])); s2d.add_sq(bf16hi(ww[i])); }
    }
}

// NJ > 0: d % 8 == 0, x 16-byte aligned and dp <= 256 * NJ -- the whole row sits in registers (8 * NJ floats per
// lane, lane owns the 8-element chunks lane, lane + 32, ...): ONE trip to DRAM per row, all of a lane's loads in
// flight at once.  (The first version read the row in a dependent loop for the norm and again for the pack: ncu
// showed it waiting on memory -- long scoreboard -- at 1.3 TB/s.)
// NJ == 0: any d / alignment, looped scalar loads.
template <int NJ>
__global__ void __launch_bounds__(INGEST_THREADS) ingest_kernel(const IngestParams p) {
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (INGEST_THREADS / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (INGEST_THREADS / 32);
    const int d = p.d, dp = p.dp, chunks = dp / 8;
    pdl_wait();          // query preparation overwrites buffers the previous query's kernels may still read
    pdl_trigger();
    if (p.flags.world && blockIdx.x == gridDim.x - 1) xchg_publish_flags(p.flags, 32);
    if (p.wait_n && blockIdx.x == 0 && (int)threadIdx.x < p.wait_n) {
        unsigned v;
        const unsigned long long t0 = globaltimer_ns();
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.wait_words + threadIdx.x) : "memory");
            if ((int)(v - p.wait_val) >= 0) break;
            if (globaltimer_ns() - t0 > 4000000000ull) __trap();      // a peer that stopped merging: fail, do not hang
            __nanosleep(200);
        }
    }
#pragma unroll
    for (int z = 0; z < 3; ++z)
        for (int i = blockIdx.x * INGEST_THREADS + threadIdx.x; i < p.zero_words[z]; i += gridDim.x * INGEST_THREADS) p.zero[z][i] = 0u;
    float max_s2 = 0.f, max_e2 = 0.f;
    for (long long r = gw; r < p.n; r += nw) {
        const float *xr = p.x + r * d;
        float inv = 1.0f;
        DFloat s2d;           // |y|^2 of the stored row (bias, max norm, eps): double-float, as accurate as fp64 was
        float e2f = 0.f;      // |y - bf16(y)|^2: only ever used as an upper bound, plain fp32 (inflated below)
        if constexpr (NJ > 0) {
            float y[NJ][8];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int c = lane + 32 * j;
                if (c * 8 < d) {                              // d % 8 == 0: a chunk is whole or padding
                    const float4 *x4 = reinterpret_cast<const float4 *>(xr + c * 8);
                    const float4 a = ldg_stream_f4(x4), b = ldg_stream_f4(x4 + 1);
                    y[j][0] = a.x; y[j][1] = a.y; y[j][2] = a.z; y[j][3] = a.w;
                    y[j][4] = b.x; y[j][5] = b.y; y[j][6] = b.z; y[j][7] = b.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[j][i] = 0.f;
                }
            }
            if (p.space == 1) {
                DFloat s;
#pragma unroll
                for (int j = 0; j < NJ; ++j)
#pragma unroll
                    for (int i = 0; i < 8; ++i) s.add_sq(y[j][i]);
                inv = hnsw_inv_norm(warp_sum(s.value()));
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int c = lane + 32 * j;
                if (c < chunks) ingest_chunk(p, r, c, chunks, dp, y[j], inv, s2d, e2f);
            }
        } else {
            if (p.space == 1) {
                DFloat s;
                for (int i = lane; i < d; i += 32) s.add_sq(__ldg(xr + i));
                inv = hnsw_inv_norm(warp_sum(s.value()));
            }
            for (int c = lane; c < chunks; c += 32) {
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = (c * 8 + i < d) ? __ldg(xr + c * 8 + i) : 0.f;
                ingest_chunk(p, r, c, chunks, dp, y, inv, s2d, e2f);
            }
        }
        const double s2 = warp_sum(s2d.value());                     // accurate to ~1e-14 relative
        const double e2 = warp_sum((double)e2f) * (1.0 + 1e-4);      // fp32 accumulation: round the bound up
        if (lane == 0) ingest_row_tail(p, r, s2, e2, max_s2, max_e2);
    }
    if (lane == 0 && p.max_norm2 && gw < p.n) {       // non-negative floats order like their bit patterns
        atomicMax(reinterpret_cast<unsigned *>(p.max_norm2), __float_as_uint(max_s2));
        atomicMax(reinterpret_cast<unsigned *>(p.max_norm2) + 1, __float_as_uint(max_e2));
    }
}

// launch table: rows of up to 256 * NJ padded dims in registers
typedef void (*ingest_fn)(const IngestParams);
inline ingest_fn ingest_lookup(int d, int dp, const void *x) {
    const bool vec = (d % 8 == 0) && (((uintptr_t)x & 15) == 0);
    if (!vec || dp > 1536) return ingest_kernel<0>;
    if (dp <= 256) return ingest_kernel<1>;
    if (dp <= 512) return ingest_kernel<2>;
    if (dp <= 768) return ingest_kernel<3>;
    if (dp <= 1024) return ingest_kernel<4>;
    return ingest_kernel<6>;
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
