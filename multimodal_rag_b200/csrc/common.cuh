// Shared device helpers: orderable keys, warp-resident sorted top-K lists, bf16 unpack.
// sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define B2R_WARP 32
#define FULL_MASK 0xffffffffu

namespace b2r {

// ---------------------------------------------------------------------------------
// Keys.  A key type K provides: better(a,b) (a ranks strictly before b), worst(),
// shfl / shfl_up, and is trivially copyable.
// ---------------------------------------------------------------------------------

// fp32 score (larger = better) + local row packed in one u64 so that a plain integer
// compare orders by (score desc, row asc).
struct KeyS {
    unsigned long long v;
    __device__ __forceinline__ static unsigned ord(float f) {
        unsigned b = __float_as_uint(f);
        return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
    }
    __device__ __forceinline__ static float unord(unsigned o) {
        unsigned b = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
        return __uint_as_float(b);
    }
    __device__ __forceinline__ static KeyS make(float score, unsigned row) {
        KeyS k; k.v = ((unsigned long long)ord(score) << 32) | (unsigned)(~row); return k;
    }
    __device__ __forceinline__ static KeyS worst() { KeyS k; k.v = 0ull; return k; }
    __device__ __forceinline__ bool valid() const { return v != 0ull; }
    __device__ __forceinline__ float score() const { return unord((unsigned)(v >> 32)); }
    __device__ __forceinline__ unsigned row() const { return ~(unsigned)v; }
    __device__ __forceinline__ static bool better(const KeyS &a, const KeyS &b) { return a.v > b.v; }
    __device__ __forceinline__ static KeyS shfl(const KeyS &a, int src) {
        KeyS k; k.v = __shfl_sync(FULL_MASK, a.v, src); return k;
    }
    __device__ __forceinline__ static KeyS shfl_up(const KeyS &a, int d) {
        KeyS k; k.v = __shfl_up_sync(FULL_MASK, a.v, d); return k;
    }
    __device__ __forceinline__ static KeyS shfl_xor(const KeyS &a, int m) {
        KeyS k; k.v = __shfl_xor_sync(FULL_MASK, a.v, m); return k;
    }
    // read a key another CTA wrote (L2, never L1)
    __device__ __forceinline__ static KeyS load_cg(const KeyS *p) { KeyS k; k.v = __ldcg(&p->v); return k; }
};

// exact fp64 distance (smaller = better) + row; order (distance asc, row asc).
struct KeyD {
    double d;
    unsigned row;
    __device__ __forceinline__ static KeyD make(double dist, unsigned r) { KeyD k; k.d = dist; k.row = r; return k; }
    __device__ __forceinline__ static KeyD worst() { KeyD k; k.d = __longlong_as_double(0x7ff0000000000000ll); k.row = 0xffffffffu; return k; }
    __device__ __forceinline__ bool valid() const { return row != 0xffffffffu; }
    __device__ __forceinline__ static bool better(const KeyD &a, const KeyD &b) {
        return a.d < b.d || (a.d == b.d && a.row < b.row);
    }
    __device__ __forceinline__ static KeyD shfl(const KeyD &a, int src) {
        KeyD k; k.d = __shfl_sync(FULL_MASK, a.d, src); k.row = __shfl_sync(FULL_MASK, a.row, src); return k;
    }
    __device__ __forceinline__ static KeyD shfl_up(const KeyD &a, int dl) {
        KeyD k; k.d = __shfl_up_sync(FULL_MASK, a.d, dl); k.row = __shfl_up_sync(FULL_MASK, a.row, dl); return k;
    }
    __device__ __forceinline__ static KeyD shfl_xor(const KeyD &a, int m) {
        KeyD k; k.d = __shfl_xor_sync(FULL_MASK, a.d, m); k.row = __shfl_xor_sync(FULL_MASK, a.row, m); return k;
    }
    __device__ __forceinline__ static KeyD load_cg(const KeyD *p) {
        KeyD k; k.d = __ldcg(&p->d); k.row = __ldcg(&p->row); return k;
    }
};

// ---------------------------------------------------------------------------------
// WarpList: the best KP = 32*EPL keys seen so far, sorted, resident in the registers
// of one warp.  Blocked layout: rank r lives in lane r/EPL, slot r%EPL, so a
// shift-by-one needs a single shuffle.  All 32 lanes must call every method with the
// same arguments (the key to insert is warp-uniform).
// ---------------------------------------------------------------------------------
template <class K, int EPL>
struct WarpList {
    K key[EPL];
    K thr;   // copy of the worst kept key (rank KP-1); warp-uniform

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int e = 0; e < EPL; ++e) key[e] = K::worst();
        thr = K::worst();
    }
    __device__ __forceinline__ bool accepts(const K &k) const { return K::better(k, thr); }

    // insert a warp-uniform key known to satisfy accepts()
    __device__ __forceinline__ void insert(const K &nk, int lane) {
        int c = 0;
#pragma unroll
        for (int e = 0; e < EPL; ++e) c += K::better(key[e], nk) ? 1 : 0;   // my keys that stay ahead
        K from_prev = K::shfl_up(key[EPL - 1], 1);
        int c_prev = __shfl_up_sync(FULL_MASK, c, 1);
        bool prev_full = (lane == 0) || (c_prev == EPL);
        if (c < EPL) {
            K ins = prev_full ? nk : from_prev;
#pragma unroll
            for (int e = EPL - 1; e >= 1; --e)
                if (e > c) key[e] = key[e - 1];
#pragma unroll
            for (int e = 0; e < EPL; ++e)
                if (e == c) key[e] = ins;
        }
        thr = K::shfl(key[EPL - 1], 31);
    }
    __device__ __forceinline__ void offer(const K &k, int lane) {
        if (accepts(k)) insert(k, lane);
    }
    // rank-ordered store: dst[r] for r in [0, KP)
    __device__ __forceinline__ void store(K *dst, int lane) const {
#pragma unroll
        for (int e = 0; e < EPL; ++e) dst[lane * EPL + e] = key[e];
    }
    // merge a rank-ordered (best first) list of `len` keys; stops at the first reject
    __device__ __forceinline__ void merge_sorted(const K *src, int len, int lane) {
        for (int i = 0; i < len; ++i) {
            K k = src[i];
            if (!k.valid() || !accepts(k)) break;
            insert(k, lane);
        }
    }
    // Bitonic top-KP merge with a full rank-ordered list of KP keys (worst() padded).
    // c[r] = best(a[r], b[KP-1-r]) is bitonic and holds the KP best of the union; log2(KP)
    // compare-exchange stages re-sort it.  Cross-lane stages use shuffles, the rest registers.
    template <bool CG>
    __device__ __forceinline__ void merge_bitonic(const K *src, int lane) {
        constexpr int KP = 32 * EPL;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const K *p = src + (KP - 1 - (lane * EPL + e));
            K b = CG ? K::load_cg(p) : *p;
            if (K::better(b, key[e])) key[e] = b;
        }
        resort_bitonic(lane);
    }
    __device__ __forceinline__ void resort_bitonic(int lane) {
        constexpr int KP = 32 * EPL;
#pragma unroll
        for (int s = KP / 2; s >= 1; s >>= 1) {
            if (s >= EPL) {
                const int ls = s / EPL;
                const bool keep_best = (lane & ls) == 0;
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    K o = K::shfl_xor(key[e], ls);
                    bool ob = K::better(o, key[e]);
                    if (ob == keep_best) key[e] = o;
                }
            } else {
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    if ((e & s) == 0) {
                        K a = key[e], b = key[e + s];
                        if (K::better(b, a)) { key[e] = b; key[e + s] = a; }
                    }
                }
            }
        }
        thr = K::shfl(key[EPL - 1], 31);
    }
};

// ---------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float bf16lo(unsigned w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(unsigned w) { return __uint_as_float(w & 0xffff0000u); }

// streaming 16-byte load: read-only path, do not allocate in L1
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// Programmatic dependent launch (PDL).  Every kernel of a query is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so it may start while its predecessor on the stream is
// still running: pdl_wait() blocks until the predecessor grid has completed and its writes are visible (call
// it before the first access to anything an earlier kernel produced OR still reads), pdl_trigger() lets the
// successor begin its own launch/prologue early.  Both are no-ops in a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// row filter: tombstones carry type code 63, which no mask ever has set
#define B2R_TYPE_DEAD 63
__device__ __forceinline__ bool row_passes(unsigned row, const uint8_t *__restrict__ type_code,
                                           unsigned long long type_mask,
                                           const uint32_t *__restrict__ allow_bits) {
    bool ok = (type_mask >> type_code[row]) & 1ull;
    if (allow_bits) ok = ok && ((allow_bits[row >> 5] >> (row & 31)) & 1u);
    return ok;
}

}  // namespace b2r
