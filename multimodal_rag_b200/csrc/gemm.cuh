// K3: batched scoring on the 5th-gen tensor cores (tcgen05 / TMEM / TMA), fused selection.
//
// One CTA = one block of 128 queries x one contiguous slice of the corpus.
//   A operand  128 prepared queries (bf16), all K-blocks, loaded ONCE by TMA and kept in shared memory
//   B operand  corpus tiles of BN rows, streamed K-block by K-block (64 bf16 = one 128-byte swizzle row)
//              through a STAGES-deep TMA/mbarrier ring
//   D          128 x BN fp32 scores in TMEM, double buffered (2*BN columns) so the tensor pipe fills
//              tile t+1 while the epilogue drains tile t
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..9 =
// epilogue, two per SM sub-partition so one hides the other's latencies.  In TMEM a lane is a query and a
// column is a corpus row, so every epilogue thread owns ONE query (and one half of each tile's columns):
// it streams that query's scores through a max-tree + one compare against a private threshold and keeps
// the best L rows it has seen in a register-resident sorted list.  The threshold is
// max(own L-th best, shared per-query bound): whenever a thread's list is full it publishes its L-th
// best score to gthr[q] (atomicMax); all slices of the same query read it once per tile, so the
// admission rate falls with the rows seen by the WHOLE grid, not by one slice.
//
// The scan starts from a seeded bound: every CTA first scans a few tiles of its slice in sampling mode, the
// CTAs of a query block exchange their best samples through global memory and fold them into gthr[q] while
// the tensor pipe already works on the real tiles (GemmParams::seed_tiles).  Pool mode (32 < k <= 128)
// keeps no lists: rows above the seeded bound go to private regions.
//
// What leaves the kernel: per query a compact pool of the admitted rows that still reach the published bound
// (KeyS: score, local row) and the final gthr[q].  Invariant used by the certificate (finalize_union_kernel):
// a row that is not in the pool was rejected by, evicted below, or filtered against a value that was
// published to gthr[q] -- the seed_rank-th best of the sampled rows' scores (k-th or L-th), or some list's L-th best -- so its bf16 score is
// <= the final gthr[q].
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "finalize.cuh"

namespace b2r {

constexpr int GEMM_BM = 128;
constexpr int GEMM_THREADS = 320;       // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_POOL_SAMPLE_RANK = 32;  // pool mode: the bound is the 32nd best sampled score
// Bounds on the waits of the in-kernel seeding phase.  List mode: 50 us, then the scan starts unseeded (its lists bound
// themselves) and picks the seed up from gthr[q] when it arrives.  Pool mode needs the bound before its first append, and
// the CTAs of a query block drift apart during a long sampling phase (1/32 of the shard: hundreds of microseconds at 25M+
// rows), so its budget scales with the sampling phase it has just measured: 2x its duration, within [100 us, 5 ms].
// Nothing depends on a wait succeeding: a fold that could not run yet is retried at every tile, and a pool-mode thread
// without a seed bounds itself from its own region (pool_region_tighten).
constexpr unsigned long long GEMM_SEED_TIMEOUT_NS = 50000ull;
constexpr unsigned long long GEMM_POOL_WAIT_MIN_NS = 100000ull, GEMM_POOL_WAIT_MAX_NS = 5000000ull;
constexpr int GEMM_HALVES = 2;          // epilogue warps w and w+4 share a TMEM lane quadrant and split a tile's columns
constexpr int GEMM_SMEM_LIMIT = 232448;       // 227 KB opt-in maximum per CTA
constexpr int GEMM_SMEM_SLACK = 1024 + 512;   // manual 1024-byte alignment + static barriers

// What the finalize needs to know about the candidate pools (K3's outputs).
struct UnionParams {
    const KeyS *lists;           // [nq][list_stride]
    int list_stride, max_entries;
    unsigned *gthr, *cnt;        // [nq] final shared bound / pool cursor (bit 31 = overflow)
    unsigned long long *pool_stats;
};

struct GemmParams {
    unsigned n;                  // rows in this shard
    int nq;                      // queries in the batch
    int qblock0;                 // first 128-query block of this launch (batches > 8 blocks are chunked)
    int n_qblocks, n_slices;     // grid = n_slices * n_qblocks (blockIdx = slice * n_qblocks + qblock)
    int list_stride;             // KeyS entries reserved per query in `lists` (>= n_slices * 2 * L)
    int tiles_total;             // tiles of the shard: ceil(n / BN)
    const uint32_t *pass_bits;   // bit r = row r is live and passes the filter; 0 for r >= n
    const float *bias;           // [n] -|x|^2/2 (l2) or nullptr
    unsigned *gthr;              // [n_qblocks*128] shared per-query bound, KeyS::ord encoding, 0 = none yet
    unsigned *cnt;               // [nq] entries appended to lists[q] so far (0 between calls)
    KeyS *lists;                 // [nq][list_stride]: every thread appends its valid entries (atomic cursor cnt[q])
    KeyS *regions;               // pool mode (L = 0): [launch queries (padded to 128)][n_slices*2][region_cap] private append regions
    int region_cap;              // entries per private region; a full region raises its thread's bound to its own
                                 // (region_cap/2)-th best score and keeps what is above it (pool_region_tighten)
    // in-kernel threshold seeding.  Before its slice every CTA scans the slice's first seed_tiles tiles in sampling
    // mode (best score of every 32-row step), posts its best few per thread and bumps arrive[qblock]; once all
    // n_slices CTAs of the block have posted, the epilogue warps of each CTA fold the posts of their share of the
    // block's queries into gthr[q] (the L-th best post; pool mode: the 32nd best) and raise seeded[q]; every epilogue
    // thread waits (bounded) for its own query's word, then the slice is scanned with that bound.
    int seed_tiles;              // 0 = off
    int seed_stride;             // slices 0, stride, 2*stride, ... sample; the others post nothing (small shards in pool mode)
    unsigned *samples;           // [launch queries (padded to 128)][n_slices*2][2 .. L] ordered score keys, 0 = empty
    unsigned *seeded;            // [all queries] 0 = not seeded yet, 1 = seeded without a bound, else the seed (= gthr[q] then)
    int seed_rank;               // which of the posted scores becomes the bound: the k-th best (list mode; any rank in [k, L] is a
                                 // valid lower bound of the k-th best overall, the k-th is the tightest), the 32nd (pool mode)
    unsigned *arrive;            // [all q-blocks] CTAs that have posted (0 between calls: the query preparation clears it)
    unsigned *tile_counter;      // nullptr: every CTA scans its static slice.  Else (one query block per launch, no pairs): the next
                                 // tile of the shard nobody has taken yet (0 at launch: cleared by the preparation); CTAs take tiles
                                 // one at a time, so a slow SM simply takes fewer and all finish within one tile of each other
    // development knobs (0 in production; b2r_create reads them from the environment)
    unsigned long long seed_wait_ns;   // overrides the wait budget of the seeding phase (1 = do not wait at all)
    int delay_us;                // every third slice sleeps this long before it posts its samples (a slow CTA)
    int trace_mode;              // 1: slot 6 = SM cycles the first epilogue warp waited for full accumulators; 2: slots 4, 5 = globaltimer at the first full accumulator / when the sampling tiles are done;
                                 // 3: slots 4, 5, 7 = MMA thread past the launch wait / queries resident / first full accumulator.  (Nothing of
                                 // this may sit inside the MMA thread's K-block loop: one extra predicated store there cost 14 % at batch 256.)
    unsigned long long *trace;   // [grid][8]: globaltimer at epilogue start, posted, seeded, done; slots 4..7 by trace_mode (nullptr = off).  The
                                 // MMA thread's and the producer's loops carry NO instrumentation: the wait totals they once kept (30 k / 37 k
                                 // cycles per launch for an empty accumulator / for operands, profiles/r2_tensor_pipe.md) were worth knowing once
};

// Up to 512 dims (KB <= 8) the query block stays resident in shared memory (KB * 16 KB) and a pipeline stage
// is one corpus K-block.  Beyond that it would crowd out the ring (768 dims: 192 KB), so the query K-block is
// streamed too: a stage = query K-block (16 KB, an L2 hit every time) + corpus K-block (32 KB).
// Up to 128 KB of resident queries: 512 dims at 128 queries per block, 1024 dims at 64 (the BM = 64 form of small batches).
__host__ __device__ constexpr bool gemm_a_resident(int KB, int bm = 128) { return KB * bm * 128 <= 128 * 1024; }
__host__ __device__ constexpr int gemm_bn(int KB) { return 256; }
// PAIR: two CTAs of a cluster (the two SMs of a TPC) score 256 queries against the same corpus tile with one
// tcgen05.mma.cta_group::2 (M = 256); each CTA stages only HALF of the tile's rows, the tensor core reads both halves,
// so a corpus tile crosses L2 -> shared memory once per pair instead of once per query block.
__host__ __device__ constexpr int gemm_stage_bytes(int KB, bool pair = false, int bm = 128) {
    return (pair ? gemm_bn(KB) / 2 : gemm_bn(KB)) * 128 + (gemm_a_resident(KB, bm) ? 0 : bm * 128);
}
__host__ __device__ constexpr int gemm_stages(int KB, bool pair = false, int bm = 128) {
    int a = gemm_a_resident(KB, bm) ? KB * bm * 128 : 0;
    int s = (GEMM_SMEM_LIMIT - GEMM_SMEM_SLACK - a) / gemm_stage_bytes(KB, pair, bm);
    return s > 8 ? 8 : s;
}
__host__ __device__ constexpr size_t gemm_smem_bytes(int KB, bool pair = false, int bm = 128) {
    return (size_t)(gemm_a_resident(KB, bm) ? KB * bm * 128 : 0) + (size_t)gemm_stages(KB, pair, bm) * gemm_stage_bytes(KB, pair, bm) + 1024;
}

// ---------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Wait with a watchdog: a protocol bug must fail the launch (trap -> CUDA error), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 4000000000ull) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier (bytes)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// ---- CTA pair (cta_group::2) forms ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {      // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the pair's loads complete on the LEADER's mbarrier (rank 0 of the pair: the same shared-memory offset with the peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tm), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
    asm volatile(
        "{\n .reg .b32 ra;\n mapa.shared::cluster.u32 ra, %0, %1;\n mbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n"
        ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t *bar, uint32_t parity) {     // acquire at cluster scope
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 4000000000ull) __trap();
    }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *slot, uint32_t ncols) {   // one warp in EACH CTA of the pair, same slot offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 in each CTA's smem] * B[N rows: N/2 in each CTA's smem]^T; the leader's elected thread issues
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0u;      // disable-output-lane mask: none
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs when every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32; one thread issues for the CTA
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 64 bf16 (128 B),
// 8-row groups 1024 B apart (SBO); version 1 (sm_100); LBO unused for swizzled K-major
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (lane = thread).  The load is
// asynchronous: v[] is defined only after tmem_ld_wait(v), which also pins the compiler's ordering.
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;\n"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
          "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
          "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :: "memory");
}

// three-input maximum (one instruction on sm_100: PTX ISA 8.6 max.f32 d, a, b, c)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// maximum of 8 / of 4 values in 4 / 2 instructions
__device__ __forceinline__ float fmax8(const float *v) {
    return fmax3(fmax3(v[0], v[1], v[2]), fmax3(v[3], v[4], v[5]), fmaxf(v[6], v[7]));
}

// ---------------------------------------------------------------------------------
// per-thread sorted list (descending score; equal scores keep the earlier = lower row first)
// ---------------------------------------------------------------------------------
template <int L>
struct RegList {
    float s[L];
    unsigned r[L];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < L; ++i) { s[i] = -INFINITY; r[i] = 0xffffffffu; }
    }
    // x must beat s[L-1]
    __device__ __forceinline__ void insert(float x, unsigned row) {
#pragma unroll
        for (int i = L - 1; i >= 1; --i) {
            const bool above = x > s[i - 1];       // x ranks before slot i-1: slot i-1 moves down
            const bool here = x > s[i];
            s[i] = above ? s[i - 1] : (here ? x : s[i]);
            r[i] = above ? r[i - 1] : (here ? row : r[i]);
        }
        const bool top = x > s[0];
        s[0] = top ? x : s[0];
        r[0] = top ? row : r[0];
    }
};

// Sampling-pass step: the best passing score among this step's 32 rows goes into the thread's list
// (every lane inserts once per step, in lockstep).  Different steps are different rows, so the list's
// entries are scores of distinct real rows -- all the threshold seed needs.
template <int L, bool HAS_BIAS>
__device__ __forceinline__ void epi_chunk_sample(const uint32_t (&raw)[32], unsigned r0, const GemmParams &p, RegList<L> &list) {
    const unsigned pm = __ldg(p.pass_bits + (r0 >> 5));
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
    if (HAS_BIAS) {
        const float4 *b4 = reinterpret_cast<const float4 *>(p.bias + r0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 b = __ldg(b4 + j);
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
    }
    if (pm != 0xffffffffu) {                              // warp-uniform
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = ((pm >> j) & 1u) ? v[j] : -INFINITY;
    }
    const float x = fmaxf(fmax3(fmax8(v), fmax8(v + 8), fmax8(v + 16)), fmax8(v + 24));
    if (x > list.s[L - 1]) list.insert(x, r0);
}

// One epilogue step: 32 scores of this thread's query (columns r0 .. r0+31 of the corpus).
// Hot path: a max tree (log depth -- a single resident warp per scheduler cannot hide serial chains)
// and one compare.  When any lane of the warp has a hit, every lane walks its own hits lowest column
// first and ALL hitting lanes insert simultaneously (one insert site), so a step costs the max over
// lanes of the hits, not their sum.
template <int L, bool HAS_BIAS>
__device__ __forceinline__ void epi_chunk(const uint32_t (&raw)[32], unsigned r0, const GemmParams &p, RegList<L> &list,
                                          float &thr, unsigned &g_seen, unsigned *gq, bool publish) {
    const unsigned pm = __ldg(p.pass_bits + (r0 >> 5));   // issued early, consumed only on the slow path
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
    if (HAS_BIAS) {
        const float4 *b4 = reinterpret_cast<const float4 *>(p.bias + r0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 b = __ldg(b4 + j);
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
    }
    float m8[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) m8[g] = fmax8(v + 8 * g);
    const float m32 = fmaxf(fmax3(m8[0], m8[1], m8[2]), m8[3]);
    if (!__any_sync(FULL_MASK, m32 > thr)) return;
    // ---- slow path ----
    unsigned hm = 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (m8[g] > thr) {
            unsigned b = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) b |= (v[8 * g + t] > thr ? 1u : 0u) << t;
            hm |= b << (8 * g);
        }
    }
    hm &= pm;                                            // dead / filtered / padding rows never enter
    while (__any_sync(FULL_MASK, hm != 0u)) {
        if (hm) {
            const int j = __ffs(hm) - 1;
            hm &= hm - 1;
            // x = v[j] by a 5-level select tree on the bits of j (depth 5, not a 31-deep chain)
            float w16[16], w8[8], w4[4], w2[2];
            const bool b4 = j & 16, b3 = j & 8, b2 = j & 4, b1 = j & 2, b0 = j & 1;
#pragma unroll
            for (int i = 0; i < 16; ++i) w16[i] = b4 ? v[16 + i] : v[i];
#pragma unroll
            for (int i = 0; i < 8; ++i) w8[i] = b3 ? w16[8 + i] : w16[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) w4[i] = b2 ? w8[4 + i] : w8[i];
#pragma unroll
            for (int i = 0; i < 2; ++i) w2[i] = b1 ? w4[2 + i] : w4[i];
            const float x = b0 ? w2[1] : w2[0];
            if (x > thr) {
                list.insert(x, r0 + j);
                const float lmin = list.s[L - 1];
                if (lmin > -INFINITY) {                  // list full: its L-th best bounds everything it rejects
                    thr = fmaxf(thr, lmin);
                    const unsigned o = KeyS::ord(lmin);
                    if (o > g_seen) { if (publish) atomicMax(gq, o); g_seen = o; }
                }
            }
        }
    }
}

// Pool mode, rare path: this thread's private region is full.  First the entries that no longer beat the query's shared
// bound go (another slice may have raised it); if more than half the region is still in use, the thread raises its own bound to
// its (cap/2)-th best score t -- found by bisection on the 32 bits of the ordered score, no sorting, no scratch -- keeps
// what beats t and publishes t.  Publishing is what keeps the certificate's invariant (a row that is not in the pool
// scores <= the final gthr[q]); it is a useful bound because cap/2 rows of this region alone reach it (cap/2 = 256 >= 2k).
// Strict compares everywhere, as in the admission test, so a region of equal scores still shrinks.
struct PoolTight { int count; float thr; };
static __device__ __noinline__ PoolTight pool_region_tighten(KeyS *region, int count, int cap, float thr, unsigned *gq, bool publish) {
    int n = count < cap ? count : cap;
    unsigned cut = thr > -INFINITY ? KeyS::ord(thr) : 0u;
    const unsigned g = *reinterpret_cast<volatile unsigned *>(gq);
    if (g > cut) {
        cut = g;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const KeyS e = region[i];
            if ((unsigned)(e.v >> 32) > cut) region[m++] = e;
        }
        n = m;
    }
    const int keep = cap / 2;
    if (n > keep) {
        unsigned t = 0u;                                   // largest t with #(key >= t) >= keep  =  the keep-th best key
#pragma unroll 1
        for (int bit = 31; bit >= 0; --bit) {
            const unsigned cand = t | (1u << bit);
            int c = 0;
#pragma unroll 8
            for (int i = 0; i < n; ++i) c += (unsigned)(region[i].v >> 32) >= cand ? 1 : 0;
            if (c >= keep) t = cand;
        }
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const KeyS e = region[i];
            if ((unsigned)(e.v >> 32) > t) region[m++] = e;
        }
        n = m;
        cut = t;
        if (publish) atomicMax(gq, t);
    }
    PoolTight r;
    r.count = n;
    r.thr = cut != 0u ? fmaxf(thr, KeyS::unord(cut)) : thr;
    return r;
}

// Pool-mode step (32 < k <= 128, no per-thread list): every passing row that beats the query's bound --
// seeded by the sampling pass at the 32nd best sampled score -- is appended to this thread's private region
// with a plain store.  With a seed the bound rarely moves during the pass, so the expected pool is
// (rows / sample rows) * 32 entries per query whatever the shard size; a region that fills up anyway (no seed yet,
// or an unlucky one) tightens itself.
template <bool HAS_BIAS>
__device__ __forceinline__ void epi_chunk_pool(const uint32_t (&raw)[32], unsigned r0, const GemmParams &p, float &thr,
                                               KeyS *region, int &count, unsigned *gq, bool publish) {
    const unsigned pm = __ldg(p.pass_bits + (r0 >> 5));
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
    if (HAS_BIAS) {
        const float4 *b4 = reinterpret_cast<const float4 *>(p.bias + r0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 b = __ldg(b4 + j);
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
    }
    float m8[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) m8[g] = fmax8(v + 8 * g);
    const float m32 = fmaxf(fmax3(m8[0], m8[1], m8[2]), m8[3]);
    if (!__any_sync(FULL_MASK, m32 > thr)) return;
    unsigned hm = 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (m8[g] > thr) {
            unsigned b = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) b |= (v[8 * g + t] > thr ? 1u : 0u) << t;
            hm |= b << (8 * g);
        }
    }
    hm &= pm;
    while (__any_sync(FULL_MASK, hm != 0u)) {
        if (hm) {
            const int j = __ffs(hm) - 1;
            hm &= hm - 1;
            float w16[16], w8[8], w4[4], w2[2];
            const bool b4 = j & 16, b3 = j & 8, b2 = j & 4, b1 = j & 2, b0 = j & 1;
#pragma unroll
            for (int i = 0; i < 16; ++i) w16[i] = b4 ? v[16 + i] : v[i];
#pragma unroll
            for (int i = 0; i < 8; ++i) w8[i] = b3 ? w16[8 + i] : w16[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) w4[i] = b2 ? w8[4 + i] : w8[i];
#pragma unroll
            for (int i = 0; i < 2; ++i) w2[i] = b1 ? w4[2 + i] : w4[i];
            const float x = b0 ? w2[1] : w2[0];
            if (count >= p.region_cap) {
                const PoolTight tg = pool_region_tighten(region, count, p.region_cap, thr, gq, publish);
                count = tg.count; thr = tg.thr;
            }
            if (x > thr) region[count++] = KeyS::make(x, r0 + j);
        }
    }
}

// ---------------------------------------------------------------------------------
// finalize: one CTA per query.  Fold the (slice) lists into the best KP by bf16 score, re-rank
// them exactly, certify against max(KP-th best candidate score, final gthr[q]), emit.
// ---------------------------------------------------------------------------------
constexpr int FU_MAX_POOL = 4096;     // pool entries staged in shared memory by the data-parallel selection
constexpr int FU_MAX_SEL = 512;       // survivors of the score cut that are ranked by counting

// bitonic sort of one key per lane, best (largest v) first
__device__ __forceinline__ KeyS warp_sort_desc(KeyS x, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j >= 1; j >>= 1) {
            const KeyS o = KeyS::shfl_xor(x, j);
            const bool lower = (lane & j) == 0;               // this lane keeps the better of the pair ...
            const bool desc = (lane & k) == 0 || k == 32;     // ... in a best-first sub-sequence
            const bool take_better = lower == desc;
            const bool o_better = KeyS::better(o, x);
            if (o_better == take_better) x = o;
        }
    }
    return x;
}
// keys of a best-first sorted run of 32 that rank before x (keys are distinct; worst() pads sort last)
__device__ __forceinline__ int run_count_better(const KeyS *run, const KeyS &x) {
    int lo = 0;                                               // invariant: run[0 .. lo) are better than x
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1)
        if (KeyS::better(run[lo + step - 1], x)) lo += step;
    if (lo == 31 && KeyS::better(run[31], x)) lo = 32;
    return lo;
}

// One query, FIN_THREADS threads (C says which, and how they synchronise), `smem_raw` = finalize_union_smem() bytes.
template <int EPL, class C>
__device__ __forceinline__ void finalize_union_query(const FinalizeParams &fin, const UnionParams &u, int qi,
                                                     unsigned char *smem_raw) {
    constexpr int KP = 32 * EPL;
    KeyS *stage = reinterpret_cast<KeyS *>(smem_raw);                       // [FIN_WARPS*KP]
    KeyD *sm_ex = reinterpret_cast<KeyD *>(stage + FIN_WARPS * KP);        // [KP]
    KeyD *sm_misc = sm_ex + KP;                                            // [4]
    float *sm_q = reinterpret_cast<float *>(sm_misc + 4);                  // [dp]
    KeyS *pool = reinterpret_cast<KeyS *>(sm_q + fin.dp);                  // [FU_MAX_POOL]
    KeyS *sel = pool + FU_MAX_POOL;                                        // [FU_MAX_SEL]
    int *s_count = reinterpret_cast<int *>(sel + FU_MAX_SEL);              // [3] + s_nsel
    int &s_nsel = s_count[3];
    const int tid = C::tid();
    const int lane = tid & 31, warp = tid >> 5;
    // written by other CTAs (or an earlier kernel): L2 loads.  The first 512 pool slots are fetched before the
    // cursor is known (every pool is at least that long), so cursor and entries cost one round trip, not two.
    const KeyS *src = u.lists + (size_t)qi * u.list_stride;
    const KeyS spec0 = KeyS::load_cg(src + tid), spec1 = KeyS::load_cg(src + tid + FIN_THREADS);
    const unsigned cnt_raw = __ldcg(u.cnt + qi);
    const bool overflow = (cnt_raw >> 31) != 0u;          // pool mode: a private region filled up, rows were dropped
    const int entries = overflow ? 0 : min((int)(cnt_raw & 0x7fffffffu), u.max_entries);   // an overflowed pool has holes
    const unsigned g = __ldcg(u.gthr + qi);

    for (int i = tid; i < fin.dp; i += FIN_THREADS) sm_q[i] = fin.q[(size_t)qi * fin.dp + i];
    if (tid == 0) { s_nsel = 0; s_count[0] = s_count[1] = s_count[2] = 0; }
    int nvalid = -1;                                       // -1: the data-parallel selection did not apply

    if (entries <= FU_MAX_POOL) {
        // ---- data-parallel selection: stage the pool, cut at the KP-th best score, rank the survivors ----
        const KeyS e0 = tid < entries ? spec0 : KeyS::worst(), e1 = tid + FIN_THREADS < entries ? spec1 : KeyS::worst();
        if (entries > FU_MAX_SEL) {
            pool[tid] = e0; pool[tid + FIN_THREADS] = e1;
            for (int i = tid + 2 * FIN_THREADS; i < entries; i += FIN_THREADS) pool[i] = KeyS::load_cg(src + i);
        }
        for (int i = tid; i < KP; i += FIN_THREADS) stage[i] = KeyS::worst();
        if (entries <= FU_MAX_SEL) {
            // ---- small pool (the usual case): warp-sorted runs of 32 + binary-search ranks ----
            // thread t holds pool entries t and t + 256; every warp sorts its 32 (best first) into run w / 8 + w;
            // an entry's rank = its position in its own run + the better entries of every other run.
            const int n_runs = (entries + 31) / 32;
            KeyS *runs = sel;                                  // [16][32]
            KeyS a = warp_sort_desc(e0, lane);
            runs[warp * 32 + lane] = a;
            KeyS b = KeyS::worst();
            if (entries > FIN_THREADS) { b = warp_sort_desc(e1, lane); runs[(FIN_WARPS + warp) * 32 + lane] = b; }
            C::sync();
            int ra = lane, rb = lane;
            for (int r = 0; r < n_runs; ++r) {
                const KeyS *run = runs + r * 32;
                if (r != warp) ra += run_count_better(run, a);
                if (r != FIN_WARPS + warp) rb += run_count_better(run, b);
            }
            if (a.valid() && ra < KP) stage[ra] = a;
            if (b.valid() && rb < KP) stage[rb] = b;
            C::sync();
            nvalid = min(entries, KP);
        } else {
        C::sync();
        unsigned t = 0;                                    // keep entries whose score key is >= t
        {
#pragma unroll 1
            for (int bit = 31, it = 0; bit >= 10; --bit, ++it) {   // 22 bits of the ordered score: a lower bound of the KP-th best
                const unsigned cand = t | (1u << bit);
                int c = 0;
                for (int i = tid; i < entries; i += FIN_THREADS) c += (unsigned)(pool[i].v >> 32) >= cand ? 1 : 0;
                c = __reduce_add_sync(FULL_MASK, c);
                if (lane == 0 && c) atomicAdd(&s_count[it % 3], c);
                if (tid == 0) s_count[(it + 1) % 3] = 0;    // last read two rounds ago: one barrier per round
                C::sync();
                if (s_count[it % 3] >= KP) t = cand;
            }
        }
        const KeyS *ranked = pool;
        int nsel = entries;
        {
            for (int i0 = warp * 32; i0 < entries; i0 += FIN_THREADS) {      // warp-aggregated compaction
                const int i = i0 + lane;
                const KeyS k = i < entries ? pool[i] : KeyS::worst();
                const bool keep = i < entries && (unsigned)(k.v >> 32) >= t;
                const unsigned m = __ballot_sync(FULL_MASK, keep);
                int base = 0;
                if (lane == 0 && m) base = atomicAdd(&s_nsel, __popc(m));
                base = __shfl_sync(FULL_MASK, base, 0);
                const int slot = base + __popc(m & ((1u << lane) - 1));
                if (keep && slot < FU_MAX_SEL) sel[slot] = k;
            }
            C::sync();
            nsel = s_nsel;
            ranked = sel;
        }
        if (nsel <= FU_MAX_SEL) {
            for (int i = tid; i < nsel; i += FIN_THREADS) {
                const KeyS me = ranked[i];
                int rank = 0;
#pragma unroll 8
                for (int j = 0; j < nsel; ++j) rank += KeyS::better(ranked[j], me) ? 1 : 0;   // broadcast LDS.64, 8 in flight
                if (rank < KP) stage[rank] = me;
            }
            C::sync();
            nvalid = min(nsel, KP);
        }
        }
    }
    if (nvalid < 0) {
        // ---- general path (very large pools or massive score ties): warp-resident sorted lists ----
        WarpList<KeyS, EPL> wl; wl.init();
        constexpr int UN = 4;                             // loads in flight per lane before the first use
        for (int b = warp * 32; b < entries; b += FIN_WARPS * 32 * UN) {
            KeyS mine[UN];
#pragma unroll
            for (int u2 = 0; u2 < UN; ++u2) {
                const int idx = b + u2 * FIN_WARPS * 32 + lane;
                mine[u2] = idx < entries ? KeyS::load_cg(src + idx) : KeyS::worst();
            }
#pragma unroll
            for (int u2 = 0; u2 < UN; ++u2) {
                unsigned hits = __ballot_sync(FULL_MASK, mine[u2].valid() && wl.accepts(mine[u2]));
                while (hits) {
                    const int sl = __ffs(hits) - 1;
                    hits &= hits - 1;
                    wl.offer(KeyS::shfl(mine[u2], sl), lane);
                }
            }
        }
        C::sync();
        cta_tree_merge<KeyS, EPL, C>(wl, stage, warp, lane);
        nvalid = 0;
        for (int i0 = 0; i0 < KP; i0 += 32) nvalid += __popc(__ballot_sync(FULL_MASK, stage[i0 + lane].valid()));
    }
    if (tid == 0) { atomicAdd(u.pool_stats, 1ull); atomicAdd(u.pool_stats + 1, (unsigned long long)entries); }
    // rows outside the candidate set: either in the pool but below the KP-th candidate, or never kept
    // by any list, hence <= the final shared bound (0 = nothing was ever rejected)
    float T = -INFINITY;
    if (nvalid == KP) T = stage[KP - 1].score();
    if (g != 0u) T = fmaxf(T, KeyS::unord(g));
    if (overflow) T = INFINITY;                             // nothing bounds the dropped rows: force the exact fix-up
    finalize_candidates<C>(fin, qi, stage, nvalid, T, sm_ex, sm_q, sm_misc);
}

// the finalize as its own launch (pool mode's sampling flow, or K3 without the fused tail): one CTA per query
template <int EPL>
__global__ void __launch_bounds__(FIN_THREADS)
finalize_union_kernel(const FinalizeParams fin, const UnionParams u, int q0) {
    extern __shared__ __align__(16) unsigned char fu_smem[];
    pdl_wait();
    pdl_trigger();
    finalize_union_query<EPL, FinCta<0>>(fin, u, q0 + blockIdx.x, fu_smem);
}

inline size_t finalize_union_smem(int EPL, int dp) {
    const int KP = 32 * EPL;
    return sizeof(KeyS) * (size_t)FIN_WARPS * KP + sizeof(KeyD) * (KP + 4) + sizeof(float) * dp +
           sizeof(KeyS) * (size_t)(FU_MAX_POOL + FU_MAX_SEL) + 16;
}

// ---------------------------------------------------------------------------------
// in-kernel threshold seeding
// ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void epi_bar_sync() {       // the 8 epilogue warps only (named barrier 1)
    asm volatile("bar.sync 1, %0;" ::"n"(GEMM_EPI_WARPS * 32) : "memory");
}

// One warp: the L-th largest of vals[0..n) (ordered score keys of DISTINCT rows, 0 = empty), 0 when fewer than
// L are set.  Every lane keeps the L best of its strided share in registers, then the warp pops the
// maximum L times.
// `rank` <= L: the rank-th largest instead (the lanes still keep L each).
template <int L>
__device__ __forceinline__ unsigned warp_lth_largest(const unsigned *vals, int n, int lane, int rank = L) {
    unsigned top[L];
#pragma unroll
    for (int i = 0; i < L; ++i) top[i] = 0u;
    constexpr int UN = 16;                               // independent L2 loads in flight per lane (the usual 296-592 posts: one or two rounds)
    for (int b = lane; b < n; b += 32 * UN) {
        unsigned x[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) x[u] = b + 32 * u < n ? __ldcg(vals + b + 32 * u) : 0u;
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            if (x[u] > top[L - 1]) {
#pragma unroll
                for (int i = L - 1; i >= 1; --i) {
                    const bool above = x[u] > top[i - 1];
                    top[i] = above ? top[i - 1] : (x[u] > top[i] ? x[u] : top[i]);
                }
                top[0] = x[u] > top[0] ? x[u] : top[0];
            }
        }
    }
    unsigned res = 0u;
#pragma unroll 1
    for (int r = 0; r < rank; ++r) {
        res = __reduce_max_sync(FULL_MASK, top[0]);
        const unsigned who = __ballot_sync(FULL_MASK, top[0] == res);
        if (lane == __ffs(who) - 1) {                    // the winning lane advances to its next best
#pragma unroll
            for (int i = 0; i < L - 1; ++i) top[i] = top[i + 1];
            top[L - 1] = 0u;
        }
    }
    return res;
}

// ---------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------
// PAIR = launched as clusters of two CTAs (query blocks 2j and 2j+1 of the same slice): the even CTA (cluster rank 0) is the
// leader -- it owns the full / accumulator-empty barriers and issues every tcgen05.mma.cta_group::2 for the pair; both CTAs
// load (their own query block, their half of each corpus tile) and both run the epilogue on their own 128 TMEM lanes.
// BM = queries per CTA: 128, or 64 for batches of at most 64 queries (tcgen05.mma M = 64: rows 16i .. 16i+15 of the block sit on
// TMEM lanes 32i .. 32i+15, so lanes 0..15 of every epilogue warp own a query and the other 16 idle).  Half the tensor work for
// the same corpus stream -- on a power-capped GPU that is bandwidth (10M x 768, batch 64, sustained: 3.03 ms with M = 128 against
// 2.50 ms for a batch of one) -- and twice the dims fit as resident queries (768 and 1024 dims stop re-fetching the query block
// with every corpus K-block).
template <int KB, int L, bool HAS_BIAS, bool PAIR, int BM = 128>
// 10 warps = 3 on some SM sub-partition, whose register file is 16K: 168 registers per thread is the hard cap
// (measured: __maxnreg__(192) compiles without spills but cannot launch)
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_x, const GemmParams p) {
    constexpr int BN = gemm_bn(KB);
    static_assert(BM == 128 || (BM == 64 && !PAIR), "64-query blocks exist only in the single-CTA form");
    constexpr int STAGES = gemm_stages(KB, PAIR, BM);
    constexpr bool A_RES = gemm_a_resident(KB, BM);
    constexpr int QL = BM == 128 ? 32 : 16;                 // epilogue lanes per warp that own a query
    constexpr int NCTA = PAIR ? 2 : 1;
    constexpr int B_ROWS = BN / NCTA;                       // corpus rows of a tile this CTA stages (tm_x's box has this many rows)
    constexpr uint32_t A_KB_BYTES = BM * 128;               // one K-block of the query block
    constexpr uint32_t B_STAGE_BYTES = B_ROWS * 128;        // one K-block of this CTA's share of a corpus tile
    constexpr uint32_t STAGE_BYTES = gemm_stage_bytes(KB, PAIR, BM);  // ring slot: corpus K-block (+ query K-block when streamed)
    constexpr uint32_t TMEM_COLS = 2 * BN;
    constexpr uint32_t IDESC = umma_idesc_bf16(BM * NCTA, BN);
    static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");

    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_a, bar_full[STAGES], bar_empty[STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ int tile_ring[8];          // dynamic tiles: the tile iteration `it` works on, at [it & 7] (bit 30 = sampling pass, -1 = no tile left)
    constexpr int TILE_SMP = 1 << 30;

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (base - smem_u32(smem_raw));
    unsigned char *smA = sm;                                  // resident: [KB][128 rows][128 B]
    unsigned char *smB = sm + (A_RES ? (size_t)KB * A_KB_BYTES : 0);   // [STAGES][BN rows][128 B] (+ [128 rows][128 B] streamed A)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;     // blockIdx.x & 1: the grid is launched in clusters of 2 along x
    const bool leader = crank == 0u;
    const int qb = p.qblock0 + blockIdx.x % p.n_qblocks, slice = blockIdx.x / p.n_qblocks;
    const int t0 = (int)((long long)p.tiles_total * slice / p.n_slices);
    const int t1 = (int)((long long)p.tiles_total * (slice + 1) / p.n_slices);
    // in-kernel threshold seeding: sampling tiles in front of the slice (every seed_stride-th slice samples, all post)
    const int S = (p.seed_tiles > 0 && slice % p.seed_stride == 0) ? min(p.seed_tiles, t1 - t0) : 0;
    const int n_iter = S + (t1 - t0);
    const bool dyn = !PAIR && p.tile_counter != nullptr;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_x);
        mbar_init(&bar_a, 1);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_tfull[i], 1); mbar_init(&bar_tempty[i], GEMM_EPI_WARPS * NCTA); }
        fence_barrier_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc_pair(&tmem_slot, TMEM_COLS); else tmem_alloc(&tmem_slot, TMEM_COLS); }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();     // the peer's barriers and TMEM exist before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            auto load = [&](void *dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1) {
                if (PAIR) tma_load_2d_pair(dst, tm, bar, c0, c1); else tma_load_2d(dst, tm, bar, c0, c1);
            };
            // The corpus does not depend on the kernel this one was launched behind (the query preparation, or the pass-bitmap
            // kernel after it: every kernel of the library waits for ITS predecessor before it lets its successor start, so all
            // ingests are complete by now).  With resident queries the first ring slots are therefore filled BEFORE the
            // programmatic-launch wait: the first tile is in shared memory when the prepared queries arrive.
            int early = 0;
            if (A_RES && !dyn) {
                early = min(STAGES, n_iter * KB);
                for (int j = 0; j < early; ++j) {
                    const int i = j / KB, kb = j % KB;
                    const int t = i < S ? t0 + i : t0 + i - S;
                    if (leader) mbar_expect_tx(&bar_full[j], NCTA * STAGE_BYTES);
                    load(smB + (size_t)j * STAGE_BYTES, &tm_x, &bar_full[j], kb * 64, t * BN + (int)crank * B_ROWS);
                }
            }
            pdl_wait();          // the prepared queries are read below
            pdl_trigger();
            if (A_RES) {
                if (leader) mbar_expect_tx(&bar_a, NCTA * KB * A_KB_BYTES);
                for (int kb = 0; kb < KB; ++kb) load(smA + (size_t)kb * A_KB_BYTES, &tm_q, &bar_a, kb * 64, qb * BM);
            }
            int stage = 0; uint32_t phase = 0;
            if (dyn) {
                // ---- dynamic tiles: take the next free tile of the shard, tell the other roles which one it is ----
                unsigned *ctr = p.tile_counter + p.qblock0;
                int it = 0, ns = 0, samp[8];
                auto issue = [&](int t, int tag) {
                    *reinterpret_cast<volatile int *>(&tile_ring[it & 7]) = t | tag;     // before the first arrive of the tile (release)
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&bar_empty[stage], phase ^ 1);
                        mbar_expect_tx(&bar_full[stage], STAGE_BYTES);
                        if (!A_RES) load(smB + (size_t)stage * STAGE_BYTES + B_STAGE_BYTES, &tm_q, &bar_full[stage], kb * 64, qb * BM);
                        load(smB + (size_t)stage * STAGE_BYTES, &tm_x, &bar_full[stage], kb * 64, t * BN);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    ++it;
                };
                for (; ns < min(p.seed_tiles, 8); ++ns) {          // the sampling pass: this CTA's first few tiles ...
                    const int t = (int)atomicAdd(ctr, 1u);
                    if (t >= p.tiles_total) break;
                    samp[ns] = t;
                    issue(t, TILE_SMP);
                }
                for (int i = 0; i < ns; ++i) issue(samp[i], 0);     // ... which the main pass scans again with the bound in place
                for (;;) {
                    const int t = (int)atomicAdd(ctr, 1u);
                    if (t >= p.tiles_total) break;
                    issue(t, 0);
                }
                *reinterpret_cast<volatile int *>(&tile_ring[it & 7]) = -1;              // nothing left: wake the MMA thread with an empty slot
                mbar_wait(&bar_empty[stage], phase ^ 1);
                mbar_arrive(&bar_full[stage]);
            } else
            for (int i = 0; i < n_iter; ++i) {
                const int t = i < S ? t0 + i : t0 + i - S;      // the seeding tiles are scanned again by the main loop
                for (int kb = 0; kb < KB; ++kb) {
                    if (i * KB + kb < early) {                  // issued above
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    if (leader) mbar_expect_tx(&bar_full[stage], NCTA * STAGE_BYTES);     // both CTAs' bytes land on the leader's barrier
                    if (!A_RES) load(smB + (size_t)stage * STAGE_BYTES + B_STAGE_BYTES, &tm_q, &bar_full[stage], kb * 64, qb * BM);
                    load(smB + (size_t)stage * STAGE_BYTES, &tm_x, &bar_full[stage], kb * 64, t * BN + (int)crank * B_ROWS);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (the leader's elected thread issues for the pair) =====
        pdl_wait();
        pdl_trigger();
        if (lane == 0 && leader) {
            auto commit = [&](uint64_t *bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
            if (p.trace && p.trace_mode == 3) p.trace[(size_t)blockIdx.x * 8 + 4] = globaltimer_ns();      // MMA thread past the launch wait
            if (A_RES) { mbar_wait(&bar_a, 0); tc_fence_after(); }
            if (p.trace && p.trace_mode == 3) p.trace[(size_t)blockIdx.x * 8 + 5] = globaltimer_ns();      // queries resident
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; dyn || it < n_iter; ++it) {
                const int buf = it & 1;
                bool last = false;
                // (PAIR: the peer's epilogue warps arrive remotely.  Default-scope arrive / try_wait, as CUTLASS's 2-SM pipelines use:
                // the cluster-scope release/acquire forms cost ~1500 cycles per tile here; what is handed over is TMEM, ordered by
                // the tcgen05 fences on both sides)
                mbar_wait(&bar_tempty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&bar_full[stage], phase);
                    if (dyn && kb == 0 && *reinterpret_cast<volatile int *>(&tile_ring[it & 7]) < 0) { last = true; break; }
                    tc_fence_after();
                    const uint64_t ad = umma_smem_desc(smem_u32(A_RES ? smA + (size_t)kb * A_KB_BYTES
                                                                          : smB + (size_t)stage * STAGE_BYTES + B_STAGE_BYTES));
                    const uint64_t bd = umma_smem_desc(smem_u32(smB + (size_t)stage * STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {    // UMMA_K = 16 bf16 = 32 B: +2 in the (>>4) address field
                        if (PAIR) umma_bf16_ss_pair(d, ad + 2 * k, bd + 2 * k, IDESC, (kb | k) != 0 ? 1u : 0u);
                        else umma_bf16_ss(d, ad + 2 * k, bd + 2 * k, IDESC, (kb | k) != 0 ? 1u : 0u);
                    }
                    commit(&bar_empty[stage]);                  // smem slot free (in both CTAs) once these MMAs retire
                    if (kb == KB - 1) commit(&bar_tfull[buf]);  // accumulator complete (in both CTAs)
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (last) { commit(&bar_tfull[buf]); break; }   // no tile left: pass the end marker on to the epilogue
            }
        }
    } else {
        // ===== epilogue: thread = query, column = corpus row =====
        pdl_wait();          // bounds / cursors / seed flags (cleared by the preparation) and the pass bitmap are read below
        pdl_trigger();
        constexpr int NC = BN / 32 / GEMM_HALVES;                    // 32-column steps per tile per warp
        static_assert(NC % 2 == 0, "steps are processed in double-buffered pairs");
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;                            // which half of a tile's columns
        const bool owner = lane < QL;                                // BM = 64: lanes 16..31 of a TMEM quadrant hold no row
        const int q = qb * BM + quad * QL + (lane & (QL - 1));       // (an idle lane aliases a real query's words, read-only)
        const bool publish = owner && q < p.nq;
        constexpr int LL = L > 0 ? L : 1;                          // pool mode (L = 0) keeps no list during the scan ...
        constexpr int LS = L > 0 ? L : GEMM_POOL_SAMPLE_RANK;      // ... but seeds its bound from the 32nd best sample
        RegList<LL> list; list.init();
        float thr = publish ? -INFINITY : INFINITY;                  // padding / idle lanes admit nothing
        KeyS *region = nullptr;
        int rcount = 0;
        if (L == 0)
            region = p.regions + ((size_t)(q - p.qblock0 * BM) * (p.n_slices * GEMM_HALVES) + (size_t)(slice * GEMM_HALVES + half)) * p.region_cap;
        unsigned g_seen = 0;
        unsigned *gq = p.gthr + q;
        unsigned g_next = *reinterpret_cast<volatile unsigned *>(gq);

        // ---- seeding state (see GemmParams::seed_tiles) ----
        // This warp folds the queries lq_next, lq_next + 8, ... < lq1 of the block (the CTA's share is [lq0, lq1)) once every
        // slice of the block has posted.  fold_pending is warp-uniform.
        // values posted per thread: its 2 (4) best when the block's sampling slices then still post >= 4 L (2 L) between
        // them -- the L-th best of that union is as good a bound -- else all it has (S * NC step maxima at most)
        const int n_samp = (p.n_slices + p.seed_stride - 1) / p.seed_stride;      // slices that sample
        const int pv = n_samp >= LS ? 2 : n_samp * GEMM_HALVES >= LS ? 4 : min(LS, (p.seed_tiles * NC + 3) & ~3);
        const int per_q = p.n_slices * GEMM_HALVES * pv;
        const int lq1 = BM * (slice + 1) / p.n_slices;
        int lq_next = BM * slice / p.n_slices + (warp - 2);
        bool fold_pending = false;
        auto fold_ready = [&]() -> bool {                  // have all slices of this block posted?
            unsigned a = 0u;
            if (lane == 0) a = ld_acquire_gpu(p.arrive + qb);
            return __shfl_sync(FULL_MASK, a, 0) >= (unsigned)p.n_slices;
        };
        auto fold_share = [&]() {                          // the L-th best posted score becomes the query's bound
            for (; lq_next < lq1; lq_next += GEMM_EPI_WARPS) {
                const int qs = qb * BM + lq_next;
                if (qs >= p.nq) break;
                const unsigned v = warp_lth_largest<LS>(p.samples + (size_t)(qs - p.qblock0 * BM) * per_q, per_q, lane, p.seed_rank);
                if (lane == 0) {
                    atomicExch(p.seeded + qs, v != 0u ? v : 1u);  // flag and seed in one word (1 = no seed: too few rows pass): the waiting threads take it from here
                    if (v != 0u) atomicMax(p.gthr + qs, v);       // the finalize (and every later tile) reads the bound from gthr[q]
                }
            }
            fold_pending = false;
        };

        const bool tracer = p.trace != nullptr && warp == 2 && lane == 0;
        long long w_tfull = 0;
        // dynamic tiles: which tile does iteration `it` hold (bit 30 = sampling pass, -1 = none left)?  Waits for its accumulator.
        auto peek = [&](int it) -> int {
            mbar_wait(&bar_tfull[it & 1], (it >> 1) & 1);
            return *reinterpret_cast<volatile int *>(&tile_ring[it & 7]);
        };
        // one tile: TMEM -> registers in 32-column steps, two register buffers so the next load flies under this step
        auto run_tile = [&](int t, int it, auto sample_c, auto &slist) {
            constexpr bool SMP = decltype(sample_c)::value;
            const int buf = it & 1;
            if (!SMP && fold_pending && fold_ready()) fold_share();     // a fold that had to be put off (a slice was late)
            if (!SMP && g_next > g_seen) { g_seen = g_next; thr = fmaxf(thr, KeyS::unord(g_next)); }
            if (tracer) {
                const long long c0 = clock64(); mbar_wait(&bar_tfull[buf], (it >> 1) & 1); w_tfull += clock64() - c0;
                if (it == 0 && p.trace_mode >= 2) p.trace[(size_t)blockIdx.x * 8 + (p.trace_mode == 2 ? 4 : 7)] = globaltimer_ns();
            } else mbar_wait(&bar_tfull[buf], (it >> 1) & 1);
            tc_fence_after();
            // the other slices' progress on this query: loaded now, consumed at the top of the next tile
            if (!SMP) g_next = *reinterpret_cast<volatile unsigned *>(gq);
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN + half * NC * 32);
            const unsigned row0 = (unsigned)t * BN + (unsigned)(half * NC * 32);
            uint32_t va[32], vb[32];
            tmem_ld_x32(trow, va);
#pragma unroll 1
            for (int c = 0; c < NC; c += 2) {
                tmem_ld_wait(va);
                tmem_ld_x32(trow + (c + 1) * 32, vb);
                if constexpr (SMP) epi_chunk_sample<LS, HAS_BIAS>(va, row0 + c * 32, p, slist);
                else if constexpr (L == 0) epi_chunk_pool<HAS_BIAS>(va, row0 + c * 32, p, thr, region, rcount, gq, publish);
                else epi_chunk<LL, HAS_BIAS>(va, row0 + c * 32, p, list, thr, g_seen, gq, publish);
                __syncwarp();
                tmem_ld_wait(vb);
                if (c + 2 < NC) tmem_ld_x32(trow + (c + 2) * 32, va);
                if constexpr (SMP) epi_chunk_sample<LS, HAS_BIAS>(vb, row0 + (c + 1) * 32, p, slist);
                else if constexpr (L == 0) epi_chunk_pool<HAS_BIAS>(vb, row0 + (c + 1) * 32, p, thr, region, rcount, gq, publish);
                else epi_chunk<LL, HAS_BIAS>(vb, row0 + (c + 1) * 32, p, list, thr, g_seen, gq, publish);
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_remote(&bar_tempty[buf], 0u); else mbar_arrive(&bar_tempty[buf]); }
        };

        if (tracer) p.trace[(size_t)blockIdx.x * 8 + 0] = globaltimer_ns();
        int it = 0;
        if (p.seed_tiles > 0) {
            // ---- seeding phase ----
            // 1. the first S tiles of the slice in sampling mode: the list collects the best step maxima
            const unsigned long long t_begin = __shfl_sync(FULL_MASK, globaltimer_ns(), 0);   // warp-uniform clock readings
            RegList<LS> slist; slist.init();
            {   // A sampling step consumes its word of the pass bitmap at once (the main loop only on a hit): fetch the words of
                // the sampling tiles now, while the first tile is still on its way, or every step pays a cold DRAM round trip
                const uint32_t *pb = p.pass_bits + (((unsigned)t0 * BN) >> 5);
                unsigned warm = 0u;
                for (int w = lane; w < S * (BN / 32); w += 32) warm |= __ldg(pb + w);
                asm volatile("" ::"r"(warm));
            }
            if (dyn) {
                for (int tv; (tv = peek(it)) >= 0 && (tv & TILE_SMP); ++it) run_tile(tv & ~TILE_SMP, it, std::true_type(), slist);
            } else {
                for (int i = 0; i < S; ++i, ++it) run_tile(t0 + i, it, std::true_type(), slist);
            }
            if (tracer && p.trace_mode == 2) p.trace[(size_t)blockIdx.x * 8 + 5] = globaltimer_ns();
            if (p.delay_us > 0 && slice % 3 == 1) {            // development: a slow CTA
                const unsigned long long t_d = globaltimer_ns();
                while (globaltimer_ns() - t_d < (unsigned long long)p.delay_us * 1000ull) __nanosleep(256);
            }
            // 2. post them, count this CTA in
            if (publish) {
                unsigned *dst = p.samples + ((size_t)(q - p.qblock0 * BM) * (p.n_slices * GEMM_HALVES) +
                                             (size_t)(slice * GEMM_HALVES + half)) * pv;
#pragma unroll
                for (int i = 0; i < LS; i += 2) {
                    if (i >= pv) break;
                    uint2 o;
                    o.x = slist.r[i] != 0xffffffffu ? KeyS::ord(slist.s[i]) : 0u;
                    o.y = slist.r[i + 1] != 0xffffffffu ? KeyS::ord(slist.s[i + 1]) : 0u;
                    *reinterpret_cast<uint2 *>(dst + i) = o;
                }
            }
            epi_bar_sync();                                   // every epilogue thread's post happens before ...
            if (warp == 2 && lane == 0) { __threadfence(); atomicAdd(p.arrive + qb, 1u); }   // ... this (cumulative) fence + arrival
            const unsigned long long t_post = __shfl_sync(FULL_MASK, globaltimer_ns(), 0);
            if (tracer) p.trace[(size_t)blockIdx.x * 8 + 1] = t_post;
            unsigned long long budget = GEMM_SEED_TIMEOUT_NS;
            if (L == 0) budget = min(max(2ull * (t_post - t_begin), GEMM_POOL_WAIT_MIN_NS), GEMM_POOL_WAIT_MAX_NS);
            if (p.seed_wait_ns) budget = p.seed_wait_ns;
            // 3. this CTA's share of the block's queries, one per epilogue warp, as soon as every slice has posted.  The
            //    wait is bounded; a fold that cannot run now is retried at the top of every tile of the main loop.
            fold_pending = lq_next < lq1 && qb * BM + lq_next < p.nq;
            while (fold_pending) {
                if (fold_ready()) { fold_share(); break; }
                if (__shfl_sync(FULL_MASK, (int)(globaltimer_ns() - t_post > budget), 0)) break;
                __nanosleep(64);
            }
            // 4. wait for this thread's own query (another CTA folds it): even with one live lane per warp an unseeded
            //    tile costs more than the wait (measured: batch 1 144 vs 149 us, batch 16 151 vs 172 us per call).  A seed
            //    that comes later still arrives through gthr[q], which every tile re-reads.
            if (publish) {
                unsigned sd;
                while ((sd = ld_acquire_gpu(p.seeded + q)) == 0u) {
                    if (globaltimer_ns() - t_post > budget) break;
                    __nanosleep(32);
                }
                if (sd > 1u) g_next = sd;
            }
            __syncwarp();
        }
        if (tracer) p.trace[(size_t)blockIdx.x * 8 + 2] = globaltimer_ns();
        if (dyn) {
            for (int tv; (tv = peek(it)) >= 0; ++it) run_tile(tv, it, std::false_type(), list);
        } else {
            for (int t = t0; t < t1; ++t, ++it) run_tile(t, it, std::false_type(), list);
        }
        if (fold_pending && fold_ready()) fold_share();          // last chance before this CTA leaves
        if (tracer) { p.trace[(size_t)blockIdx.x * 8 + 3] = globaltimer_ns(); if (p.trace_mode < 3) p.trace[(size_t)blockIdx.x * 8 + 6] = (unsigned long long)w_tfull; }

        if (L == 0) {
            if (publish && rcount > 0) {      // compact the private region into the query's pool
                // entries that do not reach the bound published so far need not travel: whatever is dropped here scores
                // <= the final gthr[q], which is all the certificate asks of a row outside the pool
                const unsigned g_now = *reinterpret_cast<volatile unsigned *>(gq);
                int nv = 0;
                for (int i = 0; i < rcount; ++i) nv += (unsigned)(region[i].v >> 32) >= g_now ? 1 : 0;
                if (nv) {
                    const unsigned base = atomicAdd(&p.cnt[q], (unsigned)nv) & 0x7fffffffu;
                    if (base + (unsigned)nv <= (unsigned)p.list_stride) {
                        KeyS *dst = p.lists + (size_t)q * p.list_stride + base;
                        for (int i = 0, j = 0; i < rcount; ++i) {
                            const KeyS e = region[i];
                            if ((unsigned)(e.v >> 32) >= g_now) dst[j++] = e;
                        }
                    } else {
                        atomicOr(&p.cnt[q], 0x80000000u);   // the pool is full, rows were dropped: not certifiable -> exact fix-up
                    }
                }
            }
        } else if (publish) {     // append this thread's entries to the query's candidate pool (compact: most lists are short)
            // entries below the bound published so far need not travel: whatever is dropped here scores
            // <= the final gthr[q], which is all the certificate asks of a row outside the pool
            int nv = 0;
            const unsigned g_now = *reinterpret_cast<volatile unsigned *>(gq);
#pragma unroll
            for (int i = 0; i < LL; ++i) nv += (list.r[i] != 0xffffffffu && KeyS::ord(list.s[i]) >= g_now) ? 1 : 0;
            if (nv) {
                KeyS *dst = p.lists + (size_t)q * p.list_stride + atomicAdd(&p.cnt[q], (unsigned)nv);
#pragma unroll
                for (int i = 0; i < LL; ++i)             // the list is sorted: valid entries are the first nv
                    if (i < nv) dst[i] = KeyS::make(list.s[i], list.r[i]);
            }
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();     // neither CTA of a pair leaves while the other may still reach its memory
    if (warp == 1) { tc_fence_after(); if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
}

}  // namespace b2r
