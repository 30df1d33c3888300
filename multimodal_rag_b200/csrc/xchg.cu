// Cross-shard exchange without a collective library on the data path (K4x).
//
// Row-sharded corpus, one process per GPU (include/b2r.h, "Row-sharded corpus").  After its local exact top-k every rank
// holds [nq, k] (global row, fp64 distance) + [nq] counts.  Every rank owns a mailbox: device memory allocated by this library
// and mapped into the other processes with CUDA IPC.  A batch is exchanged in two stream-ordered steps, neither of which ever
// keeps a CTA spinning on another GPU:
//   push   (b2r_xchg_push, on the stream of the scan)  one small kernel stores this rank's lists straight into every peer's
//          mailbox over NVLink; its last CTA fences at system scope and writes this call's sequence number into the peers'
//          arrival word for (slot, this rank).
//   merge  (b2r_xchg_merge, on the same or on another stream)  the stream itself waits until the arrival words of all peers
//          have reached the sequence number (cuStreamWaitValue32: a stream memory operation, no SM is occupied while waiting),
//          then one CTA per query merges the world x k candidates on (fp64 distance, global row) -- the order a single shard
//          uses -- and the last CTA tells every peer that the slot has been read.
// Four mailbox slots are used round-robin by sequence number; before a slot is written again the pushing stream waits (again a
// stream memory operation) until every peer has reported having read what the slot held four calls ago.  Why no spinning: the scoring kernel
// of the NEXT batch owns every SM's shared memory; a kernel that waits inside an SM either keeps that SM from starting its slice or
// cannot start itself (measured: a fused spin-wait version cost +31 us per batch on one stream and up to +50 us with outliers on a
// side stream, NCCL's all_gather + a merge launch +26 us / +8 us).  Calls are collective: every rank makes the same sequence of
// push / merge calls with the same nq and k.
//
// Fused form (b2r_query_push): no push kernel at all.  The kernels of the query store every final list into the mailboxes
// as they emit it (finalize.cuh, emit_sorted), the call's last kernel publishes the arrival words on its way out, and the first
// kernel of the call holds the stream until the slot's previous contents have been read everywhere (acks; four slots, so that is
// normally true long before).  b2r_xchg_merge for such a batch checks the arrival words INSIDE the merge kernel (a bounded spin
// on local memory): the caller enqueues the merge of batch i after the kernels of batch i+1, by which time the lists have
// arrived, so neither a stream memory operation nor a collective kernel ever stands between two scans.
#include <algorithm>
#include <cstring>
#include <new>

#include "engine.h"
#include "xchg.cuh"

using namespace b2r;

#define XCHG_REQUIRE(cond, msg)                            \
    do {                                                   \
        if (!(cond)) { set_error(msg); return B2R_EINVAL; } \
    } while (0)

namespace {

// this rank's lists -> every mailbox (its own included), then the arrival word of (slot, this rank) on every rank
__global__ void __launch_bounds__(XCHG_THREADS) xchg_push_kernel(const XchgDev p) {
    __shared__ unsigned s_ticket;
    pdl_wait();                       // the local results come from the kernels before this one on the stream
    pdl_trigger();
    const int per_q = p.k + 1;        // k (row, distance) pairs + the count
    const long long total = (long long)p.nq * per_q;
    for (long long t = (long long)blockIdx.x * XCHG_THREADS + threadIdx.x; t < total; t += (long long)gridDim.x * XCHG_THREADS) {
        const int q = (int)(t / per_q), i = (int)(t % per_q);
        const size_t e_off = (size_t)p.slot * p.slot_bytes + ((size_t)p.rank * p.nq_max + q) * p.entry_bytes;
        if (i < p.k) {
            const long long r = p.rows[(size_t)q * p.k + i];
            const double d = p.d64[(size_t)q * p.k + i];
            for (int w = 0; w < p.world; ++w) {
                char *e = p.box[w] + e_off;
                reinterpret_cast<long long *>(e)[i] = r;
                reinterpret_cast<double *>(e + (size_t)p.k_max * 8)[i] = d;
            }
        } else {
            const int c = p.cnt[q];
            for (int w = 0; w < p.world; ++w) *reinterpret_cast<int *>(p.box[w] + e_off + (size_t)p.k_max * 16) = c;
        }
    }
    // the last CTA publishes: every store of this grid is ordered before its release (threadfence + ticket), and the release is
    // at system scope because the peers' streams read the word
    unsigned *ticket = flag_words(p, p.rank) + 2 * XCHG_SLOTS * p.world + XCHG_TICKET_PUSH;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    if (s_ticket == gridDim.x - 1) {
        if (threadIdx.x == 0) *ticket = 0u;
        if (threadIdx.x < p.world) {
            __threadfence_system();
            st_release_sys(flag_words(p, threadIdx.x) + p.slot * p.world + p.rank, p.seq);
        }
    }
}

// one CTA per query
__global__ void __launch_bounds__(XCHG_THREADS) xchg_merge_kernel(const XchgDev p) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ int s_valid;
    __shared__ unsigned s_ticket;
    pdl_wait();
    pdl_trigger();
    if (p.spin) xchg_wait_arrivals(p);      // fused pushes; otherwise the stream itself has waited (cuStreamWaitValue32)
    xchg_merge_query(p, blockIdx.x, sm, &s_valid);
    xchg_publish_ack(p, &s_ticket);
}

// the flag words a fused call left behind, when the next call on the stream is not another fused call
__global__ void xchg_flush_kernel(const XchgFlags f) {
    pdl_wait();
    pdl_trigger();
    xchg_publish_flags(f, 0);
}

typedef CUresult (*wait32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
wait32_fn get_wait32() {
    static wait32_fn fn = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    if (fn) return fn;
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !sym) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<wait32_fn>(sym);
    return fn;
}

}  // namespace

struct b2r_xchg {
    int device = 0, rank = 0, world = 1, nq_max = 0, k_max = 0;
    size_t entry_bytes = 0, slot_bytes = 0, flags_off = 0, total_bytes = 0;
    char *local = nullptr;
    char *box[XCHG_MAX_WORLD] = {};
    bool opened = false;
    unsigned seq_push = 0, seq_merge = 0;
    bool fused[XCHG_SLOTS] = {};        // how the batch sitting in each slot was pushed
    int nq_of[XCHG_SLOTS] = {}, k_of[XCHG_SLOTS] = {};   // ... and its shape
    XchgFlags pending = {};             // flag words the last fused call left for the next kernel on its stream
    int sm_count = 0;
    std::mutex mu;
};

static void fill(const b2r_xchg *x, XchgDev &p, int nq, int k, unsigned seq) {
    p.rank = x->rank; p.world = x->world; p.nq = nq; p.k = k;
    p.seq = seq; p.slot = (int)(seq % XCHG_SLOTS); p.spin = 0; p.ticket = XCHG_TICKET_MERGE;
    p.nq_max = x->nq_max; p.k_max = x->k_max;
    p.entry_bytes = x->entry_bytes; p.slot_bytes = x->slot_bytes; p.flags_off = x->flags_off;
    for (int r = 0; r < XCHG_MAX_WORLD; ++r) p.box[r] = x->box[r];
    p.rows = nullptr; p.d64 = nullptr; p.cnt = nullptr; p.out_rows = nullptr; p.out_dist = nullptr; p.out_cnt = nullptr;
}

// caller holds x->mu
static int flush_pending(b2r_xchg *x, cudaStream_t s) {
    if (!x->pending.arrive_seq && !x->pending.ack_seq) return B2R_OK;
    B2R_CUDA(launch_pdl(xchg_flush_kernel, dim3(1), dim3(32), 0, s, x->pending));
    x->pending.arrive_seq = 0; x->pending.ack_seq = 0;
    return B2R_OK;
}

// make `stream` wait until word `index` of this rank's flag block has reached `value` (wrap-safe >=)
static int stream_wait(b2r_xchg *x, size_t index, unsigned value, cudaStream_t stream) {
    wait32_fn w = get_wait32();
    if (!w) { set_error("cuStreamWaitValue32 is not available from this driver"); return B2R_ECUDA; }
    const CUdeviceptr addr = (CUdeviceptr)(uintptr_t)(x->local + x->flags_off + index * sizeof(unsigned));
    const CUresult r = w((CUstream)stream, addr, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) { set_error("cuStreamWaitValue32 failed with CUresult " + std::to_string((int)r)); return B2R_ECUDA; }
    return B2R_OK;
}

extern "C" int b2r_xchg_create(int device, int rank, int world, int nq_max, int k_max, b2r_xchg_handle *out) {
    XCHG_REQUIRE(out, "b2r_xchg_create: out is NULL");
    *out = nullptr;
    XCHG_REQUIRE(world >= 1 && world <= XCHG_MAX_WORLD && rank >= 0 && rank < world, "b2r_xchg_create: bad rank / world (world <= 8)");
    XCHG_REQUIRE(nq_max >= 1 && k_max >= 1 && (size_t)world * k_max <= 8192, "b2r_xchg_create: bad sizes (world * k_max <= 8192)");
    B2R_CUDA(cudaSetDevice(device));
    b2r_xchg *x = new (std::nothrow) b2r_xchg();
    if (!x) { set_error("b2r_xchg_create: host allocation failed"); return B2R_ENOMEM; }
    x->device = device; x->rank = rank; x->world = world; x->nq_max = nq_max; x->k_max = k_max;
    x->entry_bytes = ((size_t)k_max * 16 + 4 + 15) / 16 * 16;
    x->slot_bytes = (size_t)world * nq_max * x->entry_bytes;
    x->flags_off = XCHG_SLOTS * x->slot_bytes;
    x->total_bytes = x->flags_off + sizeof(unsigned) * ((size_t)2 * XCHG_SLOTS * world + 4);
    cudaError_t e = cudaMalloc(&x->local, x->total_bytes);
    if (e != cudaSuccess) { delete x; B2R_CUDA(e); }
    e = cudaMemset(x->local, 0, x->total_bytes);
    if (e != cudaSuccess) { cudaFree(x->local); delete x; B2R_CUDA(e); }
    B2R_CUDA(cudaDeviceSynchronize());
    x->box[rank] = x->local;
    cudaDeviceProp prop;
    B2R_CUDA(cudaGetDeviceProperties(&prop, device));
    x->sm_count = prop.multiProcessorCount;
    x->opened = world == 1;
    *out = x;
    return B2R_OK;
}

extern "C" int b2r_xchg_ipc_handle(b2r_xchg_handle x, void *out64) {
    XCHG_REQUIRE(x && out64, "b2r_xchg_ipc_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    B2R_CUDA(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h;
    B2R_CUDA(cudaIpcGetMemHandle(&h, x->local));
    std::memcpy(out64, &h, 64);
    return B2R_OK;
}

extern "C" int b2r_xchg_open(b2r_xchg_handle x, const void *handles) {
    XCHG_REQUIRE(x && handles, "b2r_xchg_open: NULL argument");
    std::lock_guard<std::mutex> g(x->mu);
    B2R_CUDA(cudaSetDevice(x->device));
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank || x->box[r]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char *)handles + (size_t)r * 64, 64);
        void *p = nullptr;
        B2R_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->box[r] = (char *)p;
    }
    x->opened = true;
    return B2R_OK;
}

extern "C" int b2r_xchg_push(b2r_xchg_handle x, const int64_t *rows, const double *dist64, const int32_t *count, int nq, int k,
                             void *stream) {
    XCHG_REQUIRE(x && rows && dist64 && count, "b2r_xchg_push: NULL argument");
    XCHG_REQUIRE(x->opened, "b2r_xchg_push: the peers' mailboxes have not been opened (b2r_xchg_open)");
    XCHG_REQUIRE(nq >= 1 && nq <= x->nq_max && k >= 1 && k <= x->k_max, "b2r_xchg_push: nq / k exceed what the exchange was created for");
    std::lock_guard<std::mutex> g(x->mu);
    B2R_CUDA(cudaSetDevice(x->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if ((rc = flush_pending(x, s)) != B2R_OK) return rc;
    XchgDev p;
    fill(x, p, nq, k, ++x->seq_push);
    x->fused[p.slot] = false; x->nq_of[p.slot] = nq; x->k_of[p.slot] = k;
    p.rows = (const long long *)rows; p.d64 = dist64; p.cnt = count;
    if (p.seq > XCHG_SLOTS)   // the slot is free once every rank -- this one included: its merge may run on another stream -- has
                              // read what the slot held XCHG_SLOTS calls ago
        for (int r = 0; r < x->world; ++r)
            if ((rc = stream_wait(x, (size_t)XCHG_SLOTS * x->world + (size_t)p.slot * x->world + r, p.seq - XCHG_SLOTS, s)) != B2R_OK) return rc;
    const long long items = (long long)nq * (k + 1);
    const int grid = (int)std::max<long long>(1, std::min<long long>((items + XCHG_THREADS - 1) / XCHG_THREADS, 32));
    B2R_CUDA(launch_pdl(xchg_push_kernel, dim3(grid), dim3(XCHG_THREADS), 0, s, p));
    return B2R_OK;
}

extern "C" int b2r_xchg_merge(b2r_xchg_handle x, int nq, int k, int64_t *out_rows, float *out_dist, int32_t *out_count, void *stream) {
    XCHG_REQUIRE(x && out_rows && out_dist && out_count, "b2r_xchg_merge: NULL argument");
    XCHG_REQUIRE(x->opened, "b2r_xchg_merge: the peers' mailboxes have not been opened (b2r_xchg_open)");
    XCHG_REQUIRE(nq >= 1 && nq <= x->nq_max && k >= 1 && k <= x->k_max, "b2r_xchg_merge: nq / k exceed what the exchange was created for");
    std::lock_guard<std::mutex> g(x->mu);
    XCHG_REQUIRE(x->seq_merge < x->seq_push, "b2r_xchg_merge: no pushed batch is waiting to be merged");
    XCHG_REQUIRE(x->seq_push - x->seq_merge <= XCHG_SLOTS, "b2r_xchg_merge: more batches pushed than the mailbox has slots");
    B2R_CUDA(cudaSetDevice(x->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if ((rc = flush_pending(x, s)) != B2R_OK) return rc;     // (the arrival of the last fused batch: this merge may be waiting for it)
    XchgDev p;
    {
        const int slot = (int)((x->seq_merge + 1) % XCHG_SLOTS);
        XCHG_REQUIRE(x->nq_of[slot] == nq && x->k_of[slot] == k, "b2r_xchg_merge: nq / k differ from what the oldest unmerged batch was pushed with");
    }
    fill(x, p, nq, k, ++x->seq_merge);
    p.out_rows = (long long *)out_rows; p.out_dist = out_dist; p.out_cnt = out_count;
    p.spin = x->fused[p.slot] ? 1 : 0;
    if (!p.spin)
        for (int r = 0; r < x->world; ++r)      // own arrival word too: orders this merge after this rank's push when the streams differ
            if ((rc = stream_wait(x, (size_t)p.slot * x->world + r, p.seq, s)) != B2R_OK) return rc;
    const size_t smem = (size_t)x->world * k * 16;
    if (smem > 48 * 1024) B2R_CUDA(cudaFuncSetAttribute(xchg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // (programmatic launch: the merge's CTAs are placed while the kernel before it on the stream drains)
    B2R_CUDA(launch_pdl(xchg_merge_kernel, dim3(nq), dim3(XCHG_THREADS), smem, s, p));
    return B2R_OK;
}

// One fused call (b2r_query_push), everything decided under one lock so that a refused call changes nothing: the flag words the
// previous fused call left to publish, the rider (the oldest unmerged batch, merged inside this call's last kernel when its lists
// fit `rider_smem_limit` bytes of shared memory, by the stand-alone merge kernel behind the call otherwise), the next slot for this
// call's lists, and the acks its first kernel must see.
int b2r::xchg_begin_fused(b2r_xchg *x, int device, int nq, int k, FusedCall *c, int64_t *merge_rows, float *merge_dist,
                          int32_t *merge_count, size_t rider_smem_limit) {
    XCHG_REQUIRE(x->opened, "b2r_query_push: the peers' mailboxes have not been opened (b2r_xchg_open)");
    XCHG_REQUIRE(x->device == device, "b2r_query_push: the exchange lives on another device than the shard");
    XCHG_REQUIRE(nq >= 1 && nq <= x->nq_max && k >= 1 && k <= x->k_max, "b2r_query_push: nq / k exceed what the exchange was created for");
    std::lock_guard<std::mutex> g(x->mu);
    const unsigned unmerged = x->seq_push - x->seq_merge;
    XCHG_REQUIRE(!merge_rows || unmerged >= 1, "b2r_query_push: no earlier batch is waiting to be merged");
    XCHG_REQUIRE(unmerged < (unsigned)XCHG_SLOTS, "b2r_query_push: every mailbox slot holds a batch that has not been merged (b2r_xchg_merge)");
    c->flags = x->pending;                       // written by this call's first kernel
    std::memset(&x->pending, 0, sizeof x->pending);
    for (int r = 0; r < XCHG_MAX_WORLD; ++r) x->pending.box[r] = x->box[r];
    x->pending.world = x->world;
    c->rider.nq = 0; c->merge_after = false; c->rider_smem = 0;
    if (merge_rows) {
        const int slot = (int)((x->seq_merge + 1) % XCHG_SLOTS);
        fill(x, c->rider, x->nq_of[slot], x->k_of[slot], ++x->seq_merge);
        c->rider.spin = 1;
        c->rider.out_rows = (long long *)merge_rows; c->rider.out_dist = merge_dist; c->rider.out_cnt = merge_count;
        c->rider_smem = (size_t)x->world * c->rider.k * 16;
        if (c->rider_smem <= rider_smem_limit) {
            x->pending.ack_seq = c->rider.seq;
            x->pending.ack_off = x->flags_off + sizeof(unsigned) * ((size_t)XCHG_SLOTS * x->world + (size_t)slot * x->world + x->rank);
        } else {
            c->merge_after = true;                   // its own kernel, which also hands the slot back
        }
    }
    const unsigned seq = ++x->seq_push;
    const int slot = (int)(seq % XCHG_SLOTS);
    x->fused[slot] = true; x->nq_of[slot] = nq; x->k_of[slot] = k;
    x->pending.arrive_seq = seq;
    x->pending.arrive_off = x->flags_off + sizeof(unsigned) * ((size_t)slot * x->world + x->rank);
    for (int r = 0; r < XCHG_MAX_WORLD; ++r) c->push.box[r] = x->box[r];
    c->push.lists_off = (size_t)slot * x->slot_bytes + (size_t)x->rank * x->nq_max * x->entry_bytes;
    c->push.entry_bytes = (unsigned)x->entry_bytes; c->push.k_max = (unsigned)x->k_max;
    c->push.world = x->world;
    // before the slot is written again every rank must have merged what it held XCHG_SLOTS calls ago (ack words, local memory)
    c->wait_words = reinterpret_cast<const unsigned *>(x->local + x->flags_off) + (size_t)XCHG_SLOTS * x->world + (size_t)slot * x->world;
    c->wait_n = seq > (unsigned)XCHG_SLOTS ? x->world : 0;
    c->wait_val = seq - XCHG_SLOTS;
    return B2R_OK;
}

// a rider whose lists did not fit the exact scan's shared memory: the stand-alone merge kernel (it hands the slot back itself)
int b2r::xchg_launch_merge(b2r_xchg *x, XchgDev job, cudaStream_t s) {
    B2R_CUDA(cudaSetDevice(x->device));
    job.ticket = XCHG_TICKET_MERGE;
    const size_t smem = (size_t)job.world * job.k * 16;
    if (smem > 48 * 1024) B2R_CUDA(cudaFuncSetAttribute(xchg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2R_CUDA(launch_pdl(xchg_merge_kernel, dim3(job.nq), dim3(XCHG_THREADS), smem, s, job));
    return B2R_OK;
}

extern "C" int b2r_xchg_destroy(b2r_xchg_handle x) {
    if (!x) return B2R_OK;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank && x->box[r]) cudaIpcCloseMemHandle(x->box[r]);
    cudaFree(x->local);
    delete x;
    return B2R_OK;
}
