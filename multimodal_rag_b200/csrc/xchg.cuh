// Device side of the cross-shard exchange (xchg.cu describes the protocol): the mailbox layout, the merge of one query's
// world x k gathered candidates, the arrival wait and the "slot has been read" hand-back -- shared by the stand-alone merge kernel
// and by the exact-scan kernel, which carries the merge of the PREVIOUS fused batch as a rider (b2r_query_push).
#pragma once
#include "common.cuh"

namespace b2r {

constexpr int XCHG_MAX_WORLD = 8;
constexpr int XCHG_THREADS = 256;
constexpr int XCHG_SLOTS = 4;          // mailbox slots, used round-robin by sequence number
// flag block of a mailbox (u32 words): arrival [SLOTS][world], ack [SLOTS][world], then the exit tickets of the push and the
// merge kernel (the fused form needs none: its flag words are written by the next kernel on the stream)
constexpr int XCHG_TICKET_PUSH = 0, XCHG_TICKET_MERGE = 1;

struct XchgDev {
    int rank, world, nq, k, slot;
    int spin;                      // merge: wait for the arrival words inside the kernel (fused pushes) instead of on the stream
    int ticket;                    // exit ticket of the stand-alone merge kernel (a rider takes none)
    unsigned seq;
    int nq_max, k_max;
    size_t entry_bytes;            // one query's list in a mailbox: rows[k_max] i64 | dist[k_max] f64 | count i32 (+pad)
    size_t slot_bytes;             // world * nq_max * entry_bytes
    size_t flags_off;              // the flag block
    char *box[XCHG_MAX_WORLD];     // every rank's mailbox as mapped in this process (box[rank] = the local allocation)
    const long long *rows; const double *d64; const int *cnt;      // push: this rank's local results
    long long *out_rows; float *out_dist; int *out_cnt;            // merge: outputs
};

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned *flag_words(const XchgDev &p, int r) { return reinterpret_cast<unsigned *>(p.box[r] + p.flags_off); }

// Flag words a fused call leaves for the NEXT kernel on its stream to write (the query preparation of the following call, or a
// one-warp flush kernel in front of any other exchange call): the arrival of the batch it pushed -- a release store at system
// scope, 1.75 us (scripts/dev/sys_scope_ubench.cu), which there overlaps that kernel's own work instead of ending a kernel -- and the
// "slot has been read" of the batch its rider merged.  By then the grids that stored the lists / read the slot have completed.
struct XchgFlags {
    char *box[XCHG_MAX_WORLD];
    unsigned long long arrive_off;      // word (slot, this rank) of the arrival block; written in every mailbox
    unsigned long long ack_off;         // word (slot, this rank) of the ack block; written in every mailbox
    unsigned arrive_seq, ack_seq;       // 0 = nothing to publish
    int world;
};
// threads [first, first + 2 * world) of one CTA
__device__ __forceinline__ void xchg_publish_flags(const XchgFlags &f, int first) {
    const int t = (int)threadIdx.x - first;
    if (t < 0 || t >= 2 * f.world) return;
    if (t < f.world) {
        if (f.arrive_seq) st_release_sys(reinterpret_cast<unsigned *>(f.box[t] + f.arrive_off), f.arrive_seq);
    } else if (f.ack_seq) {
        // the mailbox loads of the merging grid returned before that grid ended: a plain store is enough
        *reinterpret_cast<volatile unsigned *>(f.box[t - f.world] + f.ack_off) = f.ack_seq;
    }
}

// All threads of the CTA: the lists of every rank (this one included) of batch p.seq have arrived.  Bounded: a peer that never
// pushed fails the launch instead of hanging the GPU.
__device__ __forceinline__ void xchg_wait_arrivals(const XchgDev &p) {
    if ((int)threadIdx.x < p.world) {
        const unsigned *w = flag_words(p, p.rank) + p.slot * p.world + threadIdx.x;
        const unsigned long long t0 = globaltimer_ns();
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(w) : "memory");
            if ((int)(v - p.seq) >= 0) break;
            if (globaltimer_ns() - t0 > 4000000000ull) __trap();
            __nanosleep(100);
        }
    }
    __syncthreads();
}

// One query, XCHG_THREADS threads: rank every gathered candidate by counting (world * k <= 8192); sm = world * k * 16 bytes
__device__ __forceinline__ void xchg_merge_query(const XchgDev &p, int q, unsigned char *sm, int *s_valid) {
    const int total = p.world * p.k;
    double *sd = reinterpret_cast<double *>(sm);
    long long *sr = reinterpret_cast<long long *>(sd + total);
    const int tid = threadIdx.x;
    if (tid == 0) *s_valid = 0;
    __syncthreads();
    const char *mine = p.box[p.rank] + (size_t)p.slot * p.slot_bytes;
    int my_valid = 0;
    for (int c = tid; c < total; c += XCHG_THREADS) {      // written by the peers: L2 loads
        const int sh = c / p.k, i = c % p.k;
        const char *e = mine + ((size_t)sh * p.nq_max + q) * p.entry_bytes;
        const bool ok = i < __ldcg(reinterpret_cast<const int *>(e + (size_t)p.k_max * 16));
        sd[c] = ok ? __ldcg(reinterpret_cast<const double *>(e + (size_t)p.k_max * 8) + i) : __longlong_as_double(0x7ff0000000000000ll);
        sr[c] = ok ? __ldcg(reinterpret_cast<const long long *>(e) + i) : -1;
        my_valid += ok ? 1 : 0;
    }
    if (my_valid) atomicAdd(s_valid, my_valid);
    __syncthreads();
    const int n_out = min(*s_valid, p.k);
    for (int c = tid; c < total; c += XCHG_THREADS) {
        const long long r = sr[c];
        if (r < 0) continue;
        const double d = sd[c];
        int rank = 0;
        for (int j = 0; j < total; ++j) {
            const long long rj = sr[j];
            const double dj = sd[j];
            rank += (rj >= 0 && (dj < d || (dj == d && rj < r))) ? 1 : 0;
        }
        if (rank < p.k) {
            p.out_rows[(size_t)q * p.k + rank] = r;
            p.out_dist[(size_t)q * p.k + rank] = (float)d;
        }
    }
    for (int t = n_out + tid; t < p.k; t += XCHG_THREADS) {
        p.out_rows[(size_t)q * p.k + t] = -1;
        p.out_dist[(size_t)q * p.k + t] = __int_as_float(0x7f800000);
    }
    if (tid == 0) p.out_cnt[q] = n_out;
}

// Every CTA of the merging grid on its way out: the last one tells every peer that this rank has read the slot.  The mailbox
// loads of this grid have RETURNED (their values were used) before the tickets are taken, so nothing a peer writes after it sees
// the word can reach them: plain stores, no system-scope release (1.75 us each on a B200; scripts/dev/sys_scope_ubench.cu).
__device__ __forceinline__ void xchg_publish_ack(const XchgDev &p, unsigned *s_ticket) {
    unsigned *ticket = flag_words(p, p.rank) + 2 * XCHG_SLOTS * p.world + p.ticket;
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); *s_ticket = atomicAdd(ticket, 1u); }
    __syncthreads();
    if (*s_ticket == gridDim.x - 1) {
        if (threadIdx.x == 0) *ticket = 0u;
        if ((int)threadIdx.x < p.world)
            *reinterpret_cast<volatile unsigned *>(flag_words(p, threadIdx.x) + XCHG_SLOTS * p.world + p.slot * p.world + p.rank) = p.seq;
    }
}

}  // namespace b2r
