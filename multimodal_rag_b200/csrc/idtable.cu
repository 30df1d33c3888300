// Host-side id table of a collection: string id -> row of the device corpus, and row -> id bytes.
//
// It replaces, for the hot path, the id index Chroma keeps in sqlite (`embeddings.embedding_id`, looked up by
// collection.add / upsert / get(ids) / delete(ids) -- the reference's call sites are app/utils/embedder.py:518, 632, 640,
// 888) and the Python dict the first host mirror used: at 10M ids every dict probe and every dict insert of an 8192-row
// upsert is a cache miss (~0.6 us each, ~12 ms per batch), several times the device's share of the step.  Here a batch is
// hashed first, then probed with the home slots PREFETCHED a fixed distance ahead: the misses overlap instead of queueing.
//
// Layout: open addressing, linear probing, 16-byte slots {64-bit hash, row + 1}; a hit is confirmed against the id bytes
// of that row (an arena holding the bytes of every appended row back to back), so two ids with one hash are two entries.
// Load factor <= 0.6 (doubling rehash).  Deletion shifts the following run back (no tombstones).  No device work.
#include "engine.h"

#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <mutex>
#include <new>
#include <vector>

namespace {

struct Slot {
    uint64_t h;
    int64_t v;      // row + 1; 0 = empty
};

inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

// 64-bit mix of the id bytes (multiply-rotate over 8-byte words, avalanche at the end): process-independent
// `safe_end`: one past the last byte that may be READ (the batch buffer's end); a tail shorter than 8 bytes is fetched with
// one unaligned 8-byte load and masked when that stays inside the buffer
inline uint64_t hash_bytes(const char *p, size_t n, const char *safe_end = nullptr) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xFF51AFD7ED558CCDull);
    while (n >= 8) {
        uint64_t k;
        memcpy(&k, p, 8);
        k *= 0x87C37B91114253D5ull; k = rotl64(k, 31); k *= 0x4CF5AD432745937Full;
        h ^= k; h = rotl64(h, 27) * 5 + 0x52DCE729ull;
        p += 8; n -= 8;
    }
    if (n) {
        uint64_t k = 0;
        if (p + 8 <= safe_end) {
            memcpy(&k, p, 8);
            k &= ~0ull >> (64 - 8 * n);         // little-endian: the first n bytes
        } else {
            memcpy(&k, p, n);
        }
        k *= 0x87C37B91114253D5ull; k = rotl64(k, 31); k *= 0x4CF5AD432745937Full;
        h ^= k;
    }
    h ^= h >> 33; h *= 0xFF51AFD7ED558CCDull; h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ull; h ^= h >> 33;
    return h;
}

// The slot array is the one structure every probe misses the cache on; at 10M ids it is 512 MB, so with 4 KB pages every
// probe is also a TLB miss and a page walk.  Anonymous mapping (zero pages on first touch, like calloc) with transparent
// huge pages asked for: one TLB entry covers 2 MB of slots.
Slot *map_slots(uint64_t n_slots) {
    const size_t bytes = (size_t)n_slots * sizeof(Slot);
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    if (bytes >= (2u << 20)) madvise(p, bytes, MADV_HUGEPAGE);      // advisory: ignored where THP is off
    return static_cast<Slot *>(p);
}
void unmap_slots(Slot *p, uint64_t n_slots) {
    if (p) munmap(p, (size_t)n_slots * sizeof(Slot));
}

}  // namespace

struct b2r_idtab {
    Slot *slots = nullptr;
    uint64_t mask = 0;                  // slot count - 1
    int64_t live = 0;
    std::vector<char> arena;            // id bytes of every appended row, one separator byte after each
    std::vector<int64_t> row_off{0};    // row r's id = arena[row_off[r] .. row_off[r + 1] - 1)
    std::vector<uint64_t> scratch;      // hashes of the batch in flight
    std::vector<int64_t> dup_scratch;   // batch-local table of b2r_idtab_lookup's duplicate check
    std::mutex mu;

    int64_t rows() const { return (int64_t)row_off.size() - 1; }
    bool same_id(int64_t row, const char *p, int64_t len) const {
        const int64_t b = row_off[row];
        return row_off[row + 1] - b - 1 == len && memcmp(arena.data() + b, p, (size_t)len) == 0;
    }
    // slot holding this id, or the empty slot where it belongs
    uint64_t find(uint64_t h, const char *p, int64_t len) const {
        uint64_t i = h & mask;
        while (slots[i].v != 0 && !(slots[i].h == h && same_id(slots[i].v - 1, p, len))) i = (i + 1) & mask;
        return i;
    }
    bool resize(uint64_t n_slots) {
        Slot *fresh = map_slots(n_slots);
        if (!fresh) return false;
        const uint64_t m = n_slots - 1;
        if (slots)
            for (uint64_t i = 0; i <= mask; ++i)
                if (slots[i].v != 0) {
                    uint64_t j = slots[i].h & m;
                    while (fresh[j].v != 0) j = (j + 1) & m;
                    fresh[j] = slots[i];
                }
        unmap_slots(slots, mask + 1);
        slots = fresh;
        mask = m;
        return true;
    }
    bool room_for(int64_t more) {
        uint64_t want = mask + 1;
        while ((double)(live + more) > 0.6 * (double)want) want *= 2;
        return want == mask + 1 || resize(want);
    }
    void hash_batch(const char *bytes, const int64_t *off, int64_t n, int64_t gap) {
        scratch.resize((size_t)n);
        uint64_t *hs = scratch.data();
        const char *end = n ? bytes + off[n] - gap : bytes;
        for (int64_t i = 0; i < n; ++i) hs[i] = hash_bytes(bytes + off[i], (size_t)(off[i + 1] - off[i] - gap), end);
    }
    // The probe loops run PREFETCH_AHEAD ids behind a prefetch of the home slot: the cache (and TLB) misses of a batch
    // overlap instead of queueing.  Measured on a 10M-id table, 8192-id batch: 0.18 ms against 0.24 ms unprefetched and
    // 0.31 ms when the whole batch is prefetched up front (that floods the fill buffers).
    // A HIT costs two more dependent misses -- the row's arena offset, then its id bytes, for the byte comparison that confirms
    // it -- so the same window looks twice more at the ids it is about to probe: 16 ahead the (by then cached) home slot is
    // read and, when its hash matches, the row's offset is prefetched; 8 ahead the offset is read and the id bytes are
    // prefetched.  (Hints only: an entry displaced from its home slot simply misses them.)  8192-id upsert batch with 10 % known
    // ids on a 10M-id table (build box, medians of 56 batches): lookup 0.88 -> 0.54-0.71 ms, append 0.65 -> 0.38-0.47 ms.
    static constexpr int64_t PREFETCH_AHEAD = 24, PEEK_OFFSET = 16, PEEK_BYTES = 8;
    void prefetch_home(int64_t i, int64_t n) const {
        if (i + PREFETCH_AHEAD < n) __builtin_prefetch(&slots[scratch[(size_t)(i + PREFETCH_AHEAD)] & mask]);
        if (i + PEEK_OFFSET < n) {
            const uint64_t h = scratch[(size_t)(i + PEEK_OFFSET)];
            const Slot &sl = slots[h & mask];
            if (sl.v != 0 && sl.h == h) __builtin_prefetch(&row_off[(size_t)(sl.v - 1)]);
        }
        if (i + PEEK_BYTES < n) {
            const uint64_t h = scratch[(size_t)(i + PEEK_BYTES)];
            const Slot &sl = slots[h & mask];
            if (sl.v != 0 && sl.h == h) __builtin_prefetch(arena.data() + row_off[(size_t)(sl.v - 1)]);
        }
    }
    void erase_slot(uint64_t i) {
        uint64_t j = i;
        for (;;) {
            j = (j + 1) & mask;
            if (slots[j].v == 0) break;
            const uint64_t k = slots[j].h & mask;       // home of the entry at j: it may move to i unless i < home <= j (cyclically)
            if (i <= j ? (i < k && k <= j) : (i < k || k <= j)) continue;
            slots[i] = slots[j];
            i = j;
        }
        slots[i].v = 0;
        --live;
    }
};

#define IDT_REQUIRE(cond, msg)                          \
    do {                                                \
        if (!(cond)) { b2r::set_error(msg); return B2R_EINVAL; } \
    } while (0)

// id i of a batch = bytes[off[i] .. off[i + 1] - gap): `gap` separator bytes follow every id (0 = packed back to back,
// 1 = the NUL-joined form a Python binding gets from "\0".join(ids) in one call)
static bool offsets_ok(const int64_t *off, int64_t n, int64_t gap) {
    if (gap < 0 || gap > 1) return false;
    for (int64_t i = 0; i < n; ++i)
        if (off[i + 1] - gap < off[i]) return false;
    return n == 0 || off[0] >= 0;
}

extern "C" {

int b2r_idtab_create(int64_t reserve_ids, b2r_idtab_handle *out) {
    IDT_REQUIRE(out && reserve_ids >= 0, "b2r_idtab_create: bad argument");
    b2r_idtab *t = new (std::nothrow) b2r_idtab();
    uint64_t n_slots = 1024;
    while (0.6 * (double)n_slots < (double)reserve_ids) n_slots *= 2;
    if (!t || !t->resize(n_slots)) {
        delete t;
        b2r::set_error("b2r_idtab_create: host allocation failed");
        return B2R_ENOMEM;
    }
    *out = t;
    return B2R_OK;
}

int b2r_idtab_destroy(b2r_idtab_handle t) {
    if (t) {
        unmap_slots(t->slots, t->mask + 1);
        delete t;
    }
    return B2R_OK;
}

int b2r_idtab_clear(b2r_idtab_handle t) {
    IDT_REQUIRE(t, "b2r_idtab_clear: null table");
    std::lock_guard<std::mutex> g(t->mu);
    memset(t->slots, 0, (size_t)(t->mask + 1) * sizeof(Slot));
    t->live = 0;
    t->arena.clear();
    t->row_off.assign(1, 0);
    return B2R_OK;
}

int64_t b2r_idtab_live(b2r_idtab_handle t) { return t ? t->live : -1; }
int64_t b2r_idtab_rows(b2r_idtab_handle t) { return t ? t->rows() : -1; }

int b2r_idtab_lookup(b2r_idtab_handle t, const char *bytes, const int64_t *offsets, int64_t n, int64_t gap,
                     int64_t *rows_out, int64_t *first_dup) {
    IDT_REQUIRE(t && n >= 0 && (n == 0 || (bytes && offsets && rows_out)), "b2r_idtab_lookup: bad argument");
    IDT_REQUIRE(offsets_ok(offsets, n, gap), "b2r_idtab_lookup: offsets must not decrease");
    std::lock_guard<std::mutex> g(t->mu);
    try {
        t->hash_batch(bytes, offsets, n, gap);
        if (first_dup) {
            // ids repeated inside the batch: a throw-away table of batch indices (2n..4n slots, cache resident for an
            // upsert-sized batch), same hash + byte comparison as the big one
            *first_dup = -1;
            uint64_t m = 16;
            while (m < (uint64_t)n * 2) m *= 2;
            t->dup_scratch.assign((size_t)m, -1);
            for (int64_t i = 0; i < n && *first_dup < 0; ++i) {
                const uint64_t h = t->scratch[(size_t)i];
                const char *p = bytes + offsets[i];
                const int64_t len = offsets[i + 1] - offsets[i] - gap;
                uint64_t s = h & (m - 1);
                for (;; s = (s + 1) & (m - 1)) {
                    const int64_t j = t->dup_scratch[(size_t)s];
                    if (j < 0) { t->dup_scratch[(size_t)s] = i; break; }
                    if (t->scratch[(size_t)j] == h && offsets[j + 1] - offsets[j] - gap == len &&
                        memcmp(bytes + offsets[j], p, (size_t)len) == 0) { *first_dup = i; break; }
                }
            }
        }
    } catch (const std::bad_alloc &) {
        b2r::set_error("b2r_idtab_lookup: host allocation failed");
        return B2R_ENOMEM;
    }
    for (int64_t i = 0; i < n; ++i) {
        t->prefetch_home(i, n);
        const uint64_t s = t->find(t->scratch[(size_t)i], bytes + offsets[i], offsets[i + 1] - offsets[i] - gap);
        rows_out[i] = t->slots[s].v - 1;        // empty slot: -1
    }
    return B2R_OK;
}

int b2r_idtab_append(b2r_idtab_handle t, const char *bytes, const int64_t *offsets, int64_t n, int64_t gap,
                     int64_t first_row, int64_t *prev_out) {
    IDT_REQUIRE(t && n >= 0 && (n == 0 || (bytes && offsets)), "b2r_idtab_append: bad argument");
    IDT_REQUIRE(offsets_ok(offsets, n, gap), "b2r_idtab_append: offsets must not decrease");
    std::lock_guard<std::mutex> g(t->mu);
    IDT_REQUIRE(first_row == t->rows(), "b2r_idtab_append: first_row must be the number of rows appended so far");
    if (n == 0) return B2R_OK;
    try {
        if (!t->room_for(n)) throw std::bad_alloc();
        const size_t need_bytes = t->arena.size() + (size_t)(offsets[n] - offsets[0] + n), need_rows = t->row_off.size() + (size_t)n;
        if (need_bytes > t->arena.capacity()) t->arena.reserve(need_bytes * 2);
        if (need_rows > t->row_off.capacity()) t->row_off.reserve(need_rows * 2);
        t->scratch.reserve((size_t)n);
    } catch (const std::bad_alloc &) {
        b2r::set_error("b2r_idtab_append: host allocation failed");
        return B2R_ENOMEM;
    }
    // the ids become rows first (find() confirms a hit against the arena), then every id is pointed at its new row
    if (gap == 1) {                             // separated already: one copy for the batch
        const int64_t base = (int64_t)t->arena.size() - offsets[0];
        t->arena.insert(t->arena.end(), bytes + offsets[0], bytes + offsets[n]);
        for (int64_t i = 1; i <= n; ++i) t->row_off.push_back(base + offsets[i]);
    } else {
        for (int64_t i = 0; i < n; ++i) {
            t->arena.insert(t->arena.end(), bytes + offsets[i], bytes + offsets[i + 1]);
            t->arena.push_back('\0');
            t->row_off.push_back((int64_t)t->arena.size());
        }
    }
    t->hash_batch(bytes, offsets, n, gap);
    for (int64_t i = 0; i < n; ++i) {
        t->prefetch_home(i, n);
        const uint64_t h = t->scratch[(size_t)i];
        const uint64_t s = t->find(h, bytes + offsets[i], offsets[i + 1] - offsets[i] - gap);
        if (prev_out) prev_out[i] = t->slots[s].v - 1;
        if (t->slots[s].v == 0) {
            t->slots[s].h = h;
            ++t->live;
        }
        t->slots[s].v = first_row + i + 1;      // an id repeated inside the batch ends at its last row
    }
    return B2R_OK;
}

int b2r_idtab_erase_rows(b2r_idtab_handle t, const int64_t *rows, int64_t n) {
    IDT_REQUIRE(t && n >= 0 && (n == 0 || rows), "b2r_idtab_erase_rows: bad argument");
    std::lock_guard<std::mutex> g(t->mu);
    for (int64_t i = 0; i < n; ++i) IDT_REQUIRE(rows[i] >= 0 && rows[i] < t->rows(), "b2r_idtab_erase_rows: row out of range");
    for (int64_t i = 0; i < n; ++i) {
        const int64_t r = rows[i], b = t->row_off[r], len = t->row_off[r + 1] - b - 1;
        const uint64_t h = hash_bytes(t->arena.data() + b, (size_t)len);
        uint64_t s = h & t->mask;
        while (t->slots[s].v != 0 && t->slots[s].v != r + 1) s = (s + 1) & t->mask;     // the id may point at a newer row by now:
        if (t->slots[s].v == r + 1) t->erase_slot(s);                                   // then nothing is unmapped
    }
    return B2R_OK;
}

int b2r_idtab_ids_of(b2r_idtab_handle t, const int64_t *rows, int64_t n, char *out, int64_t out_cap, int64_t *offsets_out,
                     int64_t *need) {
    IDT_REQUIRE(t && n >= 0 && need && (n == 0 || (rows && offsets_out)), "b2r_idtab_ids_of: bad argument");
    std::lock_guard<std::mutex> g(t->mu);
    int64_t total = 0;
    for (int64_t i = 0; i < n; ++i) {
        IDT_REQUIRE(rows[i] >= 0 && rows[i] < t->rows(), "b2r_idtab_ids_of: row out of range");
        total += t->row_off[rows[i] + 1] - t->row_off[rows[i]];
    }
    *need = total;
    if (total > out_cap || n == 0) return B2R_OK;       // the caller retries with `need` bytes
    IDT_REQUIRE(out, "b2r_idtab_ids_of: null output");
    int64_t w = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t b = t->row_off[rows[i]], len = t->row_off[rows[i] + 1] - b - 1;
        offsets_out[i] = w;
        memcpy(out + w, t->arena.data() + b, (size_t)len);
        w += len;
        out[w++] = '\0';
    }
    offsets_out[n] = w;
    return B2R_OK;
}

}  // extern "C"
