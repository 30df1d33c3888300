/* b2r -- B200 exact vector-retrieval engine: the C ABI (drop-in boundary).
 *
 * This is what a binding on the reference side would call instead of the
 * chromadb.Collection methods that app/utils/embedder.py reaches through
 * `self.collection` (reference file:line cited per entry point below).  Plain C:
 * pointers and sizes only, no torch / C++ types.  All entry points are thread-safe
 * (one host mutex per handle; every call does cudaSetDevice, because the reference
 * calls the store from asyncio.to_thread worker threads, embedder.py:517,595).
 *
 * Conventions
 *   - every function returns a b2r_status (0 = ok); b2r_last_error() gives the
 *     message for the calling thread's last failure.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - data pointers marked "host or device" are classified with
 *     cudaPointerGetAttributes.  With device pointers a call only enqueues work on
 *     `stream`; with host pointers it copies (cudaMemcpyAsync on `stream`) and, for
 *     outputs, synchronises the stream before returning.
 *   - row numbers are dense insertion indices starting at 0 (the host keeps
 *     row <-> Chroma id).  Rows are never reused; delete/upsert tombstone.
 *   - there is NO CPU fallback: without a CUDA device every call fails.
 */
#ifndef B2R_H_
#define B2R_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2R_ABI_VERSION 4

typedef struct b2r_index *b2r_handle;

typedef enum {
    B2R_OK = 0,
    B2R_EINVAL = 1,      /* bad argument / dimension mismatch -> ValueError upstream */
    B2R_ECUDA = 2,       /* CUDA runtime failure              -> RuntimeError        */
    B2R_ENOMEM = 3,      /* device allocation failed                                */
    B2R_EUNSUPPORTED = 4 /* k or dim outside what the kernels are built for         */
} b2r_status;

/* Chroma `hnsw:space` (collection metadata; default l2 -- embedder.py:179-182 passes
 * no space, the committed chroma_db/ says cosine).  Distances returned:
 *   L2     sum (q-x)^2            COSINE  1 - q^.x^  (both normalised at ingest/query)
 *   IP     1 - q.x                                                                   */
typedef enum { B2R_SPACE_L2 = 0, B2R_SPACE_COSINE = 1, B2R_SPACE_IP = 2 } b2r_space;

/* create flags */
#define B2R_FLAG_NO_F32_MASTER 1u /* keep only the packed bf16 rows: the corpus *is* its
                                     bf16 rounding, re-rank and get_rows read that.     */

/* Row filter (Chroma `where`, embedder.py:543,599).  The host evaluates the clause on
 * its metadata tables and hands the device either a type-code mask (fast path for
 * {"type": ...}) or a general allow bitmap.  Tombstoned rows never match.             */
#define B2R_MAX_COLUMNS 16      /* dictionary-encoded metadata columns kept on the device */
#define B2R_WHERE_MAX_NODES 32  /* nodes of one compiled where clause                      */

/* A compiled `where` clause, evaluated on the device against dictionary-encoded metadata columns
 * (b2r_column_set): the clause tree in postfix order.  A leaf tests one column: the row's code c
 * (-1 = key absent -> false) passes iff bit c of the leaf's look-up table is set; the host builds
 * that table by applying the comparison ($eq/$ne/$gt/$gte/$lt/$lte/$in/$nin) to the column's
 * DISTINCT values, so the device never compares strings or numbers.  AND / OR pop two, push one. */
typedef enum { B2R_WHERE_LEAF = 0, B2R_WHERE_AND = 1, B2R_WHERE_OR = 2 } b2r_where_op;
typedef struct {
    int32_t op;                 /* b2r_where_op                                          */
    int32_t column;             /* leaf: device column 0..B2R_MAX_COLUMNS-1              */
    uint32_t lut_offset;        /* leaf: first word of its table in `lut`                */
    uint32_t lut_values;        /* leaf: codes 0..lut_values-1 are covered by the table  */
} b2r_where_node;
typedef struct {
    int32_t n_nodes;            /* 1..B2R_WHERE_MAX_NODES, postfix                        */
    b2r_where_node nodes[B2R_WHERE_MAX_NODES];
    const uint32_t *lut;        /* all leaf tables, host or device                        */
    int64_t lut_words;
} b2r_where;

typedef struct {
    uint64_t type_mask;         /* bit c set = rows with type_code c pass; ~0 = all    */
    const uint32_t *allow_bits; /* NULL, or ceil(rows/32) words, host or device:
                                   bit (r&31) of word r>>5 set = row r passes          */
    const b2r_where *where;     /* NULL, or a compiled clause (host struct); a row must
                                   pass the type mask AND allow_bits AND the clause      */
} b2r_filter;

typedef struct {
    int32_t dim, dim_padded, space;
    uint32_t flags;
    int64_t rows;          /* rows appended so far (live + tombstoned)                */
    int64_t live;          /* rows not tombstoned                                     */
    int64_t capacity;      /* rows the current allocation holds                       */
    int64_t bytes_device;  /* HBM held by this handle                                 */
    int64_t n_queries;     /* queries answered                                        */
    int64_t n_exact_fallbacks; /* queries whose bf16 candidate certificate failed and
                                  were recomputed by the exact fp64 scan (only counted
                                  on calls with host outputs)                          */
    int32_t sm_count, device;
    int64_t n_pool_queries;  /* K3: queries finalized from a candidate pool ...                */
    int64_t n_pool_entries;  /* ... and the total number of pool entries they merged           */
} b2r_stats;

int b2r_abi_version(void);
const char *b2r_last_error(void);

/* replaces chromadb.Client(...).create_collection / get_collection
 * (app/utils/embedder.py:170-183).  capacity_rows is a reservation hint (grows x2). */
int b2r_create(int dim, int space, int64_t capacity_rows, int device, uint32_t flags,
               b2r_handle *out);
/* replaces client.delete_collection (app/utils/embedder.py:669-672) */
int b2r_destroy(b2r_handle h);
/* drop all rows, keep the allocation (delete_all_documents, embedder.py:658-688) */
int b2r_clear(b2r_handle h);
int b2r_reserve(b2r_handle h, int64_t capacity_rows);

/* replaces the vector half of collection.add / upsert (app/utils/embedder.py:517-523):
 * x is [n, dim] fp32 row-major, host or device.  Fused on device: (cosine) L2
 * normalise, bf16 pack, fp32 master store, per-row -|x|^2/2 for l2.  type_code is n
 * bytes (values 0..62), host or device, or NULL (= 0).  Rows become visible to every
 * later call on the same stream.  *first_row_out = row number given to x[0].        */
int b2r_ingest_f32(b2r_handle h, const float *x, int64_t n, const uint8_t *type_code,
                   int64_t *first_row_out, void *stream);

/* Metadata column `column` for rows [first_row, first_row + n): codes[i] = index of the row's value in the
 * host's dictionary of that key's distinct values, -1 = the row has no such key.  Rows never set read as -1.
 * This is the device half of Chroma's metadata segment for `where` (collection.query(where=...),
 * app/utils/embedder.py:599; collection.get(where={"doc_id": ...}), :632): the host keeps the dictionaries,
 * the device keeps one int32 per row per key and evaluates compiled clauses (b2r_where) in a bitmap kernel.  */
int b2r_column_set(b2r_handle h, int column, int64_t first_row, int64_t n, const int32_t *codes, void *stream);
/* Evaluate a filter exactly as b2r_query would and return the pass bitmap (ceil(rows/32) words, host or device):
 * live rows that pass the type mask, the allow bitmap and the clause.  Used by get(where=...) / delete(where=...)
 * and by the parity tests of the clause kernel.                                                              */
int b2r_filter_eval(b2r_handle h, const b2r_filter *filter, uint32_t *out_bits, void *stream);

/* replaces the vector half of collection.delete (app/utils/embedder.py:639-642) and the
 * overwrite half of upsert: rows (host array) stop matching any query.               */
int b2r_tombstone(b2r_handle h, const int64_t *rows, int64_t n, void *stream);

/* replaces collection.query(query_embeddings, n_results, where)
 * (app/utils/embedder.py:595-601, 900-905): q is [nq, dim] fp32 (host or device),
 * k = n_results.  Outputs, host or device (all three the same kind):
 *   out_rows  [nq, k] int64, ascending distance, ties -> lower row; padded with -1
 *   out_dist  [nq, k] fp32 distance in the collection's space; padded with +inf
 *   out_count [nq]    int32 = min(k, rows passing the filter)
 * Exact: identical rows to an fp64 brute force over the stored rows.                 */
int b2r_query(b2r_handle h, const float *q, int nq, int k, const b2r_filter *filter,
              int64_t *out_rows, float *out_dist, int32_t *out_count, void *stream);
/* same, plus the fp64 distances the ordering was decided on (out_dist64 [nq,k] or NULL);
 * the cross-shard merge needs them to stay exact.                                    */
int b2r_query_ex(b2r_handle h, const float *q, int nq, int k, const b2r_filter *filter,
                 int64_t *out_rows, float *out_dist, double *out_dist64, int32_t *out_count,
                 void *stream);

/* Pipelined form of b2r_query for HOST buffers: returns as soon as the work is enqueued -- the query batch travels on
 * an internal upload stream, the kernels run on `stream`, the results travel back on an internal download stream -- so that with two
 * calls in flight the transfers of one overlap the kernels of the other (the reference fires its queries concurrently,
 * app/utils/embedder.py:809-815; this is the batched equivalent).  q and the outputs are host arrays (page-locked
 * arrays are used in place, pageable ones go through pinned mirrors) and must stay valid until b2r_wait(ticket)
 * returns, which is also when the outputs are filled.  At most two tickets are outstanding per handle: a third call
 * first completes the oldest.  A filter, if given, must only use device-resident or type-mask forms.                */
int b2r_query_async(b2r_handle h, const float *q, int nq, int k, const b2r_filter *filter, int64_t *out_rows,
                    float *out_dist, int32_t *out_count, void *stream, uint64_t *ticket_out);
int b2r_wait(b2r_handle h, uint64_t ticket);

/* replaces collection.get(ids, include=['embeddings']) (app/utils/embedder.py:887-891):
 * the stored fp32 rows (normalised in cosine space, as hnswlib stores them).         */
int b2r_get_rows_f32(b2r_handle h, const int64_t *rows, int64_t n, float *out, void *stream);

/* replaces collection.count() (app/utils/embedder.py:700): live rows               */
int64_t b2r_count(b2r_handle h);
int b2r_get_stats(b2r_handle h, b2r_stats *out);

/* Persistence of the vector half of a collection -- what chromadb's persist_directory keeps in its HNSW segment
 * files (app/utils/embedder.py:164-170: ChromaSettings(persist_directory=settings.CHROMA_PERSIST_DIR); the committed
 * chroma_db/<segment>/{data_level0,header,length,link_lists}.bin).  b2r_save writes the shard exactly as it sits in
 * HBM (packed bf16 rows, fp32 master, l2 bias, type codes / tombstones, the measured rounding norms) to one file;
 * b2r_load rebuilds a handle from it without re-normalising or re-packing, so a loaded shard answers bit-identically.
 * The host keeps ids / documents / metadata next to it (multimodal_rag_b200/collection.py).
 * File: 128-byte header ("B2RS", version, dim, dim_padded, space, flags, rows, live, row_base, norms, payload checksum)
 * followed by the arrays in the order above.                                                                   */
int b2r_save(b2r_handle h, const char *path, void *stream);
int b2r_load(const char *path, int device, int64_t capacity_rows, b2r_handle *out);

/* Row-sharded corpus (one process per GPU): rows reported by b2r_query are
 * row_base + local row, so every shard can answer in global row numbers.            */
int b2r_set_row_base(b2r_handle h, int64_t row_base);

/* Cross-shard merge: after an allgather of every rank's local b2r_query_ex result, pick
 * the global top-k per query.  in_rows [nshards,nq,k] int64 (global rows, -1 = pad),
 * in_dist64 [nshards,nq,k] fp64, in_count [nshards,nq]; all device pointers.
 * Order: (fp64 distance, global row) -- the same total order a single shard uses.
 * No reference counterpart: the reference is single-process.                        */
int b2r_merge_shards(const int64_t *in_rows, const double *in_dist64, const int32_t *in_count,
                     int nshards, int nq, int k, int64_t *out_rows, float *out_dist,
                     int32_t *out_count, int device, void *stream);

/* Same merge over ONE gathered buffer: shard s's block starts at packed + s*shard_stride and holds rows
 * [nq,k] int64 at off_rows, fp64 distances [nq,k] at off_dist64, counts [nq] int32 at off_count (byte offsets,
 * 8-byte aligned) -- so a rank can expose rows/dist64/count as views of one allocation and exchange them
 * with a single all_gather.                                                                          */
int b2r_merge_shards_packed(const void *packed, int64_t shard_stride, int64_t off_rows, int64_t off_dist64,
                            int64_t off_count, int nshards, int nq, int k, int64_t *out_rows, float *out_dist,
                            int32_t *out_count, int device, void *stream);

/* The same exchange + merge WITHOUT a collective library on the data path.  Every rank owns a mailbox (device memory
 * allocated by this library, mapped into the other processes with CUDA IPC).  b2r_xchg_push (enqueue on the stream of the
 * scan) stores this rank's lists straight into every peer's mailbox over NVLink and publishes a sequence number there;
 * b2r_xchg_merge (same stream, or another one so that the exchange of batch i overlaps the scan of batch i+1) makes its stream
 * wait for the peers' sequence numbers with stream memory operations -- no SM spins while a peer is late -- and merges.  Four
 * mailbox slots rotate; a slot is rewritten only after every rank has read it.  Collective: every rank makes the same
 * sequence of push / merge calls (same nq, k).  Setup: b2r_xchg_create on every rank, exchange the 64-byte handles by any
 * means (torch.distributed all_gather_object), b2r_xchg_open with all of them in rank order, then a barrier.  world <= 8.
 * No reference counterpart: the reference is single-process.                                                              */
typedef struct b2r_xchg *b2r_xchg_handle;
int b2r_xchg_create(int device, int rank, int world, int nq_max, int k_max, b2r_xchg_handle *out);
int b2r_xchg_ipc_handle(b2r_xchg_handle x, void *out64);
int b2r_xchg_open(b2r_xchg_handle x, const void *handles /* world x 64 bytes, rank order */);
/* rows / dist64 / count: this rank's b2r_query_ex results (device) */
int b2r_xchg_push(b2r_xchg_handle x, const int64_t *rows, const double *dist64, const int32_t *count, int nq, int k, void *stream);
/* merges the oldest pushed batch that has not been merged yet; outputs as b2r_merge_shards */
int b2r_xchg_merge(b2r_xchg_handle x, int nq, int k, int64_t *out_rows, float *out_dist, int32_t *out_count, void *stream);
int b2r_xchg_destroy(b2r_xchg_handle x);
/* The fused form of push: b2r_query (device-resident outputs only) whose kernels ALSO store every final list into the mailboxes of
 * `x` as they emit it -- the finalize of the scoring kernels, the exact fix-up -- and whose last kernel publishes the arrival on
 * its way out.  No exchange kernel, no collective, no stream memory operation is enqueued; the call's first kernel holds the stream
 * until the mailbox slot's previous contents (four slots, round-robin) have been merged by every rank.  The batch is merged by
 * b2r_xchg_merge like a pushed one, except that the merge kernel itself checks the arrival words.  Better: hand the NEXT
 * b2r_query_push the merge outputs (merge_rows / merge_dist / merge_count, device; NULL = none) -- the oldest batch pushed before
 * this call and not merged yet is then merged INSIDE this call's last kernel (its lists arrived a whole scan ago), so a step of
 * a query stream carries its exchange without a single extra launch; b2r_xchg_merge is only needed for the last batch of the
 * stream.  At most four batches may be pushed and not yet merged.  Collective like push / merge: the same sequence of calls,
 * the same nq and k on every rank.                                                                                            */
int b2r_query_push(b2r_handle h, b2r_xchg_handle x, const float *q, int nq, int k, const b2r_filter *filter,
                   int64_t *out_rows, float *out_dist, int32_t *out_count, int64_t *merge_rows, float *merge_dist,
                   int32_t *merge_count, void *stream);

/* Host-side id table of a collection: string id <-> dense row number.  It stands where Chroma keeps its id index in sqlite
 * (`embeddings.embedding_id`, consulted by collection.add / upsert / get(ids) / delete(ids): app/utils/embedder.py:518, 632,
 * 640, 888): at 10M ids an interpreter-side dict costs a cache miss per id (~0.6 us, 4.7 ms per 8192-row upsert -- more than
 * the device's share of the step); here a batch is hashed while its home slots are prefetched, then probed.  No device work,
 * usable without a GPU.  A batch of ids = `bytes` + `offsets[n + 1]`: id i is bytes[offsets[i] .. offsets[i + 1] - gap), i.e.
 * `gap` separator bytes follow every id (0 = packed, 1 = what "\0".join(ids) gives a binding in one call).  Rows are the
 * dense insertion indices of b2r_ingest_f32; a row keeps its id bytes after it is erased (b2r_idtab_ids_of still answers).  */
typedef struct b2r_idtab *b2r_idtab_handle;
int b2r_idtab_create(int64_t reserve_ids, b2r_idtab_handle *out);
int b2r_idtab_destroy(b2r_idtab_handle t);
int b2r_idtab_clear(b2r_idtab_handle t);
int64_t b2r_idtab_live(b2r_idtab_handle t);                 /* ids currently mapped                          */
int64_t b2r_idtab_rows(b2r_idtab_handle t);                 /* rows appended so far (live or not)            */
/* rows_out[i] = the row id i maps to, or -1.  first_dup (may be NULL) = index of the first id of the batch that repeats an
 * earlier id of the same batch, or -1 (Chroma rejects such a batch: "Expected IDs to be unique").                          */
int b2r_idtab_lookup(b2r_idtab_handle t, const char *bytes, const int64_t *offsets, int64_t n, int64_t gap,
                     int64_t *rows_out, int64_t *first_dup);
/* The batch becomes rows first_row .. first_row + n - 1 (first_row must equal b2r_idtab_rows: the value b2r_ingest_f32
 * reported) and every id is (re-)pointed at its new row; prev_out[i] (may be NULL) = the row it pointed at before, or -1. */
int b2r_idtab_append(b2r_idtab_handle t, const char *bytes, const int64_t *offsets, int64_t n, int64_t gap,
                     int64_t first_row, int64_t *prev_out);
/* Unmaps the ids of these rows (delete); an id that points at a newer row by now stays mapped.                            */
int b2r_idtab_erase_rows(b2r_idtab_handle t, const int64_t *rows, int64_t n);
/* The ids of `rows`, NUL-terminated, back to back: id i = out[offsets_out[i] .. offsets_out[i + 1] - 1).  *need = bytes
 * required; nothing is written when out_cap < *need (call again with a larger buffer).                                    */
int b2r_idtab_ids_of(b2r_idtab_handle t, const int64_t *rows, int64_t n, char *out, int64_t out_cap, int64_t *offsets_out,
                     int64_t *need);

/* Diagnostics used by bench.py: run only the scoring/selection kernel selected by
 * `path` on device-resident prepared inputs, so it can be timed alone.
 *   path 0 = automatic (what b2r_query picks), 1 = warp-shuffle scan, 2 = tcgen05,
 *   3 = exact fp64 scan.                                                            */
int b2r_set_path(b2r_handle h, int path);
/* number of kernel launches issued by this handle since creation                    */
int64_t b2r_launch_count(b2r_handle h);
/* When enabled, every scoring-kernel launch (scan / tcgen05 / forced exact) is bracketed by
 * CUDA events on the caller's stream; b2r_kernel_time_ms waits for them and returns the
 * accumulated device time and launch count (reset != 0 clears the accumulators).      */
int b2r_set_kernel_timing(b2r_handle h, int enable);
int b2r_kernel_time_ms(b2r_handle h, double *total_ms, int64_t *launches, int reset);
/* Development: with B2R_TRACE=1 in the environment when the handle is created, the tcgen05 scoring kernel records eight
 * values per CTA of its last launch: four %globaltimer readings (epilogue start, samples posted, bound seeded, slice done; ns)
 * and four wait totals in SM cycles (MMA thread: empty accumulator, operands; first epilogue warp: full accumulator;
 * TMA producer: free ring slot).  Copies min(max_ctas, CTAs of that launch) * 8 values to the host array `out` after
 * synchronising the device; *n_ctas =
 * number of CTAs copied (0 when tracing is off).                                                       */
int b2r_debug_trace(b2r_handle h, uint64_t *out, int max_ctas, int *n_ctas);

#ifdef __cplusplus
}
#endif
#endif /* B2R_H_ */
