"""Drop-in for the duck-typed Chroma client/collection the reference holds in
``EmbeddingManager.client`` / ``.collection``.

Reference seam: ``chromadb.Client(...)`` + ``get_collection`` /
``create_collection(name, metadata)`` / ``delete_collection`` (app/utils/embedder.py:170-183,
669-678) and the collection methods ``add`` (:518), ``query`` (:596, :901), ``get`` (:632, :888),
``delete`` (:640), ``count`` (:700) -- same keyword arguments, same result shapes (nested lists
per query for ``query``, flat lists for ``get``), Chroma's error behaviour for bad input.

Vectors, scoring, selection and the exact re-rank live on the GPU behind the C ABI
(include/b2r.h); documents and metadata dicts stay in host tables keyed by the dense row number
the device reports, the id <-> row index is the library's native id table (b2r_idtab_*, idtable.py).
No CPU fallback exists: constructing a collection without the CUDA library or without a B200 raises.
"""
from __future__ import annotations

import ctypes
import json
import logging
import os
import threading

import numpy as np

from . import _lib
from .idtable import IdTable, IdsByRow, RowOfId, encode_ids
from .where import MetaTable, doc_mask, pack_bits

logger = logging.getLogger(__name__)

_INCLUDE_QUERY = ("metadatas", "documents", "distances")
_INCLUDE_GET = ("metadatas", "documents")
_VALID_INCLUDE = {"metadatas", "documents", "distances", "embeddings"}


def _torch():
    import sys
    return sys.modules.get("torch")


class _Matrix:
    """A [n, d] fp32 row-major matrix somewhere (host numpy or CUDA torch) + its raw pointer."""

    def __init__(self, x, what="embeddings"):
        t = _torch()
        self.keep = None
        self.stream = 0
        if t is not None and isinstance(x, t.Tensor):
            if x.dim() == 1:
                x = x[None]
            if x.dim() != 2:
                raise ValueError(f"{what} must be a 2-D [n, dim] array")
            x = x.detach().to(dtype=t.float32).contiguous()
            self.keep, self.n, self.d, self.ptr = x, int(x.shape[0]), int(x.shape[1]), x.data_ptr()
            self.is_cuda = x.is_cuda
            if x.is_cuda:
                self.device = x.device.index if x.device.index is not None else t.cuda.current_device()
                self.stream = t.cuda.current_stream(x.device).cuda_stream
            return
        try:
            a = np.asarray(x, dtype=np.float32)
        except (ValueError, TypeError) as e:
            raise ValueError(f"{what} must be a list of equal-length float vectors: {e}") from None
        if a.ndim == 1 and a.size and not isinstance(x[0], (list, tuple, np.ndarray)):
            a = a[None]
        if a.ndim != 2:
            raise ValueError(f"{what} must be a 2-D [n, dim] array")
        a = np.ascontiguousarray(a)
        self.keep, self.n, self.d, self.ptr, self.is_cuda = a, int(a.shape[0]), int(a.shape[1]), a.ctypes.data, False


class B200Collection:
    """Exact GPU collection with ``chromadb.Collection``'s add/upsert/query/get/delete/count."""

    def __init__(self, name="multimodal_rag", metadata=None, *, device=0, capacity=0,
                 keep_f32_master=True, dimension=None):
        self.name = name
        self.metadata = dict(metadata) if metadata else None
        space = (metadata or {}).get("hnsw:space", "l2")          # Chroma's default space
        if space not in _lib.SPACE_CODE:
            raise ValueError(f"hnsw:space must be one of {sorted(_lib.SPACE_CODE)}, got {space!r}")
        self.space = space
        self.device = int(device)
        self._capacity = int(capacity)
        self._flags = 0 if keep_f32_master else _lib.FLAG_NO_F32_MASTER
        self._lib = _lib.load()
        self._h = None
        self._dim = None
        self._lock = threading.RLock()
        self._idtab = IdTable(reserve=self._capacity)    # id -> live row, row -> id bytes (native, b2r_idtab_*)
        self._ids = IdsByRow(self._idtab)                # list-like / dict-like read-only views of it
        self._row_of = RowOfId(self._idtab)
        self._nrows = 0                                  # rows appended so far (live or tombstoned)
        self._docs: list = []
        self._alive_buf = np.zeros(1024, dtype=bool)     # tombstone table, grown x2 (amortised O(1) per row)
        self._meta = MetaTable()
        self.device_where = True      # general where clauses run on the device (False: host-evaluated bitmaps)
        self._where_cache: dict = {}  # (clause as JSON, table rows) -> compiled clause
        if dimension is not None:
            self._open(int(dimension))

    # ---- handle --------------------------------------------------------------------
    def _open(self, dim: int):
        h = ctypes.c_void_p()
        _lib.check(self._lib.b2r_create(dim, _lib.SPACE_CODE[self.space], self._capacity, self.device,
                                        self._flags, ctypes.byref(h)), "b2r_create")
        self._h, self._dim = h, dim

    def close(self):
        with self._lock:
            if self._h is not None:
                self._lib.b2r_destroy(self._h)
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def dimension(self):
        return self._dim

    @property
    def _alive(self) -> np.ndarray:
        """bool per appended row: not tombstoned (a view of the growable table)"""
        return self._alive_buf[: self._nrows]

    def _alive_extend(self, n_new: int) -> None:
        have, need = self._nrows, self._nrows + n_new
        if need > self._alive_buf.shape[0]:
            grown = np.zeros(max(need, 2 * self._alive_buf.shape[0]), dtype=bool)
            grown[:have] = self._alive_buf[:have]
            self._alive_buf = grown
        self._alive_buf[have:need] = True

    @property
    def handle(self):
        return self._h

    # ---- validation ------------------------------------------------------------------
    def _validate_batch(self, ids, embeddings, metadatas, documents):
        """-> (ids as a list, the batch encoded for the id table, the row each id maps to now (-1 = new), the embedding
        matrix).  Nothing runs per id in the interpreter: one join + encode, and the table's lookup reports repeats."""
        if ids is None or isinstance(ids, str):
            ids = [ids] if isinstance(ids, str) else ids
        if not isinstance(ids, (list, tuple)):
            raise ValueError("Expected ids to be a list of str")
        n = len(ids)
        try:
            enc = encode_ids(ids)
            bad = enc.has_empty()
        except TypeError:
            bad = True
        if bad:
            bad = next(i for i in ids if not (isinstance(i, str) and i))
            raise ValueError(f"Expected ID to be a non-empty str, got {bad!r}")
        found, first_dup = self._idtab.lookup(enc, want_dup=True)
        if first_dup >= 0:
            seen, dup = set(), []
            for i in ids:
                if i in seen:
                    dup.append(i)
                seen.add(i)
            raise ValueError(f"Expected IDs to be unique, found duplicates of: {', '.join(dup[:5])}")
        if embeddings is None:
            raise ValueError("embeddings are required: this collection has no embedding function")
        m = _Matrix(embeddings)
        if m.n != n:
            raise ValueError(f"Number of embeddings {m.n} must match number of ids {n}")
        for name, lst in (("metadatas", metadatas), ("documents", documents)):
            if lst is not None and len(lst) != n:
                raise ValueError(f"Number of {name} {len(lst)} must match number of ids {n}")
        if metadatas is not None:
            MetaTable.validate_batch(metadatas)
        if n and self._dim is not None and m.d != self._dim:
            raise ValueError(f"Embedding dimension {m.d} does not match collection dimensionality {self._dim}")
        return ids, enc, found, m

    # ---- mutations -------------------------------------------------------------------
    def _append_rows(self, ids, enc, m: _Matrix, sel, metadatas, documents):
        """Ingest rows `sel` (indices into the batch; None = the whole batch) and register them on the host."""
        whole = sel is None or len(sel) == m.n
        n = m.n if whole else len(sel)
        if n == 0:
            return
        if self._h is None:
            self._open(m.d)
        if metadatas is None:
            metas, codes_ptr = None, None                  # bare vectors: type code 0 for every row, no per-row host work
        else:
            metas = list(metadatas) if whole else [metadatas[i] for i in sel]
            # type codes must be known before the device call; MetaTable only commits below
            codes = np.asarray([self._meta.type_code_of(md) for md in metas], dtype=np.uint8)
            codes_ptr = codes.ctypes.data
        if whole:
            ptr, keep = m.ptr, m.keep
        elif m.is_cuda:
            keep = m.keep[_torch().as_tensor(sel, device=m.keep.device)].contiguous()
            ptr = keep.data_ptr()
        else:
            keep = np.ascontiguousarray(m.keep[sel])
            ptr = keep.ctypes.data
        first = ctypes.c_int64(-1)
        _lib.check(self._lib.b2r_ingest_f32(self._h, ptr, n, codes_ptr, ctypes.byref(first), m.stream), "b2r_ingest_f32")
        assert first.value == self._nrows, "host tables out of step with the device corpus"
        self._alive_extend(n)
        self._idtab.append(enc if whole else encode_ids([ids[i] for i in sel]), first.value)   # (re-)points the ids at the new rows
        self._nrows += n
        if documents is None:
            self._docs.extend([None] * n)
        else:
            self._docs.extend(documents if whole else [documents[i] for i in sel])
        if metas is None:
            self._meta.append_none(n)
        else:
            self._meta.append_batch(metas)
            self._push_columns(first.value, n, {k for md in metas if md for k in md}, m.stream)
        del keep

    def _push_columns(self, first: int, n: int, keys, stream=0):
        """Mirror the dictionary codes of rows [first, first + n) of the touched metadata keys to the device
        columns (b2r_column_set); keys beyond the 16 device columns stay host-only."""
        for key in keys:
            ci = self._meta.device_column(key)
            if ci is None:
                continue
            have = self._meta.cols[key].codes[first: first + n]      # the column may end before the batch does
            codes = np.full(n, -1, dtype=np.int32)
            codes[: have.shape[0]] = have
            _lib.check(self._lib.b2r_column_set(self._h, ci, first, n, codes.ctypes.data, stream), "b2r_column_set")

    def _kill_rows(self, rows, stream=0):
        if len(rows) == 0:
            return
        arr = np.ascontiguousarray(rows, dtype=np.int64)
        _lib.check(self._lib.b2r_tombstone(self._h, arr.ctypes.data, arr.shape[0], stream), "b2r_tombstone")
        self._alive_buf[arr] = False
        self._idtab.erase_rows(arr)

    def add(self, ids, embeddings=None, metadatas=None, documents=None):
        """Chroma ``Collection.add``: ids already present are skipped (with a warning)."""
        with self._lock:
            ids, enc, found, m = self._validate_batch(ids, embeddings, metadatas, documents)
            sel = None
            if found.size and found.max() >= 0:
                sel = np.flatnonzero(found < 0).tolist()
                logger.warning("Add of existing embedding ID(s) skipped: %d of %d", len(ids) - len(sel), len(ids))
            self._append_rows(ids, enc, m, sel, metadatas, documents)

    def upsert(self, ids, embeddings=None, metadatas=None, documents=None):
        """Chroma ``Collection.upsert``: existing ids are overwritten.  The new rows are appended FIRST and the old versions
        tombstoned after that succeeded, so a failed ingest (out of memory while growing, a bad batch) leaves the old
        versions in place; a query never sees both because both steps are enqueued on one stream before it."""
        with self._lock:
            ids, enc, found, m = self._validate_batch(ids, embeddings, metadatas, documents)
            old = found[found >= 0]
            if self._h is None and ids:
                self._open(m.d)
            self._append_rows(ids, enc, m, None, metadatas, documents)  # re-points the ids at the new rows
            if old.size:
                arr = np.ascontiguousarray(old)
                _lib.check(self._lib.b2r_tombstone(self._h, arr.ctypes.data, arr.shape[0], m.stream), "b2r_tombstone")
                self._alive_buf[arr] = False

    def delete(self, ids=None, where=None, where_document=None):
        """Chroma ``Collection.delete``: by ids and / or a where clause.  With neither, chromadb 0.4.22 raises instead of
        wiping the collection (the reference empties a collection through delete_collection, embedder.py:669-678)."""
        if (ids is None or (not isinstance(ids, str) and len(ids) == 0)) and not where and not where_document:
            raise ValueError("You must provide either ids, where, or where_document to delete.")
        with self._lock:
            rows = self._select_rows(ids, where, where_document)
            self._kill_rows(rows)
            return self._idtab.ids_of(rows)

    def rows_of(self, ids) -> np.ndarray:
        """int64 per id: the live row it maps to, or -1 (one native lookup for the batch)"""
        with self._lock:
            return self._idtab.lookup(encode_ids(list(ids))) if len(ids) else np.empty(0, dtype=np.int64)

    def ids_of(self, rows) -> list:
        """the ids of device rows (live or tombstoned), one native call for the batch"""
        with self._lock:
            return self._idtab.ids_of(rows)

    def count(self) -> int:
        with self._lock:
            if self._h is None:
                return 0
            n = int(self._lib.b2r_count(self._h))
            assert n == self._idtab.live, "host tables out of step with the device corpus"
            return n

    # ---- reads -----------------------------------------------------------------------
    def _where_mask(self, where, where_document=None) -> np.ndarray:
        """bool per row: live AND matching `where` (AND `where_document`).  The clause runs on the device columns
        (b2r_filter_eval) like a query's would; clauses the device cannot run fall back to the host tables inside `_filter`."""
        if self._h is None or not self._nrows:
            return np.zeros(self._nrows, dtype=bool)
        return self.filter_bits(where, where_document)

    def _select_rows(self, ids, where, where_document=None):
        if ids is not None:
            if isinstance(ids, str):
                ids = [ids]
            try:
                found = self._idtab.lookup(encode_ids(ids)) if len(ids) else np.empty(0, dtype=np.int64)
            except TypeError:
                raise ValueError("Expected ids to be a list of str") from None
            rows = np.unique(found[found >= 0]).tolist()
            if (where or where_document) and rows:
                mask = self._where_mask(where, where_document)
                rows = [r for r in rows if mask[r]]
            return rows
        if where or where_document:
            mask = self._where_mask(where, where_document)
        else:
            mask = self._alive
        return np.flatnonzero(mask).tolist()

    def _fetch_rows(self, rows):
        out = np.empty((len(rows), self._dim), dtype=np.float32)
        if rows:
            arr = np.asarray(rows, dtype=np.int64)
            _lib.check(self._lib.b2r_get_rows_f32(self._h, arr.ctypes.data, arr.shape[0], out.ctypes.data, 0),
                       "b2r_get_rows_f32")
        return out

    @staticmethod
    def _check_include(include, allowed):
        include = list(include)
        for key in include:
            if key not in _VALID_INCLUDE or key not in allowed:
                raise ValueError(f"Expected include item to be one of {sorted(allowed)}, got {key}")
        return include

    def get(self, ids=None, where=None, limit=None, offset=None, include=_INCLUDE_GET, where_document=None):
        include = self._check_include(include, {"metadatas", "documents", "embeddings"})
        with self._lock:
            rows = self._select_rows(ids, where, where_document)
            if offset:
                rows = rows[offset:]
            if limit is not None:
                rows = rows[:limit]
            out = {"ids": self._idtab.ids_of(rows), "embeddings": None, "metadatas": None, "documents": None}
            if "embeddings" in include:
                out["embeddings"] = self._fetch_rows(rows).tolist() if rows else []
            if "metadatas" in include:
                out["metadatas"] = [self._meta.meta[r] for r in rows]
            if "documents" in include:
                out["documents"] = [self._docs[r] for r in rows]
            return out

    def _filter(self, where, where_document=None):
        """where (+ where_document) -> (B2RFilter, keep-alive object)."""
        f = _lib.B2RFilter(type_mask=(1 << 64) - 1, allow_bits=None, where=None)
        doc_bits = None
        if where_document:                    # host pass over the documents -> allow bitmap, ANDed on the device with the rest
            doc_bits = pack_bits(doc_mask(self._docs, where_document))
            f.allow_bits = doc_bits.ctypes.data
        if not where:
            return f, doc_bits
        tm = self._meta.type_only_mask(where)
        if tm is not None:
            f.type_mask = tm
            return f, doc_bits
        if self.device_where:                # the clause runs on the device against the metadata columns
            # compiled clauses are kept while the table does not grow (a chat session repeats its filter; the device
            # side recognises the repeat by the clause's hash and re-uses its bitmaps)
            try:
                key = (json.dumps(where, sort_keys=True, default=repr), self._meta.nrows)
            except (TypeError, ValueError):
                key = None
            hit = self._where_cache.get(key) if key is not None else None
            if hit is None:
                prog = self._meta.compile(where)
                if prog is not None:
                    nodes, lut = prog
                    w = _lib.B2RWhere(n_nodes=len(nodes), lut=lut.ctypes.data if lut.size else None, lut_words=int(lut.size))
                    for i, (op, col, off, nv) in enumerate(nodes):
                        w.nodes[i] = _lib.B2RWhereNode(op, col, off, nv)
                    hit = (w, lut)
                    if key is not None:
                        if len(self._where_cache) >= 64:
                            self._where_cache.clear()
                        self._where_cache[key] = hit
            if hit is not None:
                f.where = ctypes.pointer(hit[0])
                return f, (hit, doc_bits)
        mask = self._meta.mask(where)                 # clause the device cannot run: host-evaluated bitmap
        if where_document:
            mask = mask & doc_mask(self._docs, where_document)
        bits = pack_bits(mask)
        f.allow_bits = bits.ctypes.data
        return f, bits

    def device_filter(self, where=None):
        """(b2r_filter struct, keep-alive object) for `where`, as `query` would hand it to the C ABI -- for callers that
        drive b2r_query themselves with device-resident buffers (bench.py).  Keep the second value alive while the
        struct is in use."""
        with self._lock:
            return self._filter(where)

    def filter_bits(self, where=None, where_document=None) -> np.ndarray:
        """The pass bitmap the device derives for `where` / `where_document` (bool per row: live AND matching) -- b2r_filter_eval."""
        with self._lock:
            n = self._nrows
            words = np.zeros((n + 31) // 32, dtype=np.uint32)
            if self._h is not None and n:
                f, keep = self._filter(where, where_document)
                _lib.check(self._lib.b2r_filter_eval(self._h, ctypes.byref(f), words.ctypes.data, 0), "b2r_filter_eval")
                del keep
            return np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)

    def query_rows(self, query_embeddings, n_results=10, where=None, want_dist64=False, where_document=None):
        """Device query returning numpy arrays: rows [nq,k] int64 (-1 pad), dist [nq,k] fp32,
        count [nq] int32 (and fp64 distances when asked)."""
        with self._lock:
            if not isinstance(n_results, int) or isinstance(n_results, bool) or n_results <= 0:
                raise ValueError(f"Expected n_results to be a positive integer, got {n_results}")
            m = _Matrix(query_embeddings, "query_embeddings")
            if m.n == 0:
                raise ValueError("Expected query_embeddings to be a non-empty list")
            if self._dim is not None and m.d != self._dim:
                raise ValueError(f"Query dimension {m.d} does not match collection dimensionality {self._dim}")
            k = n_results
            rows = np.full((m.n, k), -1, dtype=np.int64)
            dist = np.full((m.n, k), np.inf, dtype=np.float32)
            cnt = np.zeros((m.n,), dtype=np.int32)
            d64 = np.full((m.n, k), np.inf, dtype=np.float64) if want_dist64 else None
            if self._h is None or not self._idtab.live:
                return (rows, dist, cnt, d64) if want_dist64 else (rows, dist, cnt)
            f, keep = self._filter(where, where_document)
            _lib.check(self._lib.b2r_query_ex(self._h, m.ptr, m.n, k, ctypes.byref(f), rows.ctypes.data,
                                              dist.ctypes.data, None if d64 is None else d64.ctypes.data,
                                              cnt.ctypes.data, m.stream), "b2r_query")
            del keep
            return (rows, dist, cnt, d64) if want_dist64 else (rows, dist, cnt)

    def query_rows_pipelined(self, batches, n_results=10, where=None):
        """Several query batches through the pipelined form of the call (b2r_query_async / b2r_wait): two batches in
        flight, the host<->device copies of one overlap the kernels of the other.  `batches` = host arrays [nq_i, dim];
        returns a list of (rows, dist, count) like `query_rows`.  Only unfiltered or `{"type": ...}` queries take this
        route; any other clause is answered batch by batch."""
        with self._lock:
            if not isinstance(n_results, int) or isinstance(n_results, bool) or n_results <= 0:
                raise ValueError(f"Expected n_results to be a positive integer, got {n_results}")
            f, keep = self._filter(where)
            if self._h is None or not self._idtab.live or f.allow_bits or f.where:
                return [self.query_rows(b, n_results, where) for b in batches]
            k, out, pending = n_results, [], []
            for b in batches:
                m = _Matrix(b, "query_embeddings")
                if m.is_cuda:
                    raise ValueError("query_rows_pipelined takes host batches; device batches only enqueue anyway")
                if m.n == 0 or m.d != self._dim:
                    raise ValueError(f"Query dimension {m.d} does not match collection dimensionality {self._dim}")
                rows = np.empty((m.n, k), dtype=np.int64)
                dist = np.empty((m.n, k), dtype=np.float32)
                cnt = np.empty((m.n,), dtype=np.int32)
                t = ctypes.c_uint64()
                _lib.check(self._lib.b2r_query_async(self._h, m.ptr, m.n, k, ctypes.byref(f), rows.ctypes.data,
                                                     dist.ctypes.data, cnt.ctypes.data, 0, ctypes.byref(t)), "b2r_query_async")
                pending.append((t.value, m))                      # keep the batch alive until its ticket is waited for
                out.append((rows, dist, cnt))
                if len(pending) == 2:
                    _lib.check(self._lib.b2r_wait(self._h, pending.pop(0)[0]), "b2r_wait")
            for t, _ in pending:
                _lib.check(self._lib.b2r_wait(self._h, t), "b2r_wait")
            return out

    def query(self, query_embeddings=None, n_results=10, where=None, where_document=None,
              include=_INCLUDE_QUERY, query_texts=None):
        """Chroma ``Collection.query``: nested lists, one inner list per query, ascending
        distance; keys not in ``include`` are None."""
        if query_texts is not None and query_embeddings is None:
            raise ValueError("query_texts needs an embedding function; pass query_embeddings")
        include = self._check_include(include, _VALID_INCLUDE)
        with self._lock:
            rows, dist, cnt = self.query_rows(query_embeddings, n_results, where, where_document=where_document)
            nq = rows.shape[0]
            live = self._idtab.live
            if n_results > live:
                logger.warning("Number of requested results %d is greater than number of elements in index %d, "
                               "updating n_results = %d", n_results, live, live)
            res = {"ids": [], "distances": None, "metadatas": None, "documents": None, "embeddings": None}
            for key in include:
                res[key] = []
            flat = self._idtab.ids_of(np.concatenate([rows[i, : cnt[i]] for i in range(nq)])) if nq else []
            at = 0
            for i in range(nq):
                rr = rows[i, : cnt[i]].tolist()
                res["ids"].append(flat[at: at + len(rr)])
                at += len(rr)
                if res["distances"] is not None:
                    res["distances"].append(dist[i, : cnt[i]].tolist())
                if res["metadatas"] is not None:
                    res["metadatas"].append([self._meta.meta[r] for r in rr])
                if res["documents"] is not None:
                    res["documents"].append([self._docs[r] for r in rr])
                if res["embeddings"] is not None:
                    res["embeddings"].append(self._fetch_rows(rr).tolist() if rr else [])
            return res

    # ---- persistence (the reference persists through ChromaSettings(persist_directory=...), embedder.py:164-170) ----
    _FORMAT = 1

    def save(self, directory: str) -> None:
        """Write the collection under `directory`: `<name>.b2r` = the device shard exactly as it sits in HBM
        (b2r_save), `<name>.tables.json` = ids / documents / metadata / tombstones kept on the host."""
        with self._lock:
            os.makedirs(directory, exist_ok=True)
            base = os.path.join(directory, self.name)
            if self._h is not None:
                _lib.check(self._lib.b2r_save(self._h, (base + ".b2r").encode(), 0), "b2r_save")
            tables = {"format": self._FORMAT, "name": self.name, "metadata": self.metadata, "space": self.space,
                      "dimension": self._dim, "keep_f32_master": not (self._flags & _lib.FLAG_NO_F32_MASTER),
                      "ids": self._ids[:], "documents": self._docs, "metadatas": self._meta.meta,
                      "dead_rows": np.flatnonzero(~self._alive).tolist(),
                      "type_codes": self._meta.type_codes, "type_overflow": self._meta.type_overflow}
            tmp = base + ".tables.json.tmp"
            with open(tmp, "w", encoding="utf-8") as f:
                json.dump(tables, f)
            os.replace(tmp, base + ".tables.json")

    @classmethod
    def load(cls, directory: str, name: str, *, device=0, capacity=0) -> "B200Collection":
        """Rebuild a saved collection; the device shard is uploaded as stored (no re-normalisation), so queries
        answer bit-identically to the collection that was saved."""
        base = os.path.join(directory, name)
        with open(base + ".tables.json", encoding="utf-8") as f:
            t = json.load(f)
        if t.get("format") != cls._FORMAT:
            raise ValueError(f"{base}.tables.json: unknown format {t.get('format')!r}")
        c = cls(t["name"], t["metadata"], device=device, capacity=capacity, keep_f32_master=t["keep_f32_master"])
        if t["dimension"] is not None:
            h = ctypes.c_void_p()
            _lib.check(c._lib.b2r_load((base + ".b2r").encode(), int(device), int(capacity), ctypes.byref(h)), "b2r_load")
            c._h, c._dim = h, int(t["dimension"])
        c._docs = list(t["documents"])
        c._nrows = len(t["ids"])
        c._idtab.append(encode_ids(t["ids"]), 0)         # an id stored more than once ends at its last row ...
        c._meta.type_codes = {k: int(v) for k, v in t["type_codes"].items()}
        c._meta.type_overflow = bool(t["type_overflow"])
        c._meta.append_batch(t["metadatas"])
        c._alive_buf = np.ones(max(c._nrows, 1024), dtype=bool)
        dead = np.asarray(t["dead_rows"], dtype=np.int64)
        c._alive_buf[dead] = False
        c._idtab.erase_rows(dead)                        # ... and a deleted id whose last row is dead is unmapped
        if c._h is not None:
            c._push_columns(0, c._nrows, list(c._meta.cols))
            st = c.stats()
            if st["rows"] != c._nrows or st["live"] != c._idtab.live or st["dim"] != c._dim:
                c.close()
                raise ValueError(f"{base}: shard file and host tables disagree "
                                 f"(rows {st['rows']} vs {c._nrows}, live {st['live']} vs {c._idtab.live})")
        return c

    def stats(self) -> dict:
        with self._lock:
            if self._h is None:
                return {"rows": 0, "live": 0}
            st = _lib.B2RStats()
            _lib.check(self._lib.b2r_get_stats(self._h, ctypes.byref(st)), "b2r_get_stats")
            out = {name: getattr(st, name) for name, _ in st._fields_}
            out["launches"] = int(self._lib.b2r_launch_count(self._h))
            return out

    def set_path(self, path: int):
        """Diagnostics: 0 auto, 1 warp-shuffle scan, 2 tcgen05, 3 exact fp64 scan."""
        _lib.check(self._lib.b2r_set_path(self._h, int(path)), "b2r_set_path")


class B200Client:
    """The ``chromadb.Client`` trio the reference uses (app/utils/embedder.py:170-183, 669-678)."""

    def __init__(self, device=0, keep_f32_master=True, default_capacity=0, path=None):
        """`path` = the persist directory (the reference's ChromaSettings(persist_directory=...), embedder.py:164-168):
        collections saved there are loaded now, `persist()` writes them back."""
        self.device = device
        self.keep_f32_master = keep_f32_master
        self.default_capacity = default_capacity
        self.path = path
        self._collections: dict[str, B200Collection] = {}
        self._lock = threading.Lock()
        _lib.load()    # fail now, loudly, if the CUDA library is absent
        if path and os.path.isdir(path):
            for fn in sorted(os.listdir(path)):
                if fn.endswith(".tables.json"):
                    name = fn[: -len(".tables.json")]
                    self._collections[name] = B200Collection.load(path, name, device=device, capacity=default_capacity)

    def persist(self) -> None:
        """Write every collection to the persist directory (chromadb's client.persist())."""
        if not self.path:
            raise ValueError("this client was created without a persist directory")
        with self._lock:
            for c in self._collections.values():
                c.save(self.path)

    def create_collection(self, name, metadata=None, get_or_create=False, **kw):
        with self._lock:
            if name in self._collections:
                if get_or_create:
                    return self._collections[name]
                raise ValueError(f"Collection {name} already exists.")
            c = B200Collection(name, metadata, device=kw.pop("device", self.device),
                               capacity=kw.pop("capacity", self.default_capacity),
                               keep_f32_master=kw.pop("keep_f32_master", self.keep_f32_master), **kw)
            self._collections[name] = c
            return c

    def get_collection(self, name):
        with self._lock:
            if name not in self._collections:
                raise ValueError(f"Collection {name} does not exist.")
            return self._collections[name]

    def get_or_create_collection(self, name, metadata=None, **kw):
        return self.create_collection(name, metadata, get_or_create=True, **kw)

    def delete_collection(self, name):
        with self._lock:
            if name not in self._collections:
                raise ValueError(f"Collection {name} does not exist.")
            self._collections.pop(name).close()
            if self.path:
                for ext in (".b2r", ".tables.json"):
                    try:
                        os.remove(os.path.join(self.path, name + ext))
                    except FileNotFoundError:
                        pass

    def list_collections(self):
        with self._lock:
            return list(self._collections.values())
