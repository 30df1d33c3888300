"""CPU oracle for the vector add/query hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product package
(``multimodal_rag_b200``) never does and fails loudly without its CUDA library.

What it restates
----------------
The reference delegates all arithmetic of the path to two un-vendored third-party
packages, ``chromadb==0.4.22`` (-> ``chroma-hnswlib==0.7.3``), pinned at
``/root/reference/requirements.txt:21`` and called from
``/root/reference/app/utils/embedder.py:518`` (``collection.add``), ``:596,901``
(``collection.query``), ``:632,888`` (``collection.get``), ``:640``
(``collection.delete``) and ``:700`` (``collection.count``).  Neither package is in
``/root/reference`` nor installable in this image, and the reference has no tests:
**parity unpinned** against a real Chroma.  What *is* pinned: the 70 real MiniLM
vectors in the reference's committed ``chroma_db/chroma.sqlite3`` WAL and the
fp64 known answers computed from them (``tests/golden/``).

Published algorithm restated here (hnswlib ``space_l2.h`` / ``space_ip.h`` /
``bindings.cpp`` normalisation, Chroma ``SegmentAPI._query`` result shape):

* ``l2``     : squared Euclidean  sum_i (q_i - x_i)^2
* ``ip``     : 1 - sum_i q_i x_i
* ``cosine`` : vectors and queries are first scaled in fp32 by
               1 / (sqrt(sum x^2) + 1e-30), then 1 - sum_i q^_i x^_i

Chroma's HNSW search is approximate; this oracle (and the engine) return the
*exact* answer HNSW approximates.  Canonical choices where upstream is
unspecified: distances are accumulated in fp64 from the fp32 inputs and cast to
fp32 on output; results are ordered by (fp64 distance, insertion index) so ties
resolve to the earliest-inserted row; a missing metadata key never matches a
``where`` operator.
"""
from __future__ import annotations

import numpy as np

SPACES = ("l2", "cosine", "ip")


# --------------------------------------------------------------------------
# arithmetic
# --------------------------------------------------------------------------
def normalize_f32(x: np.ndarray) -> np.ndarray:
    """hnswlib python binding ``normalize_vector`` (cosine space): fp32 scale by
    1/(sqrt(sum x^2)+1e-30).  The sum of squares is taken in fp64 and rounded to
    fp32 once so the result does not depend on SIMD summation order."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    s = (x.astype(np.float64) ** 2).sum(axis=-1).astype(np.float32)
    inv = np.float32(1.0) / (np.sqrt(s, dtype=np.float32) + np.float32(1e-30))
    return (x * inv[..., None]).astype(np.float32)


def distances_f64(q: np.ndarray, X: np.ndarray, space: str) -> np.ndarray:
    """[nq, n] fp64 distances between fp32 queries and fp32 rows.

    For ``cosine`` both operands must already be ``normalize_f32``-ed (that is how
    the rows are *stored*, as in hnswlib)."""
    q = np.asarray(q, dtype=np.float32).astype(np.float64)
    X = np.asarray(X, dtype=np.float32).astype(np.float64)
    out = np.empty((q.shape[0], X.shape[0]), dtype=np.float64)
    for i in range(q.shape[0]):           # per-row reduction, independent of row position
        if space == "l2":
            diff = X - q[i]
            out[i] = (diff * diff).sum(axis=1)
        elif space in ("ip", "cosine"):
            out[i] = 1.0 - (X * q[i]).sum(axis=1)
        else:
            raise ValueError(f"unknown space {space!r}")
    return out


def topk_exact(q, X, k, space, allowed=None):
    """Exact top-k rows per query ordered by (fp64 distance, row index).

    Returns (rows [nq][<=k] int64 lists, dists [nq][<=k] fp32 lists)."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float32))
    d = distances_f64(q, X, space)
    n = X.shape[0]
    cand = np.arange(n) if allowed is None else np.flatnonzero(allowed)
    rows, dists = [], []
    for i in range(q.shape[0]):
        di = d[i, cand]
        order = np.lexsort((cand, di))[: max(0, min(k, cand.size))]
        rows.append(cand[order].astype(np.int64))
        dists.append(di[order].astype(np.float32))
    return rows, dists


def topk_bruteforce_f32(q, X, k, space):
    """The 'numpy fp32 brute force' CPU baseline of BASELINE.md §4: one sgemm
    (OpenBLAS, all host cores) + argpartition + sort.  fp32 accumulation, so it
    is a *baseline*, not the parity oracle."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float32))
    X = np.asarray(X, dtype=np.float32)
    s = q @ X.T
    if space == "l2":
        d = (q * q).sum(1)[:, None] + (X * X).sum(1)[None, :] - 2.0 * s
    else:
        d = 1.0 - s
    k = min(k, X.shape[0])
    part = np.argpartition(d, k - 1, axis=1)[:, :k]
    pd = np.take_along_axis(d, part, axis=1)
    o = np.argsort(pd, axis=1, kind="stable")
    return np.take_along_axis(part, o, axis=1), np.take_along_axis(pd, o, axis=1)


# --------------------------------------------------------------------------
# where-clause (Chroma 0.4.22 grammar), evaluated per row with Python loops
# --------------------------------------------------------------------------
_CMP = {
    "$eq": lambda a, b: a == b,
    "$ne": lambda a, b: a != b,
    "$gt": lambda a, b: a > b,
    "$gte": lambda a, b: a >= b,
    "$lt": lambda a, b: a < b,
    "$lte": lambda a, b: a <= b,
}


def _same_kind(a, b):
    num = (int, float)
    if isinstance(a, bool) or isinstance(b, bool):
        return isinstance(a, bool) and isinstance(b, bool)
    if isinstance(a, num) and isinstance(b, num):
        return True
    return type(a) is type(b)


def where_match(meta: dict | None, where: dict | None) -> bool:
    if not where:
        return True
    meta = meta or {}
    if len(where) != 1:
        raise ValueError(f"Expected where to have exactly one operator, got {where}")
    (key, cond), = where.items()
    if key == "$and":
        return all(where_match(meta, w) for w in cond)
    if key == "$or":
        return any(where_match(meta, w) for w in cond)
    if key.startswith("$"):
        raise ValueError(f"unknown where operator {key}")
    if not isinstance(cond, dict):
        cond = {"$eq": cond}
    if len(cond) != 1:
        raise ValueError(f"Expected operator expression to have one operator, got {cond}")
    (op, val), = cond.items()
    if key not in meta:
        return False
    have = meta[key]
    if op in ("$in", "$nin"):
        hit = any(_same_kind(have, v) and have == v for v in val)
        return hit if op == "$in" else not hit
    if op not in _CMP:
        raise ValueError(f"unknown where operator {op}")
    if not _same_kind(have, val):
        return False
    return bool(_CMP[op](have, val))


def where_document_match(doc, where_document) -> bool:
    """Chroma ``where_document`` (chromadb 0.4.22 validate_where_document / the sqlite full-text filter, restated from
    memory like the rest of this file): $contains, $not_contains, $and, $or; a row without a document matches neither."""
    if not where_document:
        return True
    if len(where_document) != 1:
        raise ValueError(f"Expected where document to have exactly one operator, got {where_document}")
    (op, val), = where_document.items()
    if op == "$and":
        return all(where_document_match(doc, w) for w in val)
    if op == "$or":
        return any(where_document_match(doc, w) for w in val)
    if op not in ("$contains", "$not_contains"):
        raise ValueError(f"unknown where document operator {op}")
    if not isinstance(val, str) or not val:
        raise ValueError("where document operand must be a non-empty str")
    if doc is None:
        return False
    return (val in doc) if op == "$contains" else (val not in doc)


# --------------------------------------------------------------------------
# Collection with Chroma's add/upsert/query/get/delete/count semantics
# --------------------------------------------------------------------------
class ExactCollection:
    """Exact CPU stand-in for ``chromadb.Collection`` as used by
    ``app/utils/embedder.py`` (see module docstring for the call sites)."""

    def __init__(self, name="multimodal_rag", metadata=None):
        self.name = name
        self.metadata = dict(metadata or {})
        self.space = self.metadata.get("hnsw:space", "l2")   # Chroma default
        if self.space not in SPACES:
            raise ValueError(f"unknown hnsw:space {self.space!r}")
        self.dim = None
        self._ids: list[str] = []
        self._vec: list[np.ndarray] = []      # stored rows (normalised for cosine)
        self._meta: list[dict | None] = []
        self._doc: list[str | None] = []
        self._alive: list[bool] = []
        self._row_of: dict[str, int] = {}

    @property
    def dimension(self):
        return self.dim

    # -- helpers -----------------------------------------------------------
    def _coerce(self, embeddings):
        e = np.asarray(embeddings, dtype=np.float32)
        if e.ndim != 2:
            raise ValueError("embeddings must be a list of equal-length vectors")
        if self.dim is None:
            self.dim = int(e.shape[1])
        if e.shape[1] != self.dim:
            raise ValueError(f"Embedding dimension {e.shape[1]} does not match collection dimensionality {self.dim}")
        return e

    def _store(self, e):
        return normalize_f32(e) if self.space == "cosine" else e

    @staticmethod
    def _check_lengths(ids, *others):
        n = len(ids)
        for o in others:
            if o is not None and len(o) != n:
                raise ValueError("ids, embeddings, metadatas and documents must have equal lengths")
        if len(set(ids)) != n:
            raise ValueError("Expected IDs to be unique within one call")

    def _append(self, id_, v, m, d):
        self._row_of[id_] = len(self._ids)
        self._ids.append(id_); self._vec.append(v); self._meta.append(m); self._doc.append(d)
        self._alive.append(True)

    # -- mutations ---------------------------------------------------------
    def add(self, ids, embeddings, metadatas=None, documents=None):
        self._check_lengths(ids, embeddings, metadatas, documents)
        if not len(ids):
            return
        e = self._store(self._coerce(embeddings))
        for i, id_ in enumerate(ids):
            if id_ in self._row_of:          # Chroma: existing id -> warn + skip
                continue
            self._append(id_, e[i], None if metadatas is None else metadatas[i],
                         None if documents is None else documents[i])

    def upsert(self, ids, embeddings, metadatas=None, documents=None):
        self._check_lengths(ids, embeddings, metadatas, documents)
        if not len(ids):
            return
        e = self._store(self._coerce(embeddings))
        for i, id_ in enumerate(ids):
            if id_ in self._row_of:          # overwrite = tombstone + append
                self._alive[self._row_of.pop(id_)] = False
            self._append(id_, e[i], None if metadatas is None else metadatas[i],
                         None if documents is None else documents[i])

    def delete(self, ids=None, where=None, where_document=None):
        if (ids is None or len(ids) == 0) and not where and not where_document:
            # chromadb 0.4.22 refuses a delete without ids or a clause instead of wiping the collection
            raise ValueError("You must provide either ids, where, or where_document to delete.")
        rows = self._select_rows(ids, where, where_document)
        for r in rows:
            self._alive[r] = False
            self._row_of.pop(self._ids[r], None)

    def count(self):
        return len(self._row_of)

    # -- reads -------------------------------------------------------------
    def _live_rows(self):
        return [r for r, a in enumerate(self._alive) if a]

    def _select_rows(self, ids, where, where_document=None):
        if ids is not None:
            rows = [self._row_of[i] for i in ids if i in self._row_of]
            rows.sort()
        else:
            rows = self._live_rows()
        if where:
            rows = [r for r in rows if where_match(self._meta[r], where)]
        if where_document:
            rows = [r for r in rows if where_document_match(self._doc[r], where_document)]
        return rows

    def get(self, ids=None, where=None, include=("metadatas", "documents"), where_document=None):
        rows = self._select_rows(ids, where, where_document)
        out = {"ids": [self._ids[r] for r in rows], "embeddings": None,
               "metadatas": None, "documents": None}
        if "embeddings" in include:
            out["embeddings"] = [self._vec[r].tolist() for r in rows]
        if "metadatas" in include:
            out["metadatas"] = [self._meta[r] for r in rows]
        if "documents" in include:
            out["documents"] = [self._doc[r] for r in rows]
        return out

    def query(self, query_embeddings, n_results=10, where=None,
              include=("metadatas", "documents", "distances"), where_document=None):
        q = np.asarray(query_embeddings, dtype=np.float32)
        if q.ndim == 1:
            q = q[None]
        if self.dim is not None and q.shape[1] != self.dim:
            raise ValueError(f"Query dimension {q.shape[1]} does not match collection dimensionality {self.dim}")
        if n_results <= 0:
            raise ValueError("n_results must be a positive integer")
        nq = q.shape[0]
        rows = self._select_rows(None, where, where_document)
        res = {"ids": [], "distances": None, "metadatas": None, "documents": None, "embeddings": None}
        for key in ("distances", "metadatas", "documents", "embeddings"):
            if key in include:
                res[key] = []
        if not rows:
            for key, v in res.items():
                if v is not None:
                    res[key] = [[] for _ in range(nq)]
            return res
        X = np.stack([self._vec[r] for r in rows])
        qq = normalize_f32(q) if self.space == "cosine" else q
        top_rows, top_d = topk_exact(qq, X, n_results, self.space)
        for i in range(nq):
            gl = [rows[j] for j in top_rows[i]]
            res["ids"].append([self._ids[r] for r in gl])
            if res["distances"] is not None:
                res["distances"].append([float(x) for x in top_d[i]])
            if res["metadatas"] is not None:
                res["metadatas"].append([self._meta[r] for r in gl])
            if res["documents"] is not None:
                res["documents"].append([self._doc[r] for r in gl])
            if res["embeddings"] is not None:
                res["embeddings"].append([self._vec[r].tolist() for r in gl])
        return res
