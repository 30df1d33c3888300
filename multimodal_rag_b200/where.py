"""Host-side metadata tables and Chroma `where` evaluation.

Chroma 0.4.22 resolves `where` in its sqlite metadata segment into a set of allowed ids
before the vector search (reference call sites: app/utils/embedder.py:543,599 `filter_dict`
-> `collection.query(where=...)`, and :633 `collection.get(where={"doc_id": ...})`).  Here the
metadata lives in dictionary-encoded columns (one int32 code per row per key) so a clause is
a look-up table over the distinct values of a key, applied with one numpy gather; the
result is either a type-code mask (the `{"type": ...}` fast path, evaluated inside the scan
kernel from a 1-byte/row code) or a packed allow bitmap handed to the device.

Grammar: {key: v} | {key: {"$eq|$ne|$gt|$gte|$lt|$lte": v}} | {key: {"$in|$nin": [...]}} |
{"$and"|"$or": [clause, ...]}.  A missing key never matches; values only compare within a
kind (str / number / bool).
"""
from __future__ import annotations

import operator

import numpy as np

TYPE_KEY = "type"
MAX_TYPE_CODES = 62          # device type codes 1..61 name values, 0 = no value, 62 = overflow
TYPE_NONE, TYPE_OVERFLOW = 0, 62

_CMP = {"$eq": operator.eq, "$ne": operator.ne, "$gt": operator.gt, "$gte": operator.ge,
        "$lt": operator.lt, "$lte": operator.le}


def _kind(v) -> int:
    if isinstance(v, bool):
        return 2
    if isinstance(v, (int, float)):
        return 1
    if isinstance(v, str):
        return 0
    raise ValueError(f"metadata values must be str, int, float or bool, got {type(v).__name__}")


class _Missing:
    """the key is absent from this row's metadata (batch ingestion)"""


_MISSING = _Missing()
_EXACT_KIND = {str: 0, int: 1, float: 1, bool: 2}


class _Column:
    """Dictionary-encoded column: codes[row] = index into values, -1 = key absent."""

    def __init__(self, nrows: int):
        self.codes = np.full(max(nrows, 16), -1, dtype=np.int32)
        self.n = nrows
        self.values: list = []
        self.kinds: list[int] = []
        self.by_kind: tuple = ({}, {}, {})   # raw value -> code, per kind (1 == 1.0 == True as dict keys: kinds keep them apart)
        self.kinds_np = None       # cache of `kinds` as an array (compiled $ne leaves)

    def _grow(self, n):
        if n > self.codes.shape[0]:
            new = np.full(max(n, 2 * self.codes.shape[0]), -1, dtype=np.int32)
            new[: self.n] = self.codes[: self.n]
            self.codes = new

    def set(self, row: int, v):
        self._grow(row + 1)
        self.n = max(self.n, row + 1)
        kd = _kind(v)
        c = self.by_kind[kd].get(v)
        if c is None:
            c = self._new_code(kd, v)
        self.codes[row] = c

    def code_of(self, v):
        """the code of value v in this column's dictionary, or None"""
        return self.by_kind[_kind(v)].get(v)

    def _new_code(self, kd: int, v) -> int:
        c = len(self.values)
        self.by_kind[kd][v] = c
        self.values.append(v)
        self.kinds.append(kd)
        return c

    def set_batch(self, first: int, vals: list):
        """rows first .. first + len(vals) - 1 at once (vals[i] = the row's value or _MISSING).  A batch whose values are of one
        kind -- the usual case -- costs one dict probe per row at C speed; only values never seen before (and absent keys) take
        a second, interpreted pass.  Same codes, in the same order, as `set` row by row."""
        n = len(vals)
        types = set(map(type, vals))
        types.discard(_Missing)
        if not types:
            return                                        # the key is absent from the whole batch
        kds = {_EXACT_KIND.get(t, -1) for t in types}
        self._grow(first + n)
        if len(kds) != 1 or -1 in kds:                    # mixed kinds / subclasses (numpy scalars ...): row by row
            for i, v in enumerate(vals):
                if v is not _MISSING:
                    self.set(first + i, v)
            return
        kd = kds.pop()
        d = self.by_kind[kd]
        get = d.get
        codes = np.asarray([get(v, -2) for v in vals], dtype=np.int32)      # -2: not seen before, or absent
        todo = np.flatnonzero(codes == -2)
        if todo.size:
            # the distinct unseen values, in order of first appearance, take the next codes (what `set` row by row would give);
            # then the unseen rows are probed once more
            unseen = vals if todo.size == n else [vals[i] for i in todo.tolist()]
            fresh = dict.fromkeys(unseen)
            fresh.pop(_MISSING, None)
            base = len(self.values)
            d.update(zip(fresh, range(base, base + len(fresh))))
            self.values.extend(fresh)
            self.kinds.extend([kd] * len(fresh))
            codes[todo] = [get(v, -1) for v in unseen]
        last = n - 1
        while last >= 0 and codes[last] < 0:
            last -= 1
        self.codes[first: first + n] = codes
        self.n = max(self.n, first + last + 1)

    def lut_mask(self, nrows: int, pred) -> np.ndarray:
        lut = np.zeros(len(self.values) + 1, dtype=bool)       # last slot = absent key
        for c, (v, kd) in enumerate(zip(self.values, self.kinds)):
            lut[c] = bool(pred(v, kd))
        codes = self.codes[: self.n]
        out = np.zeros(nrows, dtype=bool)
        out[: self.n] = lut[codes]                               # -1 -> last slot (False)
        return out


MAX_DEVICE_COLUMNS = 16      # include/b2r.h B2R_MAX_COLUMNS
MAX_WHERE_NODES = 32         # include/b2r.h B2R_WHERE_MAX_NODES
LEAF, AND, OR = 0, 1, 2


class MetaTable:
    """Per-collection metadata store (row-aligned with the device corpus)."""

    def __init__(self):
        self.nrows = 0
        self.cols: dict[str, _Column] = {}        # insertion order = device column number (the first 16 keys)
        self.meta: list[dict | None] = []       # original dicts, returned verbatim by query/get
        self.type_codes: dict[str, int] = {}    # value of the `type` key -> device code
        self.type_overflow = False              # more distinct type values than device codes

    def type_code_of(self, meta: dict | None) -> int:
        if not meta or TYPE_KEY not in meta or not isinstance(meta[TYPE_KEY], str):
            return TYPE_NONE
        v = meta[TYPE_KEY]
        c = self.type_codes.get(v)
        if c is None:
            if len(self.type_codes) + 1 < TYPE_OVERFLOW:
                c = self.type_codes[v] = len(self.type_codes) + 1
            else:
                c, self.type_overflow = TYPE_OVERFLOW, True
        return c

    @staticmethod
    def validate(meta):
        if meta is None:
            return
        if not isinstance(meta, dict):
            raise ValueError(f"Expected metadata to be a dict, got {type(meta).__name__}")
        for k, v in meta.items():
            if not isinstance(k, str):
                raise ValueError("Expected metadata key to be a str")
            _kind(v)

    def append(self, meta: dict | None) -> int:
        """Adds one row; returns its device type code."""
        row = self.nrows
        self.nrows += 1
        self.meta.append(meta)
        if meta:
            for k, v in meta.items():
                col = self.cols.get(k)
                if col is None:
                    col = self.cols[k] = _Column(0)
                col.set(row, v)
        return self.type_code_of(meta)

    @staticmethod
    def validate_batch(metas) -> None:
        """`validate` for a whole batch: the types of all keys and values are collected at C speed and only looked at once"""
        key_types, val_types = set(), set()
        for md in metas:
            if md is None:
                continue
            if type(md) is not dict:
                MetaTable.validate(md)                    # subclasses pass, anything else raises with the row's message
            key_types.update(map(type, md))
            val_types.update(map(type, md.values()))
        if key_types - {str} or val_types - set(_EXACT_KIND):
            for md in metas:                              # an unusual type somewhere: the per-row check decides (and words the error)
                MetaTable.validate(md)

    def append_batch(self, metas: list) -> None:
        """`append` for a whole batch, column by column (same tables as row by row; type codes are NOT returned: see
        type_code_of)."""
        first, n = self.nrows, len(metas)
        self.nrows += n
        self.meta.extend(metas)
        # the distinct key tuples of the batch, in order of first appearance (a handful, usually one), give the keys in the order
        # a row-by-row pass would meet them -- that order numbers the device columns
        shapes = dict.fromkeys(map(tuple, filter(None, metas)))
        keys = dict.fromkeys(k for shape in shapes for k in shape)
        if len(shapes) == 1 and None not in metas and all(metas):
            columns = dict(zip(keys, zip(*map(dict.values, metas))))       # every row has the same keys in the same order: transpose
        else:
            columns = {k: [md.get(k, _MISSING) if md else _MISSING for md in metas] for k in keys}
        for k in keys:
            col = self.cols.get(k)
            if col is None:
                col = self.cols[k] = _Column(0)
            col.set_batch(first, columns[k])

    def append_none(self, n: int) -> None:
        """n rows without metadata (bulk adds of bare vectors): no per-row work"""
        self.nrows += n
        self.meta.extend([None] * n)

    def clear(self):
        self.__init__()

    # ---- where ----------------------------------------------------------------
    def mask(self, where: dict | None) -> np.ndarray | None:
        """bool[nrows] of rows matching `where` (None = no restriction)."""
        if not where:
            return None
        return self._eval(where)

    def _eval(self, where) -> np.ndarray:
        if not isinstance(where, dict) or len(where) != 1:
            raise ValueError(f"Expected where to have exactly one operator, got {where}")
        (key, cond), = where.items()
        if key in ("$and", "$or"):
            if not isinstance(cond, (list, tuple)) or len(cond) < 1:
                raise ValueError(f"Expected where value for {key} to be a non-empty list")
            parts = [self._eval(w) for w in cond]
            out = parts[0].copy()
            for p in parts[1:]:
                out = (out & p) if key == "$and" else (out | p)
            return out
        if key.startswith("$"):
            raise ValueError(f"Expected where operator to be one of $and, $or, got {key}")
        if not isinstance(cond, dict):
            cond = {"$eq": cond}
        if len(cond) != 1:
            raise ValueError(f"Expected operator expression to have exactly one operator, got {cond}")
        (op, val), = cond.items()
        col = self.cols.get(key)
        if op in ("$in", "$nin"):
            if not isinstance(val, (list, tuple)) or not val:
                raise ValueError(f"Expected where value for {op} to be a non-empty list")
            wanted = {(_kind(v), v) for v in val}
            if col is None:
                return np.zeros(self.nrows, dtype=bool)
            hit = lambda v, kd: (kd, v) in wanted
            pred = hit if op == "$in" else (lambda v, kd: not hit(v, kd))
            return col.lut_mask(self.nrows, pred)
        if op not in _CMP:
            raise ValueError(f"Expected where operator to be one of {sorted(_CMP)} or $in/$nin, got {op}")
        vk = _kind(val)
        if op in ("$gt", "$gte", "$lt", "$lte") and vk != 1:
            raise ValueError(f"Expected operand value to be an int or a float for operator {op}")
        if col is None:
            return np.zeros(self.nrows, dtype=bool)
        fn = _CMP[op]
        return col.lut_mask(self.nrows, lambda v, kd: kd == vk and fn(v, val))

    # ---- where, compiled for the device (include/b2r.h b2r_where) --------------------------------
    def device_column(self, key: str) -> int | None:
        """Device column number of `key` (its position among the first 16 keys seen), else None."""
        for i, k in enumerate(self.cols):
            if i >= MAX_DEVICE_COLUMNS:
                return None
            if k == key:
                return i
        return None

    def _leaf_pred(self, key, cond):
        """(column or None, predicate over (value, kind)) of one `{key: cond}` leaf -- the same rules as _eval."""
        if not isinstance(cond, dict):
            cond = {"$eq": cond}
        if len(cond) != 1:
            raise ValueError(f"Expected operator expression to have exactly one operator, got {cond}")
        (op, val), = cond.items()
        col = self.cols.get(key)
        if op in ("$in", "$nin"):
            if not isinstance(val, (list, tuple)) or not val:
                raise ValueError(f"Expected where value for {op} to be a non-empty list")
            wanted = {(_kind(v), v) for v in val}
            hit = lambda v, kd: (kd, v) in wanted
            return col, (hit if op == "$in" else (lambda v, kd: not hit(v, kd)))
        if op not in _CMP:
            raise ValueError(f"Expected where operator to be one of {sorted(_CMP)} or $in/$nin, got {op}")
        vk = _kind(val)
        if op in ("$gt", "$gte", "$lt", "$lte") and vk != 1:
            raise ValueError(f"Expected operand value to be an int or a float for operator {op}")
        fn = _CMP[op]
        return col, (lambda v, kd: kd == vk and fn(v, val))

    @staticmethod
    def _leaf_table(col: _Column, cond, pred) -> np.ndarray:
        """bool per distinct value of `col` (padded to a multiple of 32): does the leaf's comparison hold for it?
        Equality-style operators go through the column's value index (O(operands), not O(distinct values) --
        a doc_id column has one value per document); ranges test every distinct value."""
        nv = len(col.values)
        bits = np.zeros(((nv + 31) // 32) * 32, dtype=bool)
        if not isinstance(cond, dict):
            cond = {"$eq": cond}
        (op, val), = cond.items()
        if op in ("$eq", "$in", "$nin"):
            for v in ([val] if op == "$eq" else val):
                c = col.code_of(v)
                if c is not None:
                    bits[c] = True
            if op == "$nin":
                bits[:nv] = ~bits[:nv]
        elif op == "$ne":
            if col.kinds_np is None or col.kinds_np.shape[0] != nv:
                col.kinds_np = np.asarray(col.kinds, dtype=np.int8)
            bits[:nv] = col.kinds_np == _kind(val)
            c = col.code_of(val)
            if c is not None:
                bits[c] = False
        else:
            for c, (v, kd) in enumerate(zip(col.values, col.kinds)):
                bits[c] = bool(pred(v, kd))
        return bits

    def compile(self, where: dict | None):
        """`where` -> (nodes, lut): the clause in postfix order as (op, column, lut_offset, lut_values) tuples and
        the concatenated leaf look-up tables (uint32 words; bit c of a leaf's table = its comparison holds for the
        column's c-th distinct value).  None when the clause cannot run on the device (a key beyond the 16 device
        columns, more than 32 nodes); the caller then hands the device a host-evaluated bitmap instead."""
        if not where:
            return None
        nodes, luts = [], []

        def walk(w):
            if not isinstance(w, dict) or len(w) != 1:
                raise ValueError(f"Expected where to have exactly one operator, got {w}")
            (key, cond), = w.items()
            if key in ("$and", "$or"):
                if not isinstance(cond, (list, tuple)) or len(cond) < 1:
                    raise ValueError(f"Expected where value for {key} to be a non-empty list")
                ok = walk(cond[0])
                for sub in cond[1:]:
                    ok = walk(sub) and ok
                    nodes.append((AND if key == "$and" else OR, 0, 0, 0))
                return ok
            if key.startswith("$"):
                raise ValueError(f"Expected where operator to be one of $and, $or, got {key}")
            col, pred = self._leaf_pred(key, cond)
            off = sum(len(x) for x in luts)
            if col is None:                       # no row carries the key: an empty table on column 0 is always false
                nodes.append((LEAF, 0, off, 0))
                return True
            ci = self.device_column(key)
            if ci is None:
                return False
            luts.append(np.packbits(self._leaf_table(col, cond, pred), bitorder="little").view("<u4"))
            nodes.append((LEAF, ci, off, len(col.values)))
            return True

        if not walk(where) or len(nodes) > MAX_WHERE_NODES:
            return None
        lut = np.concatenate(luts).astype(np.uint32) if luts else np.zeros(0, dtype=np.uint32)
        return nodes, np.ascontiguousarray(lut)

    def type_only_mask(self, where: dict | None) -> int | None:
        """If `where` only constrains the `type` key by equality / $in on string values that all
        have a device code, return the 64-bit type mask; else None (caller uses a bitmap)."""
        if not where or len(where) != 1 or TYPE_KEY not in where:
            return None
        if self.type_overflow:
            return None
        cond = where[TYPE_KEY]
        if not isinstance(cond, dict):
            cond = {"$eq": cond}
        if len(cond) != 1:
            return None
        (op, val), = cond.items()
        vals = [val] if op == "$eq" else list(val) if op == "$in" and isinstance(val, (list, tuple)) else None
        if vals is None or not all(isinstance(v, str) for v in vals):
            return None
        m = 0
        for v in vals:
            c = self.type_codes.get(v)
            if c is not None:
                m |= 1 << c
        return m


def doc_mask(docs: list, where_document: dict) -> np.ndarray:
    """bool[len(docs)] of rows whose DOCUMENT satisfies a Chroma ``where_document`` clause (chromadb 0.4.22:
    ``{"$contains": s}``, ``{"$not_contains": s}``, ``{"$and": [...]}``, ``{"$or": [...]}``).  Chroma answers these from its
    sqlite full-text table before the vector search; here it is a pass over the host's document list that ends as an allow
    bitmap for the device (the reference never passes where_document -- app/utils/embedder.py:595-601 -- so this is a
    compatibility path, not a hot one).  A row without a document matches neither operator."""
    if not isinstance(where_document, dict) or len(where_document) != 1:
        raise ValueError(f"Expected where document to have exactly one operator, got {where_document}")
    (op, val), = where_document.items()
    if op in ("$and", "$or"):
        if not isinstance(val, (list, tuple)) or len(val) < 1:
            raise ValueError(f"Expected where document value for {op} to be a non-empty list")
        parts = [doc_mask(docs, w) for w in val]
        out = parts[0]
        for p in parts[1:]:
            out = (out & p) if op == "$and" else (out | p)
        return out
    if op not in ("$contains", "$not_contains"):
        raise ValueError(f"Expected where document operator to be one of $contains, $not_contains, $and, $or, got {op}")
    if not isinstance(val, str) or not val:
        raise ValueError(f"Expected where document operand value for operator {op} to be a non-empty str")
    has = np.fromiter((d is not None and val in d for d in docs), dtype=bool, count=len(docs))
    if op == "$contains":
        return has
    return ~has & np.fromiter((d is not None for d in docs), dtype=bool, count=len(docs))


def pack_bits(mask: np.ndarray) -> np.ndarray:
    """bool[n] -> uint32[ceil(n/32)], bit (r & 31) of word r >> 5 (include/b2r.h b2r_filter)."""
    n = mask.shape[0]
    padded = np.zeros(((n + 31) // 32) * 32, dtype=bool)
    padded[:n] = mask
    return np.packbits(padded, bitorder="little").view("<u4").copy()
