"""Host id table of a collection (string id <-> dense device row) behind the C ABI's ``b2r_idtab_*`` entry points.

It stands where Chroma keeps its id index (the ``embeddings.embedding_id`` column consulted by ``collection.add`` /
``upsert`` / ``get(ids)`` / ``delete(ids)``: app/utils/embedder.py:518, 632, 640, 888).  A batch of Python strings crosses
the boundary as ONE buffer (``"\\0".join(ids)``) plus one offsets array, so an 8192-id upsert costs a join, an encode and
three C calls instead of 8192 dict probes that each miss the cache on a 10M-id table.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


class EncodedIds:
    """A batch of ids as the C ABI takes it: `buf` (bytes), `off` int64[n + 1], `gap` separator bytes after every id."""

    __slots__ = ("n", "buf", "off", "gap")

    def __init__(self, n, buf, off, gap):
        self.n, self.buf, self.off, self.gap = n, buf, off, gap

    def has_empty(self) -> bool:
        return bool(self.n) and bool((np.diff(self.off) == self.gap).any())


def encode_ids(ids) -> EncodedIds:
    """list[str] -> EncodedIds.  Raises TypeError when an element is not a str (the caller words the message)."""
    n = len(ids)
    if n == 0:
        return EncodedIds(0, b"", np.zeros(1, dtype=np.int64), 0)
    buf = ("\0".join(ids) + "\0").encode("utf-8", "surrogatepass")
    seps = np.flatnonzero(np.frombuffer(buf, dtype=np.uint8) == 0)
    if seps.shape[0] == n:
        off = np.empty(n + 1, dtype=np.int64)
        off[0] = 0
        off[1:] = seps + 1
        return EncodedIds(n, buf, off, 1)
    # an id holds a NUL byte itself: encode one by one, packed
    parts = [i.encode("utf-8", "surrogatepass") for i in ids]
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.fromiter(map(len, parts), dtype=np.int64, count=n), out=off[1:])
    return EncodedIds(n, b"".join(parts), off, 0)


class IdTable:
    def __init__(self, reserve: int = 0):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self._lib.b2r_idtab_create(int(reserve), ctypes.byref(h)), "b2r_idtab_create")
        self._h = h

    def close(self):
        if self._h is not None:
            self._lib.b2r_idtab_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def live(self) -> int:
        return int(self._lib.b2r_idtab_live(self._h))

    @property
    def rows(self) -> int:
        return int(self._lib.b2r_idtab_rows(self._h))

    def clear(self):
        _lib.check(self._lib.b2r_idtab_clear(self._h), "b2r_idtab_clear")

    def lookup(self, enc: EncodedIds, want_dup: bool = False):
        """-> rows int64[n] (-1 = unknown id) [, index of the first in-batch repeat or -1]"""
        rows = np.empty(enc.n, dtype=np.int64)
        dup = ctypes.c_int64(-1)
        _lib.check(self._lib.b2r_idtab_lookup(self._h, enc.buf, enc.off.ctypes.data, enc.n, enc.gap, rows.ctypes.data,
                                              ctypes.byref(dup) if want_dup else None), "b2r_idtab_lookup")
        return (rows, int(dup.value)) if want_dup else rows

    def append(self, enc: EncodedIds, first_row: int) -> np.ndarray:
        """The batch becomes rows first_row..; -> the rows the ids pointed at before (-1 = new id)."""
        prev = np.empty(enc.n, dtype=np.int64)
        _lib.check(self._lib.b2r_idtab_append(self._h, enc.buf, enc.off.ctypes.data, enc.n, enc.gap, int(first_row),
                                              prev.ctypes.data), "b2r_idtab_append")
        return prev

    def erase_rows(self, rows) -> None:
        arr = np.ascontiguousarray(rows, dtype=np.int64)
        _lib.check(self._lib.b2r_idtab_erase_rows(self._h, arr.ctypes.data, arr.shape[0]), "b2r_idtab_erase_rows")

    def ids_of(self, rows) -> list:
        """list[str]: the ids of `rows` (live or erased)"""
        arr = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1)
        n = arr.shape[0]
        if n == 0:
            return []
        off = np.empty(n + 1, dtype=np.int64)
        need = ctypes.c_int64(0)
        cap = 48 * n + 64
        while True:
            out = np.empty(cap, dtype=np.uint8)
            _lib.check(self._lib.b2r_idtab_ids_of(self._h, arr.ctypes.data, n, out.ctypes.data, cap, off.ctypes.data,
                                                  ctypes.byref(need)), "b2r_idtab_ids_of")
            if need.value <= cap:
                break
            cap = int(need.value)
        raw = out[: need.value - 1].tobytes()
        ids = raw.decode("utf-8", "surrogatepass").split("\0")
        if len(ids) != n:                       # an id holds a NUL byte: cut by the offsets instead
            ids = [raw[off[i]: off[i + 1] - 1].decode("utf-8", "surrogatepass") for i in range(n)]
        return ids

    def id_of(self, row: int) -> str:
        return self.ids_of([row])[0]


class IdsByRow:
    """Read-only, list-like view: row -> id (what the first host mirror kept as a Python list)."""

    def __init__(self, table: IdTable):
        self._t = table

    def __len__(self):
        return self._t.rows

    def __getitem__(self, r):
        if isinstance(r, slice):
            return self._t.ids_of(np.arange(*r.indices(self._t.rows), dtype=np.int64))
        r = int(r)
        n = self._t.rows
        if r < 0:
            r += n
        if not 0 <= r < n:
            raise IndexError("row out of range")
        return self._t.id_of(r)

    def __iter__(self):
        return iter(self[:])


class RowOfId:
    """Read-only, dict-like view: live id -> row."""

    def __init__(self, table: IdTable):
        self._t = table

    def __len__(self):
        return self._t.live

    def get(self, id_, default=None):
        if not isinstance(id_, str):
            return default
        r = int(self._t.lookup(encode_ids([id_]))[0])
        return default if r < 0 else r

    def __contains__(self, id_):
        return self.get(id_) is not None

    def __getitem__(self, id_):
        r = self.get(id_)
        if r is None:
            raise KeyError(id_)
        return r
