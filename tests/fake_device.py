"""TEST INFRASTRUCTURE: a numpy stand-in for the device half of libb2r.so, so that the HOST logic of B200Collection (id table,
tombstones, metadata columns, filters, result assembly, error behaviour) runs in the CPU suite.  It answers the handful of
entry points the collection calls -- create / ingest / tombstone / column_set / filter_eval / query_ex / get_rows / stats -- with the
oracle's arithmetic and passes everything else (the native id table, b2r_last_error) to the real library.  Nothing in the product
imports this; the GPU suite runs the same scenarios against the real kernels."""
from __future__ import annotations

import ctypes

import numpy as np

from multimodal_rag_b200 import _lib
from oracle import exact_oracle as eo

_SPACE = {v: k for k, v in _lib.SPACE_CODE.items()}


def _arr(ptr, ctype, shape):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=np.dtype(ctype))
    return np.ctypeslib.as_array(ctypes.cast(int(ptr), ctypes.POINTER(ctype)), shape=(n,)).reshape(shape)


class _Shard:
    def __init__(self, dim, space):
        self.dim, self.space = dim, space
        self.X = np.zeros((0, dim), dtype=np.float32)
        self.type_code = np.zeros(0, dtype=np.uint8)
        self.cols: dict[int, np.ndarray] = {}
        self.launches = 0

    @property
    def rows(self):
        return self.X.shape[0]

    def passing(self, f) -> np.ndarray:
        """bool per row: what the device derives from a b2r_filter (type mask, allow bitmap, compiled clause; dead rows never)."""
        n = self.rows
        tm = int(f.type_mask) & ~(1 << _lib.TYPE_DEAD) if f is not None else ~(1 << _lib.TYPE_DEAD)
        ok = np.asarray([(tm >> c) & 1 for c in range(64)], dtype=bool)[self.type_code]
        if f is not None and f.allow_bits:
            words = _arr(f.allow_bits, ctypes.c_uint32, ((n + 31) // 32,))
            ok &= np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)
        if f is not None and bool(f.where):
            w = f.where.contents
            lut = _arr(w.lut, ctypes.c_uint32, (int(w.lut_words),)) if w.lut_words else np.zeros(0, dtype=np.uint32)
            stack = []
            for i in range(w.n_nodes):              # the postfix program where_bits_kernel runs
                nd = w.nodes[i]
                if nd.op == _lib.WHERE_LEAF:
                    bit = np.zeros(n, dtype=bool)
                    if nd.lut_values:
                        codes = np.full(n, -1, dtype=np.int64)
                        have = self.cols.get(nd.column)
                        if have is not None:
                            codes[: have.shape[0]] = have[:n]
                        good = (codes >= 0) & (codes < nd.lut_values)
                        c = np.where(good, codes, 0)
                        bit = good & (((lut[nd.lut_offset + (c >> 5)] >> (c & 31).astype(np.uint32)) & 1) == 1)
                    stack.append(bit)
                else:
                    b, a = stack.pop(), stack.pop()
                    stack.append(a & b if nd.op == _lib.WHERE_AND else a | b)
            ok &= stack[0]
        return ok


class FakeLib:
    """Quacks like the ctypes library object for the device calls; everything else is the real library."""

    def __init__(self):
        self.real = _lib.load()
        self.shards: dict[int, _Shard] = {}
        self.next = 1

    def __getattr__(self, name):
        if name.startswith("b2r_idtab_") or name in ("b2r_last_error", "b2r_abi_version"):      # host-only entry points: the real ones
            return getattr(self.real, name)
        raise AttributeError(f"the fake device does not answer {name}")

    def _s(self, h) -> _Shard:
        return self.shards[int(h.value if hasattr(h, "value") else h)]

    def b2r_create(self, dim, space, capacity, device, flags, h_ref):
        self.shards[self.next] = _Shard(int(dim), _SPACE[int(space)])
        h_ref._obj.value = self.next
        self.next += 1
        return _lib.B2R_OK

    def b2r_destroy(self, h):
        self.shards.pop(int(h.value), None)
        return _lib.B2R_OK

    def b2r_count(self, h):
        return int((self._s(h).type_code != _lib.TYPE_DEAD).sum())

    def b2r_launch_count(self, h):
        return self._s(h).launches

    def b2r_get_stats(self, h, st_ref):
        s, st = self._s(h), st_ref._obj
        st.dim, st.rows, st.live = s.dim, s.rows, self.b2r_count(h)
        return _lib.B2R_OK

    def b2r_ingest_f32(self, h, x_ptr, n, codes_ptr, first_ref, stream):
        s = self._s(h)
        x = _arr(x_ptr, ctypes.c_float, (int(n), s.dim)).copy()
        first_ref._obj.value = s.rows
        s.X = np.concatenate([s.X, eo.normalize_f32(x) if s.space == "cosine" else x])
        codes = _arr(codes_ptr, ctypes.c_uint8, (int(n),)).copy() if codes_ptr else np.zeros(int(n), dtype=np.uint8)
        s.type_code = np.concatenate([s.type_code, codes])
        s.launches += 1
        return _lib.B2R_OK

    def b2r_tombstone(self, h, rows_ptr, n, stream):
        s = self._s(h)
        s.type_code[_arr(rows_ptr, ctypes.c_int64, (int(n),))] = _lib.TYPE_DEAD
        return _lib.B2R_OK

    def b2r_get_rows_f32(self, h, rows_ptr, n, out_ptr, stream):
        s = self._s(h)
        _arr(out_ptr, ctypes.c_float, (int(n), s.dim))[:] = s.X[_arr(rows_ptr, ctypes.c_int64, (int(n),))]
        return _lib.B2R_OK

    def b2r_column_set(self, h, column, first, n, codes_ptr, stream):
        s = self._s(h)
        col = s.cols.get(int(column))
        need = int(first) + int(n)
        if col is None or col.shape[0] < need:
            grown = np.full(max(need, 2 * (0 if col is None else col.shape[0])), -1, dtype=np.int32)
            if col is not None:
                grown[: col.shape[0]] = col
            col = s.cols[int(column)] = grown
        col[int(first): need] = _arr(codes_ptr, ctypes.c_int32, (int(n),))
        return _lib.B2R_OK

    def b2r_filter_eval(self, h, f_ref, out_ptr, stream):
        s = self._s(h)
        ok = s.passing(f_ref._obj if f_ref is not None else None)
        padded = np.zeros(((s.rows + 31) // 32) * 32, dtype=bool)
        padded[: s.rows] = ok
        _arr(out_ptr, ctypes.c_uint32, ((s.rows + 31) // 32,))[:] = np.packbits(padded, bitorder="little").view("<u4")
        return _lib.B2R_OK

    def b2r_query_ex(self, h, q_ptr, nq, k, f_ref, rows_ptr, dist_ptr, d64_ptr, cnt_ptr, stream):
        s = self._s(h)
        nq, k = int(nq), int(k)
        q = _arr(q_ptr, ctypes.c_float, (nq, s.dim)).copy()
        ok = s.passing(f_ref._obj if f_ref is not None else None)
        rows = _arr(rows_ptr, ctypes.c_int64, (nq, k)); rows[:] = -1
        dist = _arr(dist_ptr, ctypes.c_float, (nq, k)); dist[:] = np.inf
        cnt = _arr(cnt_ptr, ctypes.c_int32, (nq,)); cnt[:] = 0
        d64 = _arr(d64_ptr, ctypes.c_double, (nq, k)) if d64_ptr else None
        if d64 is not None:
            d64[:] = np.inf
        if ok.any():
            qq = eo.normalize_f32(q) if s.space == "cosine" else q
            dd = eo.distances_f64(qq, s.X, s.space)
            cand = np.flatnonzero(ok)
            for i in range(nq):
                order = cand[np.lexsort((cand, dd[i, cand]))[:k]]
                cnt[i] = len(order)
                rows[i, : len(order)] = order
                dist[i, : len(order)] = dd[i, order].astype(np.float32)
                if d64 is not None:
                    d64[i, : len(order)] = dd[i, order]
        s.launches += 1
        return _lib.B2R_OK
