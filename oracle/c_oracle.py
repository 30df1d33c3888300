"""ctypes binding for oracle/exact_topk.c -- TEST INFRASTRUCTURE ONLY (see exact_oracle.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("B2R_ORACLE_SO") or os.path.join(_HERE, "libb2r_oracle.so")    # B2R_ORACLE_SO: a sanitizer build (make -C oracle asan)
SPACE_CODE = {"l2": 0, "cosine": 1, "ip": 2}


def build(force: bool = False) -> str:
    if os.environ.get("B2R_ORACLE_SO"):
        return _SO
    srcs = [os.path.join(_HERE, f) for f in ("exact_topk.c", "hnsw_port.c", "Makefile")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libb2r_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.b2r_oracle_threads.restype = ctypes.c_int
        _lib.b2r_oracle_set_threads.argtypes = [ctypes.c_int]
        _lib.b2r_oracle_set_threads.restype = None
        _lib.b2r_oracle_normalize_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
        _lib.b2r_oracle_normalize_f32.restype = None
        _lib.b2r_oracle_topk.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.b2r_oracle_topk.restype = ctypes.c_int
        _lib.b2r_hnsw_threads.restype = ctypes.c_int
        _lib.b2r_hnsw_build.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_uint]
        _lib.b2r_hnsw_build.restype = ctypes.c_void_p
        _lib.b2r_hnsw_search.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p]
        _lib.b2r_hnsw_search.restype = None
        _lib.b2r_hnsw_set_threads.argtypes = [ctypes.c_int]
        _lib.b2r_hnsw_set_threads.restype = None
        _lib.b2r_hnsw_free.argtypes = [ctypes.c_void_p]
        _lib.b2r_hnsw_free.restype = None
    return _lib


def threads() -> int:
    return int(lib().b2r_oracle_threads())


def set_threads(n: int = 0) -> int:
    """Use n OpenMP threads (0 = every core this process may run on); returns the count now in effect."""
    if n <= 0:
        n = len(os.sched_getaffinity(0))
    lib().b2r_oracle_set_threads(int(n))
    return threads()


def normalize_f32(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().b2r_oracle_normalize_f32(x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data)
    return out


def topk(X, Q, k, space, allowed=None, acc64=True):
    """X, Q as stored (normalised for cosine).  Returns rows [nq,k] int64 (-1 pad),
    dists [nq,k] fp32 (+inf pad), counts [nq] int32."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    Q = np.ascontiguousarray(np.atleast_2d(Q), dtype=np.float32)
    nq = Q.shape[0]
    rows = np.empty((nq, k), dtype=np.int64)
    dist = np.empty((nq, k), dtype=np.float32)
    cnt = np.empty((nq,), dtype=np.int32)
    a = None
    if allowed is not None:
        a = np.ascontiguousarray(allowed, dtype=np.uint8)
    rc = lib().b2r_oracle_topk(X.ctypes.data, X.shape[0], X.shape[1], Q.ctypes.data, nq, k,
                               SPACE_CODE[space], None if a is None else a.ctypes.data,
                               1 if acc64 else 0, rows.ctypes.data, dist.ctypes.data, cnt.ctypes.data)
    if rc != 0:
        raise ValueError("b2r_oracle_topk: bad arguments")
    return rows, dist, cnt


class Hnsw:
    """hnsw_port.c: restatement of the chroma-hnswlib index the reference queries (Chroma defaults
    M=16, ef_construction=100, search ef = max(10, k)).  X must be stored rows (normalised for cosine)
    and stay alive while the index is used."""

    def __init__(self, X: np.ndarray, space: str = "cosine", M: int = 16, ef_construction: int = 100, seed: int = 100,
                 build_threads: int = 0):
        self.X = np.ascontiguousarray(X, dtype=np.float32)
        self.space = space
        self._h = None
        if build_threads > 0:                      # 1 = deterministic insertion order
            lib().b2r_hnsw_set_threads(build_threads)
        try:
            self._h = lib().b2r_hnsw_build(self.X.ctypes.data, self.X.shape[0], self.X.shape[1],
                                           0 if space == "l2" else 1, M, ef_construction, seed)
        finally:
            if build_threads > 0:
                lib().b2r_hnsw_set_threads(0)
        if not self._h:
            raise ValueError("b2r_hnsw_build failed")

    def query(self, Q, k: int, ef: int = 10):
        Q = np.ascontiguousarray(np.atleast_2d(Q), dtype=np.float32)
        rows = np.empty((Q.shape[0], k), dtype=np.int64)
        dist = np.empty((Q.shape[0], k), dtype=np.float32)
        lib().b2r_hnsw_search(self._h, Q.ctypes.data, Q.shape[0], k, max(ef, k), rows.ctypes.data, dist.ctypes.data)
        return rows, dist

    def close(self):
        if self._h:
            lib().b2r_hnsw_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
