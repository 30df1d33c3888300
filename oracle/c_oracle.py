"""ctypes binding for oracle/exact_topk.c -- TEST INFRASTRUCTURE ONLY (see exact_oracle.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libb2r_oracle.so")
SPACE_CODE = {"l2": 0, "cosine": 1, "ip": 2}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "exact_topk.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libb2r_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.b2r_oracle_threads.restype = ctypes.c_int
        _lib.b2r_oracle_normalize_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
        _lib.b2r_oracle_normalize_f32.restype = None
        _lib.b2r_oracle_topk.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.b2r_oracle_topk.restype = ctypes.c_int
    return _lib


def threads() -> int:
    return int(lib().b2r_oracle_threads())


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
