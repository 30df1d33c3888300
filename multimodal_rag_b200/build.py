"""Builds libb2r.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

    python -m multimodal_rag_b200.build [--force] [--verbose]

Kernel families are separate translation units compiled in parallel; the scan kernel is
compiled once per supported padded dimension (-DB2R_DP=...).  nvcc cross-compiles here
without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb2r.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O2,-Wall", "-ccbin", "/usr/bin/g++"]
SCAN_DPS = [64, 128, 256, 384, 512, 768, 1024]
GEMM_KBS = [2, 4, 6, 8, 12, 16, 24]  # padded dim / 64: 128, 256, 384, 512, 768, 1024, 1536


def _units():
    units = [("b2r_api", "b2r_api.cu", []), ("exact_kernels", "exact_kernels.cu", []), ("xchg", "xchg.cu", []), ("idtable", "idtable.cu", []),
             ("scan_dispatch", "scan_kernels.cu", [])]
    units += [(f"scan_dp{dp}", "scan_kernels.cu", [f"-DB2R_DP={dp}"]) for dp in SCAN_DPS]
    units.append(("gemm_dispatch", "gemm_kernels.cu", []))
    units += [(f"gemm_kb{kb}", "gemm_kernels.cu", [f"-DB2R_KB={kb}"]) for kb in GEMM_KBS]
    return units


def _src_digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for fn in sorted(os.listdir(root)):
            if fn.endswith((".cu", ".cuh", ".h")):
                h.update(fn.encode())
                with open(os.path.join(root, fn), "rb") as f:
                    h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(name, src, defs, verbose):
    out = os.path.join(OBJ, name + ".o")
    cmd = [NVCC, *FLAGS, *defs, "-c", os.path.join(CSRC, src), "-o", out]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    return out, r.stderr


def build_asan() -> str:
    """The same library with its HOST code instrumented by AddressSanitizer + UBSan (device code is untouched): the CPU tests that
    drive the host layer without a GPU -- the id table, argument validation, shard-file parsing -- run against it with
        LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 B2R_LIB=multimodal_rag_b200/build_asan/libb2r_asan.so \
            python -m pytest tests/test_abi.py tests/test_idtable.py tests/test_host_tables.py -q
    (compute-sanitizer is closed on the GPU pool; this covers the half of the library that is plain C++)."""
    out_dir = os.path.join(HERE, "build_asan")
    os.makedirs(out_dir, exist_ok=True)
    san = "-fsanitize=address,-fsanitize=undefined"          # (-Xcompiler splits at commas)
    flags = ARCH + ["-O1", "-g", "-std=c++17", "-lineinfo", "-Xcompiler", f"-fPIC,-fno-omit-frame-pointer,{san}", "-ccbin", "/usr/bin/g++"]

    def one(u):
        name, src, defs = u
        out = os.path.join(out_dir, name + ".o")
        r = subprocess.run([NVCC, *flags, *defs, "-c", os.path.join(CSRC, src), "-o", out], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
        return out

    units = _units()
    with cf.ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(one, units))
    lib = os.path.join(out_dir, "libb2r_asan.so")
    r = subprocess.run([NVCC, *ARCH, "-shared", "-o", lib, *objs, "-ccbin", "/usr/bin/g++", "-Xcompiler", san], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _src_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    units = _units()
    objs = []
    with cf.ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        for out, log in ex.map(lambda u: _compile(*u, verbose), units):
            objs.append(out)
            if verbose and log:
                sys.stderr.write(log)
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--asan", action="store_true", help="build build_asan/libb2r_asan.so (host code under ASan + UBSan) instead")
    a = ap.parse_args()
    print(build_asan() if a.asan else build(a.force, a.verbose))
