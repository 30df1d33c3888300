"""CPU: the C-ABI library loads here (no GPU), exports every symbol include/b2r.h declares, and
fails loudly -- never falls back -- when asked to compute without a device."""
import ctypes
import os
import re
import subprocess

import pytest

from multimodal_rag_b200 import _lib, build as b2r_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    b2r_build.build()
    return _lib.load()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "b2r.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(lib):
    decl = _declared()
    assert sorted(_lib.ABI_SYMBOLS) == decl
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(decl) <= exported
    assert {s for s in exported if s.startswith("b2r_")} == set(decl)       # nothing undocumented
    assert lib.b2r_abi_version() == 4


def test_library_is_sm100a_native(lib):
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.b2r_create(384, 1, 0, 0, 0, ctypes.byref(h))
    assert rc == _lib.B2R_ECUDA and b"no CPU fallback" in lib.b2r_last_error()
    from multimodal_rag_b200 import B200Collection
    c = B200Collection("c", {"hnsw:space": "cosine"})
    with pytest.raises(RuntimeError):
        c.add(ids=["a"], embeddings=[[0.0] * 384])
    with pytest.raises(ValueError):
        B200Collection("c", {"hnsw:space": "hamming"})


def test_argument_validation_without_device(lib):
    h = ctypes.c_void_p()
    assert lib.b2r_create(0, 1, 0, 0, 0, ctypes.byref(h)) == _lib.B2R_EINVAL
    assert lib.b2r_create(384, 7, 0, 0, 0, ctypes.byref(h)) == _lib.B2R_EINVAL
    assert lib.b2r_query(None, None, 1, 5, None, None, None, None, None) == _lib.B2R_EINVAL
    assert lib.b2r_count(None) == -1
    with pytest.raises(ValueError):
        _lib.check(lib.b2r_create(384, 7, 0, 0, 0, ctypes.byref(h)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "multimodal_rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("the oracle's tie-break key", ""), f"{fn} mentions the oracle"


def test_load_rejects_non_shard_files_without_a_gpu(tmp_path):
    """b2r_load validates the file before it touches the device: a missing file or a foreign file is EINVAL
    (-> ValueError upstream), whatever the machine."""
    import ctypes
    from multimodal_rag_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.b2r_load(str(tmp_path / "nope.b2r").encode(), 0, 0, ctypes.byref(h)) == _lib.B2R_EINVAL
    assert b"cannot open" in lib.b2r_last_error()
    bad = tmp_path / "bad.b2r"
    bad.write_bytes(b"SQLite format 3\x00" + bytes(200))
    assert lib.b2r_load(str(bad).encode(), 0, 0, ctypes.byref(h)) == _lib.B2R_EINVAL
    assert b"not a b2r shard file" in lib.b2r_last_error()
    with pytest.raises(ValueError):
        _lib.check(lib.b2r_load(str(bad).encode(), 0, 0, ctypes.byref(h)))


def test_load_validates_the_header_before_touching_the_device(tmp_path):
    """Shard file header (include/b2r.h): 128 bytes = "B2RS", version, dim, padded dim, space, flags,
    rows, live, row base, two norms, checksum, payload bytes.  Every inconsistency is EINVAL, with or without a GPU."""
    import ctypes
    import struct
    from multimodal_rag_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()

    def header(version=1, dim=384, dp=384, space=1, flags=0, rows=10, live=10, payload=None):
        if payload is None:
            payload = max(0, rows * dp * 2 + rows * dp * 4 + rows)    # corpus + master + type codes (cosine: no bias)
        b = b"B2RS" + struct.pack("<IiiiIqqq", version, dim, dp, space, flags, rows, live, 0) + struct.pack("<ff", 1.0, 1e-6)
        b += struct.pack("<QQ", 0, payload)
        return b + bytes(128 - len(b))

    def load(blob):
        p = tmp_path / "x.b2r"
        p.write_bytes(blob)
        rc = lib.b2r_load(str(p).encode(), 0, 0, ctypes.byref(h))
        return rc, lib.b2r_last_error()

    assert len(header()) == 128
    assert load(header(version=7)) == (_lib.B2R_EINVAL, b"b2r_load: unknown shard file version")
    for bad in (header(dim=0), header(dp=400), header(space=5), header(rows=-1), header(rows=3, live=4)):
        rc, msg = load(bad)
        assert rc == _lib.B2R_EINVAL and b"corrupt header" in msg
    rc, msg = load(header(payload=12345))
    assert rc == _lib.B2R_EINVAL and b"payload size" in msg
