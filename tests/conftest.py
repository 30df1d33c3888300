import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    """The reference's committed Chroma WAL: 70 real MiniLM vectors + ids + metadata + known answers."""
    here = os.path.join(ROOT, "tests", "golden")
    meta = json.load(open(os.path.join(here, "chroma_wal.json")))
    meta["vectors"] = np.load(os.path.join(here, "chroma_wal.npz"))["vectors"]
    return meta


def make_unit(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)
