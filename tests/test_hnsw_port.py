"""The HNSW restatement used as the CPU baseline (oracle/hnsw_port.c): it must behave like an HNSW index --
exact on tiny collections, high recall on structured data, results ascending -- so its speed/recall numbers mean
what bench.py says they mean."""
import numpy as np

from conftest import make_unit


def test_hnsw_is_exact_on_the_golden_fixture(golden):
    from oracle import c_oracle, exact_oracle as eo
    X = eo.normalize_f32(golden["vectors"])
    idx = c_oracle.Hnsw(X[1:], "cosine", build_threads=1)
    rows, dist = idx.query(X[:1], 5, ef=100)
    ids = [golden["ids"][1:][r] for r in rows[0]]
    assert ids == [a["id"] for a in golden["known"]["top5"]]
    np.testing.assert_allclose(dist[0], [a["cosine"] for a in golden["known"]["top5"]], rtol=1e-4)


def test_hnsw_recall_on_structured_data_and_ordering():
    from oracle import c_oracle
    rng = np.random.default_rng(0)
    A = rng.standard_normal((8, 128)).astype(np.float32)
    X = rng.standard_normal((20000, 8)).astype(np.float32) @ A
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    Q = rng.standard_normal((200, 8)).astype(np.float32) @ A
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    idx = c_oracle.Hnsw(X, "cosine", build_threads=1)       # single-threaded build: deterministic graph
    exact, _, _ = c_oracle.topk(X, Q, 10, "cosine")
    for ef, floor in ((10, 0.80), (100, 0.98)):
        rows, dist = idx.query(Q, 10, ef)
        rec = np.mean([len(set(rows[i]) & set(exact[i])) / 10 for i in range(len(Q))])
        assert rec >= floor, (ef, rec)
        assert (np.diff(dist, axis=1) >= 0).all()
    # l2 space, odd dimension, k > ef floor
    Y = make_unit(3000, 50, 3) * 2.0
    idx2 = c_oracle.Hnsw(Y, "l2", build_threads=1)
    rows, dist = idx2.query(Y[:20], 3, ef=200)
    hit = rows[:, 0] == np.arange(20)             # approximate index, isotropic data: allow a miss
    assert hit.sum() >= 16 and (dist[hit, 0] < 1e-6).all()
