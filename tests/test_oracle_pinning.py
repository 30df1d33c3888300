"""Pins the oracle against the REAL dependency the moment one is importable (chromadb==0.4.22 / chroma-hnswlib==0.7.3,
/root/reference/requirements.txt:21).  Neither is in this image -- every test here skips, and DESIGN.md says "parity unpinned" --
but an environment that has them turns the restatement (oracle/exact_oracle.py) into a checked one: hnswlib's distance arithmetic
and cosine normalisation through its brute-force index, and Chroma's Collection semantics (query / get / delete with where and
where_document, result shapes, error behaviour) as the reference calls them (app/utils/embedder.py:518, 596, 632, 640)."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "chroma_wal.json")))
    X = np.load(os.path.join(ROOT, "tests", "golden", "chroma_wal.npz"))["vectors"]
    return g, X


@pytest.mark.parametrize("space", ["cosine", "l2", "ip"])
def test_hnswlib_bruteforce_index_agrees_with_the_oracle(space):
    hnswlib = pytest.importorskip("hnswlib")
    from oracle import exact_oracle as eo
    _, X = _golden()
    rng = np.random.default_rng(3)
    X = np.concatenate([X, rng.standard_normal((500, X.shape[1])).astype(np.float32) * 3.0])
    Q = np.concatenate([X[:4] + 0.01, rng.standard_normal((4, X.shape[1])).astype(np.float32)])
    index = hnswlib.BFIndex(space=space, dim=X.shape[1])
    index.init_index(max_elements=X.shape[0])
    index.add_items(X, np.arange(X.shape[0]))
    labels, dists = index.knn_query(Q, k=10)
    Xs, Qs = (eo.normalize_f32(X), eo.normalize_f32(Q)) if space == "cosine" else (X, Q)
    rows, want = eo.topk_exact(Qs, Xs, 10, space)
    for i in range(Q.shape[0]):
        np.testing.assert_allclose(np.sort(dists[i]), want[i], rtol=1e-4, atol=1e-5)     # hnswlib accumulates in fp32 SIMD lanes
        # ids may differ only where hnswlib's fp32 distances tie or cross within that tolerance
        d64 = eo.distances_f64(Qs[i:i + 1], Xs, space)[0]
        assert np.all(d64[np.asarray(labels[i], dtype=np.int64)] <= d64[rows[i][-1]] * (1 + 1e-4) + 1e-5)


def test_chroma_collection_semantics_agree_with_the_oracle():
    chromadb = pytest.importorskip("chromadb")
    from oracle import exact_oracle as eo
    g, X = _golden()
    ids, metas = g["ids"], g["metadatas"]
    docs = [m.get("chroma:document") if m else None for m in metas]
    metas = [{k: v for k, v in (m or {}).items() if not k.startswith("chroma:")} or None for m in metas]
    client = chromadb.Client()
    # exhaustive search parameters: with 70 vectors HNSW then returns the exact neighbours
    col = client.create_collection("pin", metadata={"hnsw:space": g["space"], "hnsw:search_ef": 200, "hnsw:construction_ef": 200, "hnsw:M": 64})
    o = eo.ExactCollection("pin", {"hnsw:space": g["space"]})
    for c in (col, o):
        c.add(ids=ids[1:], embeddings=X[1:].tolist(), metadatas=metas[1:], documents=docs[1:])
    assert col.count() == o.count()
    q = [X[0].tolist()]
    first_type = next(m["type"] for m in metas[1:] if m and "type" in m)
    word = next(d.split()[0] for d in docs[1:] if d)
    for where, wd in ((None, None), ({"type": first_type}, None), (None, {"$contains": word}), ({"type": {"$ne": first_type}}, {"$not_contains": word})):
        kw = {k: v for k, v in (("where", where), ("where_document", wd)) if v}
        got = col.query(query_embeddings=q, n_results=5, include=["metadatas", "documents", "distances"], **kw)
        want = o.query(q, n_results=5, where=where, where_document=wd)
        assert got["ids"] == want["ids"], (where, wd)
        np.testing.assert_allclose(got["distances"][0], want["distances"][0], rtol=1e-4, atol=1e-5)
        assert got["documents"] == want["documents"] and got["metadatas"] == want["metadatas"]
        assert sorted(col.get(**kw)["ids"]) == sorted(o.get(where=where, where_document=wd)["ids"])
    assert got.keys() >= {"ids", "distances", "metadatas", "documents"}
    with pytest.raises(ValueError):
        col.delete()                                   # neither ids nor a clause: an error, not a wipe
    col.add(ids=ids[1:3], embeddings=X[1:3].tolist())  # existing ids: skipped, not duplicated
    assert col.count() == o.count()
    col.delete(ids=ids[1:4]); o.delete(ids=ids[1:4])
    col.upsert(ids=[ids[5], "new"], embeddings=X[[2, 3]].tolist()); o.upsert(ids=[ids[5], "new"], embeddings=X[[2, 3]].tolist())
    assert col.count() == o.count()
    assert col.query(query_embeddings=q, n_results=5)["ids"] == o.query(q, n_results=5)["ids"]
