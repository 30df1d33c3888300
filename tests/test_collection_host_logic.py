"""CPU: the HOST logic of B200Collection -- native id table, tombstones, dictionary-encoded metadata columns and compiled clauses,
where_document, result assembly, Chroma's error behaviour -- against the oracle's ExactCollection, with tests/fake_device.py standing
in for the device half of the library (the GPU suite runs the same calls against the real kernels).  Reference call sites mirrored:
collection.add / upsert / query / get / delete / count (app/utils/embedder.py:518, 596, 632, 640, 700, 888, 901)."""
import numpy as np
import pytest

from conftest import make_unit


@pytest.fixture()
def fake(monkeypatch):
    from multimodal_rag_b200 import _lib, build as b2r_build
    from fake_device import FakeLib
    b2r_build.build()
    lib = FakeLib()
    monkeypatch.setattr(_lib, "load", lambda: lib)
    return lib


def _metas(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        m = {"doc_id": f"doc_{i % 13:04d}", "type": str(rng.choice(["text", "table", "image"]))}
        if i % 3:
            m["page"] = int(i % 9)
        if i % 4 == 0:
            m["score"] = float(i) / 5
        out.append(None if i % 37 == 0 else m)
    return out


def _docs(n, tag):
    return [None if i % 11 == 0 else f"{tag} {('text on page', 'table row', 'image of Figure')[i % 3]} {i % 23}" for i in range(n)]


def _same(got, want, include=("ids", "documents", "metadatas")):
    for key in include:
        assert got[key] == want[key], key
    if want.get("distances") is not None:
        for a, b in zip(got["distances"], want["distances"]):
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("space", ["cosine", "l2", "ip"])
def test_collection_follows_the_oracle_through_a_mixed_workload(fake, space):
    from multimodal_rag_b200 import B200Collection
    from oracle import exact_oracle as eo
    d = 24
    c = B200Collection("h", {"hnsw:space": space})
    o = eo.ExactCollection("h", {"hnsw:space": space})
    rng = np.random.default_rng(11)
    Q = make_unit(5, d, 77)
    wheres = [None, {"type": "image"}, {"type": {"$in": ["text", "table"]}}, {"$and": [{"page": {"$gte": 3}}, {"type": {"$ne": "text"}}]},
              {"$or": [{"doc_id": "doc_0003"}, {"score": {"$gt": 30.0}}]}, {"missing": 1}]
    wdocs = [None, {"$contains": "table row"}, {"$not_contains": "page"}, {"$and": [{"$contains": "b1"}, {"$not_contains": "Figure 2"}]}]
    next_id = 0
    for step in range(8):
        n = int(rng.integers(20, 120))
        ids = [f"id{next_id + i}" for i in range(n)]
        if step >= 2:                                    # some ids of the batch exist already
            for j in rng.choice(n, size=n // 4, replace=False):
                ids[j] = f"id{int(rng.integers(0, next_id))}"
            ids = list(dict.fromkeys(ids))
        next_id += n
        X = make_unit(len(ids), d, 100 + step) * (1.0 + step)
        metas, docs = _metas(len(ids), step), _docs(len(ids), f"b{step}")
        for col in (c, o):
            (col.upsert if step % 2 else col.add)(ids=ids, embeddings=X, metadatas=metas, documents=docs)
        assert c.count() == o.count()
        if step % 3 == 2:
            kill = [f"id{int(i)}" for i in rng.integers(0, next_id, size=15)]
            c.delete(ids=kill); o.delete(ids=kill)
            c.delete(where={"page": int(step)}); o.delete(where={"page": int(step)})
        for where in wheres:
            for wd in wdocs[: 2 if where else 4]:
                want = o.query(Q, n_results=6, where=where, where_document=wd)
                got = c.query(query_embeddings=Q, n_results=6, where=where, where_document=wd)
                _same(got, {**want, "distances": want["distances"]})
                assert c.get(where=where, where_document=wd)["ids"] == o.get(where=where, where_document=wd)["ids"]
    some = [f"id{i}" for i in range(0, next_id, 7)] + ["never added"]
    _same(c.get(ids=some), o.get(ids=some))
    got = c.get(ids=some, include=["embeddings"])
    want = o.get(ids=some, include=["embeddings"])
    np.testing.assert_array_equal(np.asarray(got["embeddings"], dtype=np.float32), np.asarray(want["embeddings"], dtype=np.float32))
    assert c.get(limit=5, offset=3)["ids"] == o.get()["ids"][3:8]
    big = c.query(query_embeddings=Q[:1], n_results=10_000)          # more than there are rows: every live row, once
    assert len(big["ids"][0]) == c.count() and len(set(big["ids"][0])) == c.count()
    gone = c.delete(where_document={"$contains": "image of"})
    before = set(o._row_of)
    o.delete(where_document={"$contains": "image of"})
    assert sorted(gone) == sorted(before - set(o._row_of)) and c.count() == o.count()
    c.close()


def test_bad_batches_change_nothing(fake):
    from multimodal_rag_b200 import B200Collection
    c = B200Collection("e", {"hnsw:space": "cosine"})
    X = make_unit(4, 8, 1)
    c.add(ids=["a", "b", "c", "d"], embeddings=X)
    for kw in (dict(ids=["x", "x"], embeddings=X[:2]), dict(ids=["x", ""], embeddings=X[:2]), dict(ids=["x", 3], embeddings=X[:2]),
               dict(ids=["x"], embeddings=X[:2]), dict(ids=["x", "y"], embeddings=make_unit(2, 9, 2)),
               dict(ids=["x", "y"], embeddings=X[:2], metadatas=[{"k": [1]}, None]), dict(ids="x", embeddings=None)):
        for call in (c.add, c.upsert):
            with pytest.raises(ValueError):
                call(**kw)
    assert c.count() == 4 and c.get()["ids"] == ["a", "b", "c", "d"]
    c.add(ids=["b", "e"], embeddings=X[:2])                          # an existing id is skipped, the new one is stored
    assert c.count() == 5 and c.get(ids=["e"])["ids"] == ["e"]
    Y = make_unit(3, 8, 9)
    c.upsert(ids=["a\0b", "naïve", "a"], embeddings=Y)               # ids with a NUL byte / non-ASCII round-trip; "a" is overwritten
    assert c.count() == 7 and sorted(c.get()["ids"]) == sorted(["b", "c", "d", "e", "a\0b", "naïve", "a"])
    r = c.query(query_embeddings=Y, n_results=1)
    assert [x[0] for x in r["ids"]] == ["a\0b", "naïve", "a"]
    for bad in (0, -1, 2.5, True):
        with pytest.raises(ValueError):
            c.query(query_embeddings=X[:1], n_results=bad)
    with pytest.raises(ValueError):
        c.query(query_embeddings=make_unit(1, 9, 3), n_results=1)
    with pytest.raises(ValueError):
        c.get(include=["distances"])
    with pytest.raises(ValueError):
        c.delete()
    c.close()
