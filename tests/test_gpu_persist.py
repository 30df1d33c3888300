"""Persistence: a saved collection reloads to the same device state and answers bit-identically
(b2r_save / b2r_load behind B200Collection.save / load and the persistent B200Client -- the reference keeps its
collection under ChromaSettings(persist_directory=...), app/utils/embedder.py:164-170)."""
import os

import numpy as np
import pytest

from conftest import make_unit

pytestmark = pytest.mark.gpu


def _build(space, n=6000, d=384, **kw):
    from multimodal_rag_b200 import B200Collection
    X = make_unit(n, d, 21) * (1.0 if space == "cosine" else 1.3)
    types = np.random.default_rng(5).choice(["text", "image", "table"], size=n, p=[0.6, 0.3, 0.1])
    c = B200Collection("persisted", {"hnsw:space": space}, **kw)
    c.add(ids=[f"doc_{i:06d}" for i in range(n)], embeddings=X,
          metadatas=[{"type": str(t), "doc_id": f"d{i % 17}", "page": int(i % 5)} for i, t in enumerate(types)],
          documents=[f"summary {i}" for i in range(n)])
    c.delete(ids=[f"doc_{i:06d}" for i in range(0, n, 13)])                    # tombstones
    c.upsert(ids=["doc_000001", "brand_new"], embeddings=make_unit(2, d, 77),   # overwrite + append
             metadatas=[{"type": "image"}, {"type": "text"}], documents=["u1", "u2"])
    return c


@pytest.mark.parametrize("space,kw", [("cosine", {}), ("l2", {}), ("ip", {"keep_f32_master": False})])
def test_save_load_round_trip_is_bit_identical(tmp_path, space, kw):
    from multimodal_rag_b200 import B200Collection
    c = _build(space, **kw)
    Q = make_unit(9, 384, 3)
    before = [c.query_rows(Q, 10, w, want_dist64=True) for w in (None, {"type": "image"}, {"doc_id": "d3"})]
    res_before = c.query(query_embeddings=Q[:2].tolist(), n_results=4, where={"type": "text"})
    emb_before = c.get(ids=["brand_new", "doc_000002"], include=["embeddings", "metadatas", "documents"])
    c.save(str(tmp_path))
    assert os.path.exists(tmp_path / "persisted.b2r") and os.path.exists(tmp_path / "persisted.tables.json")
    n_live = c.count()
    c.close()

    r = B200Collection.load(str(tmp_path), "persisted")
    assert r.count() == n_live and r.space == space and r.dimension == 384
    after = [r.query_rows(Q, 10, w, want_dist64=True) for w in (None, {"type": "image"}, {"doc_id": "d3"})]
    for (r0, d0, c0, e0), (r1, d1, c1, e1) in zip(before, after):
        np.testing.assert_array_equal(r0, r1)
        np.testing.assert_array_equal(c0, c1)
        np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))      # bit-identical distances
        np.testing.assert_array_equal(e0.view(np.uint64), e1.view(np.uint64))
    assert r.query(query_embeddings=Q[:2].tolist(), n_results=4, where={"type": "text"}) == res_before
    assert r.get(ids=["brand_new", "doc_000002"], include=["embeddings", "metadatas", "documents"]) == emb_before
    # the reloaded collection keeps working: deleted ids stay deleted, new rows land after the old ones
    assert r.get(ids=["doc_000000"])["ids"] == []
    x = make_unit(1, 384, 99)
    r.add(ids=["after_reload"], embeddings=x, metadatas=[{"type": "table"}])
    got = r.query(query_embeddings=x.tolist(), n_results=1)
    assert got["ids"][0] == ["after_reload"]
    r.close()


def test_corrupt_and_truncated_files_are_rejected(tmp_path):
    from multimodal_rag_b200 import B200Collection
    c = _build("cosine", n=2000)
    c.save(str(tmp_path))
    c.close()
    path = tmp_path / "persisted.b2r"
    blob = path.read_bytes()
    path.write_bytes(blob[:200] + bytes([blob[200] ^ 0x40]) + blob[201:])          # one flipped payload bit
    with pytest.raises(ValueError, match="checksum"):
        B200Collection.load(str(tmp_path), "persisted")
    path.write_bytes(blob[: len(blob) // 2])
    with pytest.raises(ValueError, match="truncated"):
        B200Collection.load(str(tmp_path), "persisted")
    path.write_bytes(blob)
    B200Collection.load(str(tmp_path), "persisted").close()


def test_persistent_client(tmp_path):
    from multimodal_rag_b200 import B200Client
    cl = B200Client(path=str(tmp_path))
    col = cl.create_collection("rag", {"hnsw:space": "cosine"})
    X = make_unit(300, 384, 1)
    col.add(ids=[f"a{i}" for i in range(300)], embeddings=X, metadatas=[{"type": "text"}] * 300)
    empty = cl.create_collection("empty_one")
    assert empty.count() == 0
    cl.persist()
    want = col.query(query_embeddings=X[:3].tolist(), n_results=3)

    cl2 = B200Client(path=str(tmp_path))
    assert sorted(c.name for c in cl2.list_collections()) == ["empty_one", "rag"]
    assert cl2.get_collection("rag").query(query_embeddings=X[:3].tolist(), n_results=3) == want
    assert cl2.get_collection("empty_one").count() == 0
    cl2.delete_collection("rag")
    assert not os.path.exists(tmp_path / "rag.b2r") and not os.path.exists(tmp_path / "rag.tables.json")
    assert [c.name for c in B200Client(path=str(tmp_path)).list_collections()] == ["empty_one"]
