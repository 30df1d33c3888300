"""CPU: the oracle (numpy + C) against the reference's committed Chroma WAL fixture and its
known answers, and the two restatements against each other."""
import numpy as np
import pytest

from oracle import c_oracle, exact_oracle as eo
from conftest import make_unit


def test_fixture_shape(golden):
    X = golden["vectors"]
    assert X.shape == (70, 384) and X.dtype == np.float32
    assert golden["space"] == "cosine" and golden["dimension"] == 384
    assert sum(1 for op in golden["wal_ops"] if op[1] == 0) == 70 and sum(1 for op in golden["wal_ops"] if op[1] == 3) == 70
    n = np.linalg.norm(X.astype(np.float64), axis=1)
    assert n.min() > 0.999999 and n.max() < 1.000001
    types = [m["type"] for m in golden["metadatas"]]
    assert types.count("text") == 48 and types.count("image") == 22


# SURVEY.md App. B, computed independently there; hard-coded here so the fixture json cannot drift
APP_B = [("doc_d8164983ea8e_text_35", 0.8956640, 1.7913280), ("doc_d8164983ea8e_text_9", 0.8957669, 1.7915338),
         ("doc_d8164983ea8e_text_0", 0.9043092, 1.8086183), ("doc_d8164983ea8e_text_12", 0.9060158, 1.8120315),
         ("doc_d8164983ea8e_text_23", 0.9253784, 1.8507566)]
APP_B_IMAGE = [("page_19_19130f87", 0.938270), ("page_14_dca538e9", 0.950331), ("page_7_45d34ba7", 0.953245),
               ("page_1_c3ebbd03", 0.957183), ("page_11_cc89b5d3", 0.961343)]


@pytest.mark.parametrize("space", ["cosine", "l2"])
def test_collection_known_answers(golden, space):
    X, ids, metas = golden["vectors"], golden["ids"], golden["metadatas"]
    c = eo.ExactCollection(golden["collection"], {"hnsw:space": space})
    c.add(ids=ids[1:], embeddings=X[1:].tolist(), metadatas=metas[1:],
          documents=[m["chroma:document"] for m in metas[1:]])
    assert c.count() == 69
    r = c.query(query_embeddings=[X[0].tolist()], n_results=5)
    col = 1 if space == "cosine" else 2
    assert r["ids"][0] == [a[0] for a in APP_B]
    np.testing.assert_allclose(r["distances"][0], [a[col] for a in APP_B], rtol=2e-7 * 5, atol=6e-8)
    assert r["ids"][0] == [a["id"] for a in golden["known"]["top5"]]
    np.testing.assert_allclose(r["distances"][0], [a[space] for a in golden["known"]["top5"]], rtol=1e-6)
    assert [m["type"] for m in r["metadatas"][0]] == ["text"] * 5
    assert r["documents"][0][0] == metas[ids.index(APP_B[0][0])]["chroma:document"]
    r = c.query(query_embeddings=[X[0].tolist()], n_results=5, where={"type": "image"})
    assert [i.split("_C_")[-1] for i in r["ids"][0]] == [a[0] for a in APP_B_IMAGE]
    if space == "cosine":
        np.testing.assert_allclose(r["distances"][0], [a[1] for a in APP_B_IMAGE], atol=1e-6)


def test_similar_documents_flow(golden):
    """get(ids, include=embeddings) -> query(n+1) -> drop self (app/utils/embedder.py:861-930)."""
    X, ids = golden["vectors"], golden["ids"]
    c = eo.ExactCollection("c", {"hnsw:space": "cosine"})
    c.add(ids=ids, embeddings=X.tolist())
    src = c.get(ids=[ids[7]], include=["embeddings"])
    r = c.query(query_embeddings=[src["embeddings"][0]], n_results=11)
    assert r["ids"][0][0] == ids[7] and abs(r["distances"][0][0]) < 1e-6
    rest = [i for i in r["ids"][0] if i != ids[7]][:10]
    assert rest == [a["id"] for a in golden["known"]["top10_self_excluded_row7"]]


def test_wal_replay_leaves_empty_collection(golden):
    """The fixture's WAL is 70 ADDs then 70 DELETEs: replaying it yields count()==0 and [[]]."""
    X, ids = golden["vectors"], golden["ids"]
    row = {i: n for n, i in enumerate(ids)}
    c = eo.ExactCollection("c", {"hnsw:space": "cosine"})
    for _, op, id_ in golden["wal_ops"]:
        if op == 0:
            c.add(ids=[id_], embeddings=[X[row[id_]].tolist()])
        elif op == 3:
            c.delete(ids=[id_])
    assert c.count() == 0
    assert c.query(query_embeddings=[X[0].tolist()], n_results=5)["ids"] == [[]]


@pytest.mark.parametrize("space", ["l2", "cosine", "ip"])
def test_c_port_matches_numpy_oracle(space):
    X = make_unit(5000, 96, 1) * np.random.default_rng(2).uniform(0.5, 2, (5000, 1)).astype(np.float32)
    Q = make_unit(6, 96, 3)
    Xs, Qs = (eo.normalize_f32(X), eo.normalize_f32(Q)) if space == "cosine" else (X, Q)
    np.testing.assert_array_equal(eo.normalize_f32(X), c_oracle.normalize_f32(X))
    allowed = (np.arange(5000) % 3 != 0)
    for mask in (None, allowed):
        rn, dn = eo.topk_exact(Qs, Xs, 17, space, allowed=mask)
        rc, dc, cc = c_oracle.topk(Xs, Qs, 17, space, allowed=mask, acc64=True)
        assert (cc == 17).all()
        np.testing.assert_array_equal(rc, np.stack(rn))
        np.testing.assert_allclose(dc, np.stack(dn), rtol=1e-6, atol=1e-7)
    # fp32 fast mode (the timed baseline) agrees on distances to fp32 accuracy
    rf, df, _ = c_oracle.topk(Xs, Qs, 17, space, acc64=False)
    np.testing.assert_allclose(df, np.stack(eo.topk_exact(Qs, Xs, 17, space)[1]), rtol=1e-4, atol=1e-5)
    bi, bd = eo.topk_bruteforce_f32(Qs, Xs, 17, space)
    np.testing.assert_allclose(bd, np.stack(eo.topk_exact(Qs, Xs, 17, space)[1]), rtol=1e-3, atol=1e-4)


def test_ties_resolve_to_lowest_insertion_index_and_k_clamps():
    E = np.zeros((30, 8), dtype=np.float32)
    E[np.arange(30), np.arange(30) % 3] = 1.0
    q = np.zeros((1, 8), dtype=np.float32); q[0, 1] = 1.0
    r, d = eo.topk_exact(q, E, 12, "l2")
    assert r[0].tolist() == [1, 4, 7, 10, 13, 16, 19, 22, 25, 28, 0, 2]
    rc, dc, cc = c_oracle.topk(E, q, 12, "l2")
    assert rc[0].tolist() == r[0].tolist()
    rc, dc, cc = c_oracle.topk(E, q, 40, "l2")
    assert cc[0] == 30 and (rc[0, 30:] == -1).all() and np.isinf(dc[0, 30:]).all()
    c = eo.ExactCollection("c")                     # default space is l2 (embedder.py:179-182)
    assert c.space == "l2"
    c.add(ids=[f"i{j}" for j in range(30)], embeddings=E)
    assert len(c.query(query_embeddings=q, n_results=40)["ids"][0]) == 30


def test_add_semantics():
    c = eo.ExactCollection("c", {"hnsw:space": "cosine"})
    c.add(ids=["a", "b"], embeddings=[[1, 0], [0, 1]], metadatas=[{"type": "text"}, {"type": "image"}])
    c.add(ids=["a", "c"], embeddings=[[0, 1], [1, 1]])          # existing id skipped
    assert c.count() == 3 and c.get(ids=["a"], include=["embeddings"])["embeddings"][0] == [1.0, 0.0]
    with pytest.raises(ValueError):
        c.add(ids=["d", "d"], embeddings=[[1, 0], [1, 0]])
    with pytest.raises(ValueError):
        c.add(ids=["e"], embeddings=[[1, 0, 0]])
    c.upsert(ids=["a"], embeddings=[[0, 1]])
    assert c.get(ids=["a"], include=["embeddings"])["embeddings"][0] == [0.0, 1.0]
    assert c.get(where={"type": "image"}, include=[])["ids"] == ["b"]
    c.delete(ids=["b"])
    assert c.count() == 2 and c.get(where={"type": "image"}, include=[])["ids"] == []
