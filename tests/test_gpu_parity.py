"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from conftest import make_unit

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import exact_oracle as eo
    return eo


def _mk(space, d, n, seed=0, unit=True, **kw):
    from multimodal_rag_b200 import B200Collection
    X = make_unit(n, d, seed)
    if not unit:
        X = X * np.random.default_rng(seed + 1).uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)
    c = B200Collection("t", {"hnsw:space": space}, **kw)
    ids = [f"doc_{i // 7:012x}_text_{i}" for i in range(n)]
    c.add(ids=ids, embeddings=X)
    return c, X, ids


def _check(c, X, Q, k, space, where_mask=None, where=None, path=None):
    eo = _oracle()
    if path is not None:
        c.set_path(path)
    rows, dist, cnt = c.query_rows(Q, k, where)
    Xs = eo.normalize_f32(X) if space == "cosine" else X
    Qs = eo.normalize_f32(Q) if space == "cosine" else Q
    er, ed = eo.topk_exact(Qs, Xs, k, space, allowed=where_mask)
    for i in range(Q.shape[0]):
        assert cnt[i] == len(er[i]), (i, cnt[i], len(er[i]))
        np.testing.assert_array_equal(rows[i, : cnt[i]], er[i])          # bit-exact ids
        np.testing.assert_allclose(dist[i, : cnt[i]], ed[i], rtol=1e-5, atol=1e-7)   # north_star: 1e-5 relative
        assert (rows[i, cnt[i]:] == -1).all() and np.isinf(dist[i, cnt[i]:]).all()


@pytest.mark.parametrize("space", ["cosine", "l2", "ip"])
@pytest.mark.parametrize("d", [384, 512, 768])
def test_scan_matches_oracle(space, d):
    c, X, _ = _mk(space, d, 20000, seed=d, unit=(space == "cosine"))
    Q = make_unit(7, d, 99)
    _check(c, X, Q, 5, space, path=1)
    _check(c, X, Q[:1], 10, space, path=1)
    _check(c, X, Q[:3], 16, space, path=1)
    if space == "cosine":       # unit-norm random data: the bf16 certificate must hold without help
        assert c.stats()["n_exact_fallbacks"] == 0


@pytest.mark.parametrize("space", ["cosine", "l2", "ip"])
def test_exact_path_matches_oracle(space):
    c, X, _ = _mk(space, 384, 5000, seed=3, unit=False)
    Q = make_unit(5, 384, 7) * 1.7
    _check(c, X, Q, 20, space, path=3)
    _check(c, X, Q, 100, space, path=3)


@pytest.mark.parametrize("k", [20, 32, 50, 100])
def test_scan_large_k(k):
    c, X, _ = _mk("cosine", 384, 30000, seed=11)
    Q = make_unit(3, 384, 5)
    _check(c, X, Q, k, "cosine", path=1)


def test_odd_dimension_goes_through_exact_path():
    c, X, _ = _mk("l2", 100, 3000, seed=5, unit=False)
    Q = make_unit(4, 100, 8)
    _check(c, X, Q, 5, "l2")


def test_k_larger_than_collection_and_empty():
    from multimodal_rag_b200 import B200Collection
    c, X, ids = _mk("cosine", 384, 9, seed=1)
    Q = make_unit(2, 384, 2)
    _check(c, X, Q, 16, "cosine")
    r = c.query(query_embeddings=Q.tolist(), n_results=16)
    assert [len(x) for x in r["ids"]] == [9, 9]
    e = B200Collection("e", {"hnsw:space": "cosine"})
    r = e.query(query_embeddings=Q.tolist(), n_results=5)
    assert r["ids"] == [[], []] and r["distances"] == [[], []]


def test_duplicates_and_ties_resolve_to_lowest_row():
    from multimodal_rag_b200 import B200Collection
    d = 384
    base = make_unit(50, d, 4)
    X = np.concatenate([base, base, base[:10]])           # exact duplicates -> exact distance ties
    c = B200Collection("dup", {"hnsw:space": "cosine"})
    c.add(ids=[f"id{i}" for i in range(X.shape[0])], embeddings=X)
    Q = base[:6] + 0.01 * make_unit(6, d, 6)
    _check(c, X, Q, 8, "cosine")
    # one-hot rows: every distance is exactly representable, massive ties
    E = np.zeros((300, d), dtype=np.float32)
    E[np.arange(300), np.arange(300) % 3] = 1.0
    c2 = B200Collection("onehot", {"hnsw:space": "l2"})
    c2.add(ids=[f"e{i}" for i in range(300)], embeddings=E)
    q = np.zeros((1, d), dtype=np.float32); q[0, 1] = 1.0
    _check(c2, E, q, 16, "l2")


def test_type_filter_and_bitmap_filter():
    from multimodal_rag_b200 import B200Collection
    d, n = 512, 20000
    X = make_unit(n, d, 21)
    rng = np.random.default_rng(0x7E57)
    types = rng.choice(["text", "table", "image"], size=n, p=[0.6, 0.1, 0.3])
    metas = [{"type": str(t), "doc_id": f"doc_{i % 50}", "page": int(i % 13)} for i, t in enumerate(types)]
    c = B200Collection("mm", {"hnsw:space": "cosine"})
    c.add(ids=[f"r{i}" for i in range(n)], embeddings=X, metadatas=metas)
    Q = make_unit(4, d, 22)
    _check(c, X, Q, 10, "cosine", where_mask=(types == "image"), where={"type": "image"})
    _check(c, X, Q, 10, "cosine", where_mask=np.isin(types, ["image", "table"]),
           where={"type": {"$in": ["image", "table"]}})
    page = np.arange(n) % 13
    _check(c, X, Q, 10, "cosine", where_mask=(types == "text") & (page >= 11),
           where={"$and": [{"type": "text"}, {"page": {"$gte": 11}}]})
    _check(c, X, Q[:1], 10, "cosine", where_mask=(np.arange(n) % 50 == 7), where={"doc_id": "doc_7"})
    r = c.query(query_embeddings=Q[:1], n_results=10, where={"type": "video"})
    assert r["ids"] == [[]]


def test_delete_upsert_visibility():
    c, X, ids = _mk("cosine", 384, 4000, seed=31)
    eo = _oracle()
    Q = X[:3] + 0.05 * make_unit(3, 384, 32)
    c.delete(ids=ids[:3])
    alive = np.ones(4000, dtype=bool); alive[:3] = False
    _check(c, X, Q, 5, "cosine", where_mask=alive)
    assert c.count() == 3997
    # upsert: overwrite 10 existing ids with new vectors, add 5 new ones
    newv = make_unit(15, 384, 33)
    up_ids = ids[100:110] + [f"new_{i}" for i in range(5)]
    c.upsert(ids=up_ids, embeddings=newv)
    assert c.count() == 3997 + 5
    r = c.query(query_embeddings=newv[:15], n_results=1)
    assert [x[0] for x in r["ids"]] == up_ids          # each upserted row is its own top-1
    alive[100:110] = False
    Xall = np.concatenate([X, newv]); alive_all = np.concatenate([alive, np.ones(15, dtype=bool)])
    _check(c, Xall, Q, 5, "cosine", where_mask=alive_all)


def test_golden_fixture_known_answers(golden):
    from multimodal_rag_b200 import B200Collection
    X, ids, metas = golden["vectors"], golden["ids"], golden["metadatas"]
    for space, key in (("cosine", "cosine"), ("l2", "l2")):
        c = B200Collection(golden["collection"], {"hnsw:space": space})
        c.add(ids=ids[1:], embeddings=X[1:].tolist(), metadatas=metas[1:],
              documents=[m["chroma:document"] for m in metas[1:]])
        r = c.query(query_embeddings=[X[0].tolist()], n_results=5,
                    include=["metadatas", "documents", "distances"])
        assert r["ids"][0] == [a["id"] for a in golden["known"]["top5"]]
        np.testing.assert_allclose(r["distances"][0], [a[key] for a in golden["known"]["top5"]], rtol=1e-5)
        assert all(m["type"] == "text" for m in r["metadatas"][0])
        r = c.query(query_embeddings=[X[0].tolist()], n_results=5, where={"type": "image"})
        assert r["ids"][0] == [a["id"] for a in golden["known"]["top5_image"]]
        np.testing.assert_allclose(r["distances"][0], [a[key] for a in golden["known"]["top5_image"]], rtol=1e-5)


def test_get_embeddings_roundtrip_and_bf16_only_mode():
    eo = _oracle()
    c, X, ids = _mk("cosine", 384, 1000, seed=41)
    g = c.get(ids=[ids[5], ids[900]], include=["embeddings"])
    np.testing.assert_array_equal(np.asarray(g["embeddings"], dtype=np.float32), eo.normalize_f32(X[[5, 900]]))
    # bf16-only corpus: the stored rows are the bf16 roundings, and the answer is exact w.r.t. them
    import torch
    c2, X2, _ = _mk("ip", 384, 8000, seed=42, keep_f32_master=False)
    Xb = torch.from_numpy(X2).to(torch.bfloat16).to(torch.float32).numpy()
    g = c2.get(ids=["doc_000000000000_text_3"], include=["embeddings"])
    np.testing.assert_array_equal(np.asarray(g["embeddings"], dtype=np.float32)[0], Xb[3])
    Q = make_unit(3, 384, 43)
    _check(c2, Xb, Q, 5, "ip")


def test_large_corpus_properties():
    """BASELINE config 2 size (1M x 384): checked through size-independent properties --
    planted neighbours are found, results are sorted, and the top-k of the union equals the
    merge of the shard top-ks (the identity the multi-GPU path relies on)."""
    import torch
    from multimodal_rag_b200 import B200Collection
    n, d, k = 1_000_000, 384, 5
    g = torch.Generator(device="cuda").manual_seed(0xC0FFEE)
    X = torch.randn(n, d, generator=g, device="cuda", dtype=torch.float32)
    X = torch.nn.functional.normalize(X, dim=1)
    c = B200Collection("big", {"hnsw:space": "cosine"}, capacity=n, dimension=d)
    ids = [str(i) for i in range(n)]
    c.add(ids=ids, embeddings=X)
    planted = torch.tensor([3, 499_999, 999_999], device="cuda")
    Q = torch.nn.functional.normalize(X[planted] + 0.02 * torch.randn(3, d, generator=g, device="cuda"), dim=1)
    rows, dist, cnt = c.query_rows(Q, k)
    assert rows[:, 0].tolist() == planted.tolist()
    assert (np.diff(dist, axis=1) >= 0).all() and (cnt == k).all()
    # the fp64 C oracle over the same stored rows: ids bit-exact, distances to 1e-5 relative
    from oracle import c_oracle
    c_oracle.set_threads(0)
    er, ed, _ = c_oracle.topk(c_oracle.normalize_f32(X.cpu().numpy()), c_oracle.normalize_f32(Q.cpu().numpy()), k, "cosine", acc64=True)
    np.testing.assert_array_equal(rows, er)
    np.testing.assert_allclose(dist, ed, rtol=1e-5, atol=1e-7)
    assert c.stats()["n_exact_fallbacks"] == 0
