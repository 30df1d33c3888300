"""The collection's host tables at sizes where per-row Python work shows: bulk adds of bare vectors, upserts that overwrite,
get / delete by a where clause evaluated on the device, and an upsert that fails half way."""
import numpy as np
import pytest

from conftest import make_unit

pytestmark = pytest.mark.gpu


def test_bulk_add_upsert_overwrite_and_device_where():
    import torch
    from multimodal_rag_b200 import B200Collection
    from oracle import exact_oracle as eo
    n, d = 300_000, 128
    c = B200Collection("big", {"hnsw:space": "cosine"}, capacity=n + 20_000)
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.nn.functional.normalize(torch.randn(n, d, generator=g, device="cuda"), dim=1)
    c.add(ids=[f"r{i}" for i in range(n)], embeddings=X)                     # bare vectors: no per-row metadata work
    assert c.count() == n
    # upsert: 10 % of the batch overwrites existing ids, the rest is new
    m = 8192
    U = torch.nn.functional.normalize(torch.randn(m, d, generator=g, device="cuda"), dim=1)
    ids = [f"n{i}" for i in range(m)]
    over = np.random.default_rng(2).choice(n, size=m // 10, replace=False)
    for j, o in enumerate(over.tolist()):
        ids[j] = f"r{o}"
    c.upsert(ids=ids, embeddings=U, metadatas=[{"type": "image" if i % 2 else "text", "page": i % 7} for i in range(m)])
    assert c.count() == n + m - len(over)
    rows, dist, cnt = c.query_rows(U[:16], 3)
    assert [c._ids[r] for r in rows[:, 0]] == ids[:16] and (np.abs(dist[:, 0]) < 1e-5).all()
    assert not np.isin(rows, over).any()                                     # the overwritten rows never come back
    # get / delete through a where clause: evaluated by the device's clause kernel over the columns
    got = c.get(where={"$and": [{"type": "image"}, {"page": {"$gte": 5}}]}, include=[])
    want = [ids[i] for i in range(m) if i % 2 and i % 7 >= 5]
    assert sorted(got["ids"]) == sorted(want)
    gone = c.delete(where={"page": 6})
    assert sorted(gone) == sorted(ids[i] for i in range(m) if i % 7 == 6)
    assert c.count() == n + m - len(over) - len(gone)
    # final state against the oracle on a sample of queries
    Xall = np.concatenate([X.cpu().numpy(), U.cpu().numpy()])
    alive = c._alive.copy()
    Q = make_unit(8, d, 9)
    rows, dist, cnt = c.query_rows(Q, 10)
    er, ed = eo.topk_exact(eo.normalize_f32(Q), eo.normalize_f32(Xall), 10, "cosine", allowed=alive)
    for i in range(8):
        np.testing.assert_array_equal(rows[i], er[i])
        np.testing.assert_allclose(dist[i], ed[i], rtol=1e-5, atol=1e-7)
    c.close()


def test_failed_upsert_leaves_the_old_rows():
    from multimodal_rag_b200 import B200Collection
    X = make_unit(100, 64, 3)
    c = B200Collection("u", {"hnsw:space": "cosine"})
    c.add(ids=[f"a{i}" for i in range(100)], embeddings=X)
    with pytest.raises(ValueError):
        c.upsert(ids=["a1", "a2"], embeddings=make_unit(2, 32, 4))           # wrong dimension: rejected before anything changes
    with pytest.raises(ValueError):
        c.upsert(ids=["a1", "a1"], embeddings=make_unit(2, 64, 4))
    assert c.count() == 100
    rows, dist, cnt = c.query_rows(X[1:3], 1)
    assert [c._ids[r] for r in rows[:, 0]] == ["a1", "a2"] and (np.abs(dist[:, 0]) < 1e-6).all()
    c.close()
