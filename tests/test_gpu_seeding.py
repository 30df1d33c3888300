"""K3's in-kernel threshold seeding (sampling tiles -> posts -> cross-CTA fold -> bounded wait) only switches on when
every CTA owns >= 8 tiles, i.e. from ~303k rows: these cases run it at 330k rows x 128 dims against the fp64 C oracle
(oracle/exact_topk.c) -- ids bit-exact, distances to 1e-5 relative -- over batch sizes that cover one lane per warp,
ragged query blocks and three query blocks, every list length, filters that leave almost nothing to seed from, planted
neighbours inside and outside the sampled tiles, and exact ties at the bound."""
import numpy as np
import pytest

from conftest import make_unit

pytestmark = pytest.mark.gpu

N, D = 330_000, 128
SPECIAL = (999, 70_999, 200_999, N - 1)


def _check(c, Xs, Q, k, space, allowed=None, where=None):
    from oracle import c_oracle
    rows, dist, cnt = c.query_rows(Q, k, where)
    Qs = c_oracle.normalize_f32(Q) if space == "cosine" else Q
    er, ed, ec = c_oracle.topk(Xs, Qs, k, space, allowed=allowed, acc64=True)
    np.testing.assert_array_equal(cnt, ec)
    for i in range(Q.shape[0]):
        np.testing.assert_array_equal(rows[i, : cnt[i]], er[i, : ec[i]])
        np.testing.assert_allclose(dist[i, : cnt[i]], ed[i, : ec[i]], rtol=1e-5, atol=1e-7)


@pytest.fixture(scope="module")
def cosine():
    from multimodal_rag_b200 import B200Collection
    from oracle import c_oracle
    X = make_unit(N, D, 31)
    types = np.where(np.arange(N) % 97 == 0, "rare", np.where(np.arange(N) % 3 == 0, "image", "text"))
    c = B200Collection("seed", {"hnsw:space": "cosine"}, capacity=N)
    c.add(ids=[f"r{i}" for i in range(N)], embeddings=X,
          metadatas=[dict({"type": str(t), "bucket": int(i % 1000)}, **({"special": 1} if i in SPECIAL else {}))
                     for i, t in enumerate(types)])
    yield c, X, c_oracle.normalize_f32(X), types
    c.close()


@pytest.mark.parametrize("nq,k", [(1, 5), (3, 8), (33, 5), (130, 16), (300, 5), (64, 32), (257, 20)])
def test_seeded_scan_matches_oracle(cosine, nq, k):
    c, X, Xs, _ = cosine
    Q = make_unit(nq, D, 100 + nq)
    # planted neighbours: in the first tile of a slice (sampled), deep inside a slice, in the last ragged tile
    plant = [0, 5, 2229 * 3 + 1000, N // 2 + 17, N - 1][: min(5, nq)]
    Q[: len(plant)] = X[plant] + 0.02 * make_unit(len(plant), D, 7)
    c.set_path(2)
    before = c.stats()["n_exact_fallbacks"]
    _check(c, Xs, Q, k, "cosine")
    assert c.stats()["n_exact_fallbacks"] == before          # the certificate holds without the exact fix-up


def test_filters_that_starve_the_seed(cosine):
    c, X, Xs, types = cosine
    Q = make_unit(40, D, 55)
    c.set_path(2)
    _check(c, Xs, Q, 5, "cosine", allowed=(types == "rare"), where={"type": "rare"})          # ~1 % of the rows
    # 330 rows pass (one per thousand): far fewer than the posts a seed needs from most slices
    _check(c, Xs, Q[:9], 5, "cosine", allowed=(np.arange(N) % 1000 == 999), where={"bucket": 999})
    # 4 rows pass: no slice can post L values, the seed word says "no bound"
    four = np.zeros(N, dtype=bool)
    four[list(SPECIAL)] = True
    _check(c, Xs, Q[:9], 5, "cosine", allowed=four, where={"special": 1})
    rows, dist, cnt = c.query_rows(Q[:3], 5, {"bucket": {"$in": [-1]}})                      # nothing passes
    assert (cnt == 0).all() and (rows == -1).all()


def test_ties_at_the_bound_fall_back_exactly():
    """20k exact copies of one row: the seed equals the best score, nothing beats it strictly, the certificate must
    refuse and the exact scan must return the lowest rows."""
    from multimodal_rag_b200 import B200Collection
    from oracle import c_oracle
    X = make_unit(N, D, 77)
    dup = np.arange(1000, N, 16)[:20_000]
    X[dup] = X[123]
    c = B200Collection("ties", {"hnsw:space": "cosine"}, capacity=N)
    c.add(ids=[f"t{i}" for i in range(N)], embeddings=X)
    Q = np.concatenate([X[123:124], make_unit(30, D, 78)])
    c.set_path(2)
    _check(c, c_oracle.normalize_f32(X), Q, 8, "cosine")
    c.close()


@pytest.mark.parametrize("space", ["l2", "ip"])
def test_seeded_scan_other_spaces(space):
    from multimodal_rag_b200 import B200Collection
    X = make_unit(N, D, 91) * np.random.default_rng(5).uniform(0.5, 2.0, size=(N, 1)).astype(np.float32)
    c = B200Collection("sp", {"hnsw:space": space}, capacity=N)
    c.add(ids=[f"s{i}" for i in range(N)], embeddings=X)
    Q = make_unit(140, D, 92) * 1.3
    c.set_path(2)
    _check(c, X, Q, 10, space)
    _check(c, X, Q[:1], 5, space)
    c.close()
