"""Randomised parity sweep (SURVEY.md §4 iii): random shapes, spaces, k, batch sizes, filters, tombstones and
adversarial value patterns through every scoring path, each compared with the fp64 oracle -- ids bit-exact,
distances to 1e-5 relative.  Seeds are fixed so a failure reproduces."""
import numpy as np
import pytest

from test_gpu_parity import _check

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(seed)
    d = int(rng.choice([64, 100, 128, 200, 256, 384, 512, 768, 1024]))
    n = int(rng.choice([1, 5, 33, 257, 1000, 4097, 12000]))
    nq = int(rng.choice([1, 2, 3, 5, 9, 64, 130]))
    k = int(rng.choice([1, 3, 5, 8, 9, 16, 20, 32, 33, 64, 100, 128]))
    space = str(rng.choice(["cosine", "l2", "ip"]))
    kind = str(rng.choice(["gauss", "clustered", "duplicates", "lowrank", "tiny_scale"]))
    X = rng.standard_normal((n, d)).astype(np.float32)
    if kind == "clustered":                      # tight clusters: near-ties everywhere
        centres = rng.standard_normal((max(1, n // 50), d)).astype(np.float32)
        X = centres[rng.integers(0, len(centres), n)] + 0.01 * X
    elif kind == "duplicates":                   # exact duplicates: exact ties -> lowest row must win
        X = X[rng.integers(0, max(1, n // 3), n)]
    elif kind == "lowrank":
        X = (rng.standard_normal((n, 4)).astype(np.float32) @ rng.standard_normal((4, d)).astype(np.float32))
    elif kind == "tiny_scale":
        X = X * 1e-3
    if space != "cosine":
        X = X * rng.uniform(0.2, 3.0, size=(n, 1)).astype(np.float32)
    X[np.abs(X).sum(axis=1) == 0] = 1.0           # no all-zero rows (cosine of a zero vector is undefined upstream)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    if n > 3 and rng.random() < 0.5:
        Q[: min(nq, n)] = X[rng.integers(0, n, min(nq, n))] + 0.001 * Q[: min(nq, n)]
    return rng, d, n, nq, k, space, kind, X, Q


@pytest.mark.parametrize("seed", range(40))
def test_random_case_all_paths(seed):
    from multimodal_rag_b200 import B200Collection
    rng, d, n, nq, k, space, kind, X, Q = _case(1000 + seed)
    c = B200Collection("p", {"hnsw:space": space}, keep_f32_master=bool(rng.random() < 0.8))
    types = rng.choice(["text", "table", "image"], size=n)
    ids = [f"doc_{i:06d}" for i in range(n)]
    # ingest in ragged batches (exercises capacity growth and the pass-bitmap cache)
    cuts = sorted(set([0, n] + rng.integers(0, n + 1, 3).tolist()))
    for a, b in zip(cuts[:-1], cuts[1:]):
        c.add(ids=ids[a:b], embeddings=X[a:b], metadatas=[{"type": str(t), "page": int(i % 7)} for i, t in zip(range(a, b), types[a:b])])
    Xs = X
    if not c._flags == 0:                         # bf16-only corpus: the stored rows are the bf16 roundings
        import torch
        from oracle import exact_oracle as eo
        Xn = eo.normalize_f32(X) if space == "cosine" else X
        Xs = torch.from_numpy(Xn).to(torch.bfloat16).to(torch.float32).numpy()
    alive = np.ones(n, dtype=bool)
    if n > 4 and rng.random() < 0.5:
        dead = rng.choice(n, size=max(1, n // 5), replace=False)
        c.delete(ids=[ids[i] for i in dead])
        alive[dead] = False
    where, mask = None, alive
    r = rng.random()
    if r < 0.3:
        where, mask = {"type": "image"}, alive & (types == "image")
    elif r < 0.5:
        where, mask = {"$or": [{"type": "table"}, {"page": {"$gte": 5}}]}, alive & ((types == "table") | (np.arange(n) % 7 >= 5))
    for path in (0, 1, 2, 3):
        _check_stored(c, Xs, Q, k, space, mask, where, path, normalised=(c._flags != 0 and space == "cosine"))


def _check_stored(c, Xs, Q, k, space, mask, where, path, normalised):
    """_check, except that for a bf16-only cosine corpus Xs is already the stored (normalised, rounded) rows"""
    if not normalised:
        return _check(c, Xs, Q, k, space, where_mask=mask, where=where, path=path)
    from oracle import exact_oracle as eo
    c.set_path(path)
    rows, dist, cnt = c.query_rows(Q, k, where)
    er, ed = eo.topk_exact(eo.normalize_f32(Q), Xs, k, "ip", allowed=mask)     # stored rows are used as they are
    for i in range(Q.shape[0]):
        assert cnt[i] == len(er[i])
        np.testing.assert_array_equal(rows[i, : cnt[i]], er[i])
        np.testing.assert_allclose(dist[i, : cnt[i]], ed[i], rtol=1e-5, atol=1e-7)
