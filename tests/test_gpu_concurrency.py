"""Two collections queried from two threads at once (the reference calls the store from asyncio.to_thread workers,
app/utils/embedder.py:517,595,809-815).  K3's in-kernel threshold seeding makes CTAs of one launch wait for each other;
when two launches share the GPU that wait must stay bounded and the answers must not change."""
import threading

import numpy as np
import pytest

from conftest import make_unit

pytestmark = pytest.mark.gpu


def test_two_collections_two_threads():
    import torch
    from multimodal_rag_b200 import B200Collection
    from oracle import exact_oracle as eo
    n, d, k = 300_000, 384, 5
    cols, data = [], []
    for seed in (1, 2):
        X = make_unit(n, d, seed)
        c = B200Collection(f"c{seed}", {"hnsw:space": "cosine"}, capacity=n)
        c.add(ids=[f"r{i}" for i in range(n)], embeddings=X)
        cols.append(c)
        data.append(X)
    Q = make_unit(64, d, 9)
    want = [eo.topk_exact(eo.normalize_f32(Q), eo.normalize_f32(X), k, "cosine") for X in data]
    errors = []

    def worker(ci):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                qd = torch.as_tensor(Q, device="cuda")
                for it in range(25):
                    nq = (1, 7, 64)[it % 3]
                    rows, dist, cnt = cols[ci].query_rows(qd[:nq], k)
                    er, ed = want[ci]
                    for i in range(nq):
                        np.testing.assert_array_equal(rows[i, : cnt[i]], er[i])
                        np.testing.assert_allclose(dist[i, : cnt[i]], ed[i], rtol=1e-5, atol=1e-7)
        except Exception as e:                                   # noqa: BLE001
            errors.append((ci, repr(e)))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in ts), "a query thread hung"
    assert not errors, errors
    for c in cols:
        assert c.stats()["n_exact_fallbacks"] == 0
        c.close()
