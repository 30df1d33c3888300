"""BASELINE.json configs 3 and 5 through the C ABI: full-size property checks plus oracle parity on a
slice the oracle finishes in seconds."""
import ctypes

import numpy as np
import pytest

from conftest import make_unit

pytestmark = pytest.mark.gpu


def test_config3_mixed_collection_type_filter_1m_x_512():
    """1M x 512-d mixed text+table+image rows, where={'type': ...}, top_k=10, batch 1 and batch 256.
    Properties: every returned row carries the requested type, results are sorted, planted neighbours
    of the requested type are found, a planted neighbour of the WRONG type is never returned.  (The bit-exact
    comparison of this shape against the fp64 oracle is in test_gpu_fullsize.py.)"""
    import torch
    from multimodal_rag_b200 import _lib
    from multimodal_rag_b200.sharded import DeviceShard
    n, d, k = 1_000_000, 512, 10
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0x7E57)
    sh = DeviceShard(d, "cosine", capacity=n, device=0)
    codes = torch.multinomial(torch.tensor([0.6, 0.1, 0.3]), n, replacement=True, generator=torch.Generator().manual_seed(1)).to(torch.uint8)
    lib = _lib.load()
    X = torch.empty((n, d), device=dev)
    for s in range(0, n, 1 << 18):
        m = min(1 << 18, n - s)
        X[s:s + m] = torch.nn.functional.normalize(torch.randn(m, d, generator=g, device=dev), dim=1)
        first = ctypes.c_int64()
        cd = codes[s:s + m].contiguous().to(dev)
        _lib.check(lib.b2r_ingest_f32(sh.h, X[s:s + m].data_ptr(), m, cd.data_ptr(), ctypes.byref(first),
                                      torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    codes_d = codes.to(dev)
    img = torch.nonzero(codes_d == 2).flatten()
    txt = torch.nonzero(codes_d == 0).flatten()
    nq = 256
    Q = torch.nn.functional.normalize(torch.randn(nq, d, generator=g, device=dev), dim=1)
    planted = img[torch.tensor([5, 1000, 99_999, 250_000], device=dev)]
    Q[:4] = torch.nn.functional.normalize(X[planted] + 0.02 * torch.randn(4, d, generator=g, device=dev), dim=1)
    wrong = txt[torch.tensor([7, 70_000], device=dev)]           # nearest rows are text: must NOT come back
    Q[4:6] = torch.nn.functional.normalize(X[wrong] + 0.02 * torch.randn(2, d, generator=g, device=dev), dim=1)
    f = _lib.B2RFilter(type_mask=1 << 2, allow_bits=None)
    for batch in (256, 1):
        rows = torch.empty((batch, k), dtype=torch.int64, device=dev)
        dist = torch.empty((batch, k), dtype=torch.float32, device=dev)
        cnt = torch.empty((batch,), dtype=torch.int32, device=dev)
        _lib.check(lib.b2r_query(sh.h, Q.data_ptr(), batch, k, ctypes.byref(f), rows.data_ptr(), dist.data_ptr(),
                                 cnt.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert (cnt == k).all()
        assert (codes_d[rows.flatten()] == 2).all()
        assert (dist[:, 1:] >= dist[:, :-1]).all()
        assert rows[:min(4, batch), 0].tolist() == planted[:min(4, batch)].tolist()
        if batch > 4:
            assert not torch.isin(rows[4:6], wrong).any()
        # (ids bit-exact against the fp64 oracle restricted to the same rows: test_gpu_fullsize.py)
    st = _lib.B2RStats()
    _lib.check(lib.b2r_get_stats(sh.h, ctypes.byref(st)))
    assert st.n_exact_fallbacks == 0
    sh.close()


def test_config5_streaming_upserts_and_queries_768():
    """Interleaved batched upserts (normalise + pack) and batch-64 queries, top_k=20, 768-d.  After
    every upsert the just-written rows must be their own top-1 (visibility on the same stream), ids
    overwritten by an upsert never come back, and the final state equals the oracle."""
    from multimodal_rag_b200 import B200Collection
    from oracle import exact_oracle as eo
    d, k, rounds, per = 768, 20, 6, 2048
    c = B200Collection("stream", {"hnsw:space": "cosine"}, capacity=rounds * per + 8)
    rng = np.random.default_rng(5)
    live = {}                                                  # id -> vector (the oracle's view)
    order = []                                                 # insertion order of live ids (row order)
    for r in range(rounds):
        X = make_unit(per, d, 100 + r) * rng.uniform(0.5, 2.0, size=(per, 1)).astype(np.float32)
        ids = [f"doc_{r:02d}_{i}" for i in range(per)]
        if r > 0:                                              # ~10 % of the batch overwrites existing ids
            old = rng.choice(len(order), size=per // 10, replace=False)
            for j, o in enumerate(old):
                ids[j] = order[o]
        c.upsert(ids=ids, embeddings=X)
        for i, v in zip(ids, X):
            if i in live:
                order.remove(i)
            live[i] = v
            order.append(i)
        probe = rng.choice(per, size=64, replace=False)
        res = c.query(query_embeddings=X[probe], n_results=k, include=["distances"])
        assert [x[0] for x in res["ids"]] == [ids[p] for p in probe]          # visibility
        assert all(abs(x[0]) < 1e-5 for x in res["distances"])
        assert c.count() == len(live)
    # final state against the oracle (rows in current insertion order; ties impossible on random data)
    Xl = np.stack([live[i] for i in order])
    Q = make_unit(64, d, 999)
    res = c.query(query_embeddings=Q, n_results=k, include=["distances"])
    er, ed = eo.topk_exact(eo.normalize_f32(Q), eo.normalize_f32(Xl), k, "cosine")
    for i in range(64):
        assert res["ids"][i] == [order[j] for j in er[i]]
        np.testing.assert_allclose(res["distances"][i], ed[i], rtol=1e-5, atol=1e-7)
