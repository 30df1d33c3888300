"""BASELINE.json's configs at FULL size, ids bit-exact against the fp64 C oracle (oracle/exact_topk.c, acc64) on a sample
of the batch, distances to north_star's 1e-5 relative.  The oracle scans 1M rows for 32-48 queries in a few seconds on the
GPU box's host cores, so nothing here needs a tolerance on the ids:

  configs[1]  1M x 384, batch 256, top_k 5             (list mode, CTA pairs)
  configs[2]  1M x 512, type filter, batch 256 and 1, top_k 10
  configs[3]  1M x 384 per GPU, batch 1024, top_k 100  (pool mode)
  configs[4]  1M x 768, top_k 20 after interleaved upserts and deletes (tombstones, appended rows)

and the paired (cta_group::2) and unpaired forms of K3 must return identical rows and distances.
"""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _corpus(n, d, seed, dev="cuda"):
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    X = torch.empty((n, d), device=dev)
    for s in range(0, n, 1 << 18):
        m = min(1 << 18, n - s)
        X[s:s + m] = torch.nn.functional.normalize(torch.randn(m, d, generator=g, device=dev), dim=1)
    return X, g


def _oracle_check(rows, dist, cnt, Xh, Qh, k, sample, allowed=None):
    """rows/dist/cnt: the engine's answers (numpy) for the whole batch; the first `sample` queries are checked."""
    from oracle import c_oracle
    c_oracle.set_threads(0)
    Xs, Qs = c_oracle.normalize_f32(Xh), c_oracle.normalize_f32(Qh[:sample])
    er, ed, ec = c_oracle.topk(Xs, Qs, k, "cosine", allowed=allowed, acc64=True)
    np.testing.assert_array_equal(cnt[:sample], ec)
    np.testing.assert_array_equal(rows[:sample], er)                      # bit-exact ids, in order
    np.testing.assert_allclose(dist[:sample], ed, rtol=1e-5, atol=1e-7)   # north_star: 1e-5 relative


def _shard_query(sh, Q, k, f=None):
    import torch
    from multimodal_rag_b200 import _lib
    nq = Q.shape[0]
    rows = torch.empty((nq, k), dtype=torch.int64, device=Q.device)
    dist = torch.empty((nq, k), dtype=torch.float32, device=Q.device)
    cnt = torch.empty((nq,), dtype=torch.int32, device=Q.device)
    _lib.check(sh.lib.b2r_query(sh.h, Q.data_ptr(), nq, k, None if f is None else ctypes.byref(f), rows.data_ptr(),
                                dist.data_ptr(), cnt.data_ptr(), torch.cuda.current_stream().cuda_stream), "b2r_query")
    torch.cuda.synchronize()
    return rows.cpu().numpy(), dist.cpu().numpy(), cnt.cpu().numpy()


def _shard(X, codes=None, env=None):
    import torch
    from multimodal_rag_b200 import _lib
    from multimodal_rag_b200.sharded import DeviceShard
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        sh = DeviceShard(X.shape[1], "cosine", capacity=X.shape[0], device=0, world=1)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for s in range(0, X.shape[0], 1 << 18):
        m = min(1 << 18, X.shape[0] - s)
        first = ctypes.c_int64()
        cd = None if codes is None else codes[s:s + m].contiguous()
        _lib.check(sh.lib.b2r_ingest_f32(sh.h, X[s:s + m].data_ptr(), m, None if cd is None else cd.data_ptr(),
                                         ctypes.byref(first), torch.cuda.current_stream().cuda_stream), "b2r_ingest_f32")
    torch.cuda.synchronize()
    return sh


def test_config2_1m_x_384_batch256_k5_bit_exact():
    import torch
    X, g = _corpus(1_000_000, 384, 0xC0FFEE)
    Q = torch.nn.functional.normalize(torch.randn(256, 384, generator=g, device="cuda"), dim=1)
    Q[:4] = torch.nn.functional.normalize(X[[3, 499_999, 777_777, 999_999]] + 0.02 * torch.randn(4, 384, generator=g, device="cuda"), dim=1)
    sh = _shard(X)
    rows, dist, cnt = _shard_query(sh, Q, 5)
    assert rows[:4, 0].tolist() == [3, 499_999, 777_777, 999_999]
    Xh, Qh = X.cpu().numpy(), Q.cpu().numpy()
    _oracle_check(rows, dist, cnt, Xh, Qh, 5, 48)
    assert sh.fallbacks() == 0
    # batch 1 (the HBM-bound scan) and the warp-shuffle scan path give the same answers as the batch did
    r1, d1, c1 = _shard_query(sh, Q[:1].contiguous(), 5)
    np.testing.assert_array_equal(r1[0], rows[0]); np.testing.assert_array_equal(d1[0], dist[0])
    from multimodal_rag_b200 import _lib
    _lib.check(sh.lib.b2r_set_path(sh.h, 1))
    r2, d2, _ = _shard_query(sh, Q[:16].contiguous(), 5)
    np.testing.assert_array_equal(r2, rows[:16]); np.testing.assert_array_equal(d2, dist[:16])
    sh.close()
    # the unpaired form of K3 (one CTA per query block) must agree bit for bit with the paired one
    sh2 = _shard(X, env={"B2R_NO_PAIR": "1"})
    rows2, dist2, cnt2 = _shard_query(sh2, Q, 5)
    np.testing.assert_array_equal(rows2, rows); np.testing.assert_array_equal(dist2, dist); np.testing.assert_array_equal(cnt2, cnt)
    sh2.close()


def test_config3_1m_x_512_type_filter_k10_bit_exact():
    import torch
    from multimodal_rag_b200 import _lib
    X, g = _corpus(1_000_000, 512, 0x7E57)
    codes = torch.multinomial(torch.tensor([0.6, 0.1, 0.3]), 1_000_000, replacement=True,
                              generator=torch.Generator().manual_seed(1)).to(torch.uint8).cuda()
    sh = _shard(X, codes)
    Q = torch.nn.functional.normalize(torch.randn(256, 512, generator=g, device="cuda"), dim=1)
    img = torch.nonzero(codes == 2).flatten()
    txt = torch.nonzero(codes == 0).flatten()
    Q[:2] = torch.nn.functional.normalize(X[img[[5, 250_000]]] + 0.02 * torch.randn(2, 512, generator=g, device="cuda"), dim=1)
    Q[2:4] = torch.nn.functional.normalize(X[txt[[7, 70_000]]] + 0.02 * torch.randn(2, 512, generator=g, device="cuda"), dim=1)   # wrong type
    f = _lib.B2RFilter(type_mask=1 << 2, allow_bits=None)
    Xh, Qh = X.cpu().numpy(), Q.cpu().numpy()
    allowed = (codes == 2).cpu().numpy()
    for batch, sample in ((256, 32), (1, 1)):
        rows, dist, cnt = _shard_query(sh, Q[:batch].contiguous(), 10, f)
        assert allowed[rows.ravel()].all()                                # every returned row carries the requested type
        _oracle_check(rows, dist, cnt, Xh, Qh, 10, sample, allowed=allowed)
    assert sh.fallbacks() == 0
    sh.close()


def test_config4_shard_1m_x_384_batch1024_k100_bit_exact():
    import torch
    X, g = _corpus(1_000_000, 384, 0xC4)
    Q = torch.nn.functional.normalize(torch.randn(1024, 384, generator=g, device="cuda"), dim=1)
    sh = _shard(X)
    rows, dist, cnt = _shard_query(sh, Q, 100)
    assert (np.diff(dist, axis=1) >= 0).all()
    _oracle_check(rows, dist, cnt, X.cpu().numpy(), Q.cpu().numpy(), 100, 32)
    assert sh.fallbacks() == 0
    st = _stats(sh)
    assert 300 < st.n_pool_entries / st.n_pool_queries < 4000
    sh.close()


def _stats(sh):
    from multimodal_rag_b200 import _lib
    st = _lib.B2RStats()
    _lib.check(sh.lib.b2r_get_stats(sh.h, ctypes.byref(st)))
    return st


def test_config5_1m_x_768_k20_after_interleaved_upserts_bit_exact():
    """Upserts = tombstone + append: after 6 rounds of {overwrite 10 % of a batch, append 8192 rows, query} the answer over
    the surviving rows must be the oracle's over the same rows (dead rows excluded through `allowed`)."""
    import torch
    from multimodal_rag_b200 import _lib
    n0, d, k, up = 1_000_000, 768, 20, 8192
    X0, g = _corpus(n0, d, 0xC5)
    from multimodal_rag_b200.sharded import DeviceShard
    sh = DeviceShard(d, "cosine", capacity=n0 + 6 * up, device=0, world=1)
    st = torch.cuda.current_stream().cuda_stream
    for s in range(0, n0, 1 << 18):
        sh.ingest(X0[s:s + (1 << 18)])
    rng = np.random.default_rng(5)
    allX = [X0]
    dead = np.zeros(n0 + 6 * up, dtype=bool)
    for r in range(6):
        old = rng.integers(0, n0 + r * up, size=up // 10, dtype=np.int64)
        _lib.check(sh.lib.b2r_tombstone(sh.h, old.ctypes.data, old.shape[0], st), "b2r_tombstone")
        dead[old] = True
        U = torch.nn.functional.normalize(torch.randn(up, d, generator=g, device="cuda"), dim=1)
        first = sh.ingest(U)
        assert first == n0 + r * up
        allX.append(U)
        probe = U[:64].contiguous()
        rows, dist, cnt = _shard_query(sh, probe, k)
        assert rows[:, 0].tolist() == list(range(first, first + 64))       # visible to the next query, its own nearest neighbour
        assert (np.abs(dist[:, 0]) < 1e-5).all()
        assert not dead[rows.ravel()].any()                                # overwritten rows never come back
    Xall = torch.cat(allX).cpu().numpy()
    Q = torch.nn.functional.normalize(torch.randn(64, d, generator=g, device="cuda"), dim=1)
    Q[:8] = torch.nn.functional.normalize(torch.from_numpy(Xall[np.flatnonzero(dead)[:8]]).cuda() + 0.01 * torch.randn(8, d, generator=g, device="cuda"), dim=1)   # next to dead rows
    rows, dist, cnt = _shard_query(sh, Q, k)
    _oracle_check(rows, dist, cnt, Xall, Q.cpu().numpy(), k, 32, allowed=~dead[: Xall.shape[0]])
    assert sh.fallbacks() == 0
    sh.close()
