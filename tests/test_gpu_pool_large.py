"""Pool mode (32 < k <= 128, BASELINE config 4's top_k=100) at the shard sizes config 4 puts on one GPU.

Round 1's driver run fell off a cliff here (13.9 q/s on 2 GPUs): the CTAs of a query block drift apart during the long
sampling phase of a 25M+ row shard, a fixed 50 us wait for the cross-CTA fold gave up, nothing bounded the appends,
the regions overflowed and every query was redone by the exact scan.  These cases pin the repaired behaviour:
  * 25M x 384 rows, batch 1024, k=100 on one GPU: no exact fall-back, a wall-time bound, and the fast path's rows
    identical to the exact fp64 scan's (K5, itself checked against the C oracle in test_gpu_parity) on a sample;
  * a slice that posts its samples 200 us late (development knob): the fold is put off, not abandoned -- same answers,
    no fall-back;
  * no wait at all / no seed at all: every thread bounds itself from its own region -- same answers, no fall-back.
"""
import ctypes
import os

import pytest

pytestmark = pytest.mark.gpu


def _shard(rows, dim=384, seed=0xC4):
    import torch
    from multimodal_rag_b200.sharded import DeviceShard
    sh = DeviceShard(dim, "cosine", capacity=rows, row_base=0, device=0)
    g = torch.Generator(device="cuda").manual_seed(seed)
    for s0 in range(0, rows, 1 << 18):
        m = min(1 << 18, rows - s0)
        sh.ingest(torch.nn.functional.normalize(torch.randn(m, dim, generator=g, device="cuda"), dim=1))
    torch.cuda.synchronize()
    return sh


def _queries(nq, dim=384):
    import torch
    g = torch.Generator(device="cuda").manual_seed(0xBEEF4)
    return torch.nn.functional.normalize(torch.randn(nq, dim, generator=g, device="cuda"), dim=1)


def _fast_vs_exact(sh, Q, k, sample):
    """rows of the fast path for the whole batch == rows of the forced exact scan on the first `sample` queries"""
    import torch
    from multimodal_rag_b200 import _lib
    lib = _lib.load()
    o = sh.alloc_out(Q.shape[0], k)
    sh.query_local(Q, k, o)
    torch.cuda.synchronize()
    fast_rows, fast_d = o["rows"][:sample].clone(), o["d64"][:sample].clone()
    assert int(o["cnt"].min()) == k
    _lib.check(lib.b2r_set_path(sh.h, 3))
    o2 = sh.alloc_out(sample, k)
    sh.query_local(Q[:sample].contiguous(), k, o2)
    torch.cuda.synchronize()
    _lib.check(lib.b2r_set_path(sh.h, 0))
    assert torch.equal(fast_rows, o2["rows"])
    assert float((fast_d - o2["d64"]).abs().max()) < 1e-12
    return o


def test_config4_shard_on_one_gpu_25m_rows():
    import torch
    free_b, _ = torch.cuda.mem_get_info()
    rows = 25_000_000
    if free_b < rows * (384 * 6 + 16) * 1.1:
        pytest.skip("needs ~64 GB of free HBM")
    sh = _shard(rows)
    try:
        Q = _queries(1024)
        o = _fast_vs_exact(sh, Q, 100, 8)
        assert sh.fallbacks() == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            sh.query_local(Q, 100, o)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        # 2 * 1024 * 25M * 384 flop = 19.7 TFLOP: 11.8 ms at the burst peak, 14 ms sustained; the cliff was 8.6 s
        assert ms < 20.0, f"{ms:.1f} ms per batch"
        assert sh.fallbacks() == 0
    finally:
        sh.close()


@pytest.mark.parametrize("env", [{"B2R_DELAY_US": "200"}, {"B2R_DELAY_US": "200", "B2R_SEED_WAIT_NS": "20000"},
                                 {"B2R_SEED_WAIT_NS": "1"}, {"B2R_NO_SEED": "1"}],
                         ids=["late-slice", "late-slice-short-wait", "no-wait", "no-seed"])
def test_pool_mode_degrades_gracefully(env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)                      # read when the handle is created
    try:
        sh = _shard(2_000_000)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    try:
        _fast_vs_exact(sh, _queries(1024), 100, 8)
        assert sh.fallbacks() == 0
        _fast_vs_exact(sh, _queries(300), 64, 4)     # ragged last query block, another list length
        # with no seed at all and ~100 threads per query, the self-bounded regions together exceed the query's pool:
        # those queries are flagged and redone by the exact scan -- still the right answer, which is what is pinned here
        if "B2R_NO_SEED" not in env:
            assert sh.fallbacks() == 0
    finally:
        sh.close()
