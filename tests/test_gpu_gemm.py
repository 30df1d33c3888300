"""GPU parity for K3 (tcgen05 batched scoring, path 2) against the CPU oracle, through the C ABI.
(The full-size BASELINE shapes -- 1M rows, batch 256 / 1024 -- are in test_gpu_fullsize.py, bit-exact against the C oracle.)"""
import numpy as np
import pytest

from conftest import make_unit
from test_gpu_parity import _check, _mk, _oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("space", ["cosine", "ip", "l2"])
@pytest.mark.parametrize("d", [384, 512, 768])
def test_gemm_matches_oracle(space, d):
    c, X, _ = _mk(space, d, 20000, seed=d + 1, unit=(space == "cosine"))
    Q = make_unit(40, d, 77)
    _check(c, X, Q, 5, space, path=2)
    if space == "cosine":       # unit-norm random data, k=5: the bf16 x bf16 certificate must hold without help
        assert c.stats()["n_exact_fallbacks"] == 0
    _check(c, X, Q[:9], 10, space, path=2)
    _check(c, X, Q[:17], 16, space, path=2)   # 16th vs 32nd best of 20k rows: fall-backs are legitimate here


@pytest.mark.parametrize("d", [128, 256, 1000, 1536])
def test_gemm_other_dims(d):
    """padded dims 128 / 256 (resident queries) and 1024 / 1536 (streamed query K-blocks; 1000 pads to 1024)"""
    c, X, _ = _mk("cosine", d, 6000, seed=d, unit=True)
    Q = make_unit(33, d, 5)
    _check(c, X, Q, 5, "cosine", path=2)


def test_gemm_batch_crosses_query_blocks():
    """nq = 300: three 128-query blocks, the last one ragged; n not a multiple of the 256-row tile."""
    c, X, _ = _mk("cosine", 384, 33333, seed=9)
    Q = make_unit(300, 384, 10)
    _check(c, X, Q, 5, "cosine", path=2)
    assert c.stats()["n_exact_fallbacks"] == 0


def test_gemm_mid_size_shard_planted_neighbours():
    """274 tiles (too few per CTA for the in-kernel seeding, which tests/test_gpu_seeding.py covers): results must
    equal the oracle, including for planted neighbours at the edges of slices."""
    n, d = 70001, 384
    c, X, _ = _mk("cosine", d, n, seed=19)
    Q = make_unit(24, d, 20)
    Q[:8] = X[[1, 300, 777, 12345, 40000, 65535, 69999, 70000]] + 0.03 * make_unit(8, d, 21)
    _check(c, X, Q, 5, "cosine", path=2)
    _check(c, X, Q, 16, "cosine", path=2)
    types = np.arange(n) % 3
    c2, X2, _ = _mk("l2", d, n, seed=23, unit=False)
    _check(c2, X2, Q, 5, "l2", path=2)


def test_gemm_auto_path_for_batches():
    c, X, _ = _mk("cosine", 384, 20000, seed=12)
    Q = make_unit(64, 384, 13)
    before = c.stats()["launches"]
    _check(c, X, Q, 5, "cosine")           # automatic: nq >= 5 takes K3
    used = c.stats()["launches"] - before
    assert used <= 8, used                  # prepare + pass bits + gemm + finalize + exact fix-up, not 16 scans


def test_gemm_k_up_to_32_and_tiny_corpus():
    c, X, _ = _mk("cosine", 384, 3000, seed=14)
    Q = make_unit(20, 384, 15)
    _check(c, X, Q, 20, "cosine", path=2)
    _check(c, X, Q, 32, "cosine", path=2)
    c2, X2, _ = _mk("cosine", 384, 7, seed=16)          # fewer rows than k: every row is a candidate
    _check(c2, X2, Q, 16, "cosine", path=2)


@pytest.mark.parametrize("k", [33, 50, 100, 128])
def test_gemm_pool_mode_large_k(k):
    """32 < k <= 128: no per-thread lists; sampling seeds a bound, rows above it are pooled (config 4's k=100)."""
    c, X, _ = _mk("cosine", 384, 40000, seed=31)
    Q = make_unit(40, 384, 32)
    _check(c, X, Q, k, "cosine", path=2)
    _check(c, X, Q[:3], k, "cosine", path=2)              # small batch still samples in pool mode
    assert c.stats()["n_exact_fallbacks"] == 0


def test_gemm_pool_mode_spaces_filters_and_small_shards():
    c, X, _ = _mk("l2", 512, 9000, seed=33, unit=False)
    Q = make_unit(20, 512, 34)
    _check(c, X, Q, 64, "l2", path=2)
    c2, X2, _ = _mk("ip", 256, 1500, seed=35, unit=False)  # < 8 tiles: no sampling, every row is pooled
    Q2 = make_unit(10, 256, 36)
    _check(c2, X2, Q2, 100, "ip", path=2)
    c3, X3, _ = _mk("cosine", 384, 300, seed=37)           # fewer rows than k in the filtered set
    mask = np.arange(300) % 4 == 0
    c3.delete(ids=[f"doc_{i // 7:012x}_text_{i}" for i in range(300) if not mask[i]])
    _check(c3, X3, make_unit(6, 384, 38), 100, "cosine", where_mask=mask, path=2)


def test_gemm_filters_and_tombstones():
    from multimodal_rag_b200 import B200Collection
    d, n = 512, 20000
    X = make_unit(n, d, 21)
    rng = np.random.default_rng(0x7E57)
    types = rng.choice(["text", "table", "image"], size=n, p=[0.6, 0.1, 0.3])
    metas = [{"type": str(t), "page": int(i % 13)} for i, t in enumerate(types)]
    ids = [f"r{i}" for i in range(n)]
    c = B200Collection("mm", {"hnsw:space": "cosine"})
    c.add(ids=ids, embeddings=X, metadatas=metas)
    Q = make_unit(48, d, 22)
    _check(c, X, Q, 10, "cosine", where_mask=(types == "image"), where={"type": "image"}, path=2)
    page = np.arange(n) % 13
    _check(c, X, Q, 10, "cosine", where_mask=(types == "text") & (page >= 11),
           where={"$and": [{"type": "text"}, {"page": {"$gte": 11}}]}, path=2)
    c.delete(ids=ids[:500])
    alive = np.ones(n, dtype=bool); alive[:500] = False
    Qn = X[:48] + 0.05 * make_unit(48, d, 23)             # nearest neighbours were just deleted
    _check(c, X, Qn, 10, "cosine", where_mask=alive, path=2)


def test_gemm_duplicates_fall_back_exactly():
    """Exact duplicates tie at every precision: the certificate cannot separate them, the exact
    fix-up must still produce the oracle's order."""
    from multimodal_rag_b200 import B200Collection
    d = 384
    base = make_unit(400, d, 4)
    X = np.concatenate([base, base, base[:100]])
    c = B200Collection("dup", {"hnsw:space": "cosine"})
    c.add(ids=[f"id{i}" for i in range(X.shape[0])], embeddings=X)
    Q = base[:24] + 0.01 * make_unit(24, d, 6)
    _check(c, X, Q, 8, "cosine", path=2)


def test_gemm_bf16_only_corpus():
    import torch
    c2, X2, _ = _mk("ip", 384, 8000, seed=42, keep_f32_master=False)
    Xb = torch.from_numpy(X2).to(torch.bfloat16).to(torch.float32).numpy()
    Q = make_unit(12, 384, 43)
    _check(c2, Xb, Q, 5, "ip", path=2)
