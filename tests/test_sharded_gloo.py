"""CPU, world_size=2, gloo: the host-side logic of the row-sharded collection (slicing, global
sequence numbers, the all_gather exchange, merge order, payload assembly) with the oracle standing
in for each rank's GPU shard and a numpy merge standing in for the CUDA merge kernel."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShard:
    """Adapter: oracle ExactCollection with the attributes ShardedCollection uses of B200Collection."""

    def __init__(self, space):
        from oracle import exact_oracle as eo
        self.eo = eo
        self.c = eo.ExactCollection("shard", {"hnsw:space": space})

    def rows_of(self, ids):
        return np.asarray([self.c._row_of.get(i, -1) for i in ids], dtype=np.int64)

    def ids_of(self, rows):
        return [self.c._ids[int(r)] for r in rows]

    @property
    def _docs(self):
        return self.c._doc

    @property
    def _meta(self):
        class M:
            pass
        m = M()
        m.meta = self.c._meta
        return m

    def add(self, **kw):
        self.c.add(**kw)

    def upsert(self, **kw):
        self.c.upsert(**kw)

    @property
    def dimension(self):
        return self.c.dim

    def delete(self, ids=None, where=None):
        self.c.delete(ids=ids, where=where)

    def count(self):
        return self.c.count()

    def query_rows(self, q, k, where=None, want_dist64=False):
        eo = self.eo
        q = np.atleast_2d(np.asarray(q, dtype=np.float32))
        nq = q.shape[0]
        rows = np.full((nq, k), -1, dtype=np.int64); d32 = np.full((nq, k), np.inf, dtype=np.float32)
        d64 = np.full((nq, k), np.inf); cnt = np.zeros(nq, dtype=np.int32)
        live = self.c._select_rows(None, where)
        if live:
            X = np.stack([self.c._vec[r] for r in live])
            qq = eo.normalize_f32(q) if self.c.space == "cosine" else q
            dd = eo.distances_f64(qq, X, self.c.space)
            for i in range(nq):
                order = np.lexsort((np.asarray(live), dd[i]))[:k]
                cnt[i] = len(order)
                rows[i, : len(order)] = np.asarray(live)[order]
                d64[i, : len(order)] = dd[i][order]
                d32[i, : len(order)] = dd[i][order].astype(np.float32)
        return rows, d32, cnt, d64


def numpy_merge(rows, d64, cnt, k):
    R, nq, _ = rows.shape
    o_rows = torch.full((nq, k), -1, dtype=torch.int64); o_dist = torch.full((nq, k), float("inf"))
    o_cnt = torch.zeros(nq, dtype=torch.int32)
    for i in range(nq):
        cand = [(float(d64[r, i, j]), int(rows[r, i, j])) for r in range(R) for j in range(int(cnt[r, i]))]
        cand.sort()
        cand = cand[:k]
        o_cnt[i] = len(cand)
        for j, (d, g) in enumerate(cand):
            o_rows[i, j] = g; o_dist[i, j] = float(np.float32(d))
    return o_rows, o_dist, o_cnt


def _real_collection_on_the_fake_device(space):
    """The product's B200Collection (native id table, metadata columns, compiled clauses, result assembly) with
    tests/fake_device.py answering the device entry points: sharded.py and collection.py are exercised TOGETHER on CPU."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from multimodal_rag_b200 import _lib, B200Collection
    from fake_device import FakeLib
    if not isinstance(_lib.load(), FakeLib):
        lib = FakeLib()
        _lib.load = lambda: lib
    return B200Collection("shard", {"hnsw:space": space})


def _worker(rank, world, port, q, kind="oracle"):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodal_rag_b200.sharded import ShardedCollection
        from oracle import exact_oracle as eo
        make_shard = (lambda: OracleShard("cosine")) if kind == "oracle" else (lambda: _real_collection_on_the_fake_device("cosine"))
        rng = np.random.default_rng(5)
        n, d = 300, 32
        X = rng.standard_normal((n, d), dtype=np.float32)
        X[40:50] = X[10:20]                                   # duplicates that land on different shards
        ids = [f"doc_{i // 9:03d}_text_{i}" for i in range(n)]
        metas = [{"type": "image" if i % 4 == 0 else "text", "doc_id": f"doc_{i // 9:03d}"} for i in range(n)]
        docs = [f"summary {i}" for i in range(n)]
        sc = ShardedCollection("c", {"hnsw:space": "cosine"}, shard_factory=make_shard, merge_fn=numpy_merge)
        ref = eo.ExactCollection("c", {"hnsw:space": "cosine"})
        for lo in range(0, n, 77):                            # several ragged batches
            sl = slice(lo, min(n, lo + 77))
            sc.add(ids=ids[sl], embeddings=X[sl], metadatas=metas[sl], documents=docs[sl])
            ref.add(ids=ids[sl], embeddings=X[sl], metadatas=metas[sl], documents=docs[sl])
        sc.add(ids=ids[:5], embeddings=X[:5])                 # existing ids skipped on every rank
        assert sc.count() == ref.count() == n
        assert abs(sc.shard.count() - n / world) <= 4         # balanced slices
        Q = np.concatenate([X[10:13] + 0.01, rng.standard_normal((3, d), dtype=np.float32)])
        for where in (None, {"type": "image"}, {"doc_id": "doc_003"}):
            a = sc.query(query_embeddings=Q, n_results=7, where=where)
            b = ref.query(query_embeddings=Q, n_results=7, where=where)
            assert a["ids"] == b["ids"], (where, a["ids"][0], b["ids"][0])
            np.testing.assert_allclose(np.asarray(a["distances"], dtype=np.float64),
                                       np.asarray(b["distances"], dtype=np.float64), rtol=1e-6, atol=1e-7)
            assert a["metadatas"] == b["metadatas"] and a["documents"] == b["documents"]
        sc.delete(ids=ids[10:13]); ref.delete(ids=ids[10:13])
        sc.upsert(ids=[ids[100], "new_1"], embeddings=X[[12, 11]], metadatas=[{"type": "text"}, {"type": "text"}])
        ref.upsert(ids=[ids[100], "new_1"], embeddings=X[[12, 11]], metadatas=[{"type": "text"}, {"type": "text"}])
        assert sc.count() == ref.count()
        a = sc.query(query_embeddings=Q, n_results=5); b = ref.query(query_embeddings=Q, n_results=5)
        assert a["ids"] == b["ids"]
        # a bad batch is rejected on EVERY rank before any rank-local work (so nobody is left waiting in a collective), and
        # changes nothing; delete() without ids or a clause raises instead of wiping the collection
        before = sc.count()
        for bad in (dict(ids=["x1", "x2"], embeddings=X[:2, :16]),                       # wrong dimension
                    dict(ids=["x1", "x1"], embeddings=X[:2]),                            # duplicate ids
                    dict(ids=["x1", "x2"], embeddings=X[:3]),                            # length mismatch
                    dict(ids=["x1", "x2"], embeddings=X[:2], metadatas=[{"a": [1]}, {}])):   # bad metadata value
            try:
                sc.upsert(**bad)
                raise AssertionError(f"accepted {list(bad)}")
            except ValueError:
                pass
        for kw in ({}, {"ids": []}, {"where": {}}):
            try:
                sc.delete(**kw)
                raise AssertionError("delete() without arguments must raise")
            except ValueError:
                pass
        assert sc.count() == before
        # an upsert that moves an id to another rank leaves exactly one live copy
        sc.upsert(ids=[ids[5], ids[290], "new_2", "new_3"], embeddings=X[[1, 2, 3, 4]])
        ref.upsert(ids=[ids[5], ids[290], "new_2", "new_3"], embeddings=X[[1, 2, 3, 4]])
        assert sc.count() == ref.count()
        a = sc.query(query_embeddings=X[[1, 2]], n_results=3); b = ref.query(query_embeddings=X[[1, 2]], n_results=3)
        assert a["ids"] == b["ids"]
        q.put((rank, "ok"))
    except Exception as e:                                    # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("kind", ["oracle", "b200_collection_on_fake_device"])
def test_sharded_collection_world2_gloo(kind):
    from multimodal_rag_b200 import build as b2r_build
    b2r_build.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + (7 if kind == "oracle" else 13)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, kind)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}:\n{msg}"
