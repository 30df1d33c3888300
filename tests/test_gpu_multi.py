"""Row-sharded path on real GPUs over NCCL (needs >= 2 GPUs; skipped on the 1-GPU box).
Every rank owns a shard; the merged answer must equal the oracle's answer over the union."""
import os
import socket

import numpy as np
import pytest

from conftest import make_unit

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, nq, k, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from multimodal_rag_b200.sharded import DeviceShard, ShardedCollection
        X = make_unit(n, d, 5)
        Q = make_unit(nq, d, 6)
        per = n // world
        lo, hi = rank * per, (n if rank == world - 1 else (rank + 1) * per)
        # (1) device-resident shard: local exact top-k -> all_gather -> merge kernel, global row numbers
        sh = DeviceShard(d, "cosine", capacity=hi - lo, row_base=lo, device=rank)
        sh.ingest(torch.from_numpy(X[lo:hi]).cuda())
        o = sh.alloc_out(nq, k)
        rows, dist_, cnt = sh.query_device(torch.from_numpy(Q).cuda(), k, o)
        torch.cuda.synchronize()
        res = {"rows": rows.cpu().numpy(), "dist": dist_.cpu().numpy(), "cnt": cnt.cpu().numpy()}
        # the pipelined form (exchange of batch i on a side stream under the scan of batch i+1), two output sets
        os_ = [sh.alloc_out(nq, k) for _ in range(2)]
        Qd = torch.from_numpy(Q).cuda()
        for i in range(5):
            sh.query_device_pipelined(Qd, k, os_[i % 2])
        sh.drain()
        torch.cuda.synchronize()
        res["pipelined_same"] = all(bool(torch.equal(o_["m_rows"], rows)) and bool(torch.equal(o_["m_dist"], dist_)) for o_ in os_)
        # pool mode (top_k = 100) across the shards, a batch of two query blocks (CTA pairs)
        Q2 = torch.from_numpy(make_unit(130, d, 8)).cuda()
        o2 = sh.alloc_out(130, 100)
        r2, d2, c2 = sh.query_device(Q2, 100, o2)
        torch.cuda.synchronize()
        res["rows100"], res["dist100"], res["cnt100"] = r2.cpu().numpy(), d2.cpu().numpy(), c2.cpu().numpy()
        # the library's own exchange (peer-to-peer stores into the peers' mailboxes + flags + merge in ONE kernel, no NCCL on
        # the data path) must give what the all_gather + merge gave -- several calls in a row (both slots, the slot
        # hand-back), a different batch size in between, and the pipelined form
        sh.enable_p2p_exchange(nq_max=256, k_max=128)
        ok = True
        for rep in range(5):
            o3 = sh.alloc_out(nq, k)
            r3, d3, c3 = sh.query_device(Qd, k, o3)
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(r3, rows)) and bool(torch.equal(d3, dist_)) and bool(torch.equal(c3, cnt))
            if rep == 2:
                o4 = sh.alloc_out(130, 100)
                r4, d4, c4 = sh.query_device(Q2, 100, o4)
                torch.cuda.synchronize()
                ok = ok and bool(torch.equal(r4, r2)) and bool(torch.equal(d4, d2))
        for i in range(6):
            sh.query_device_pipelined(Qd, k, os_[i % 2])
        sh.drain()
        torch.cuda.synchronize()
        ok = ok and all(bool(torch.equal(o_["m_rows"], rows)) and bool(torch.equal(o_["m_dist"], dist_)) for o_ in os_)
        res["p2p_same"] = ok
        # the fused form (b2r_query_push): the query's own kernels store the lists into the peers' mailboxes, the merge of a batch
        # rides behind the next batch -- 9 batches in a row (every slot twice, the ack wait), k = 100 with forced exact fix-ups in
        # between (the fix-up kernel pushes the redone queries), then the old push / merge form again on the same mailboxes
        of = [sh.alloc_out(nq, k) for _ in range(2)]
        ok = True
        for i in range(9):
            sh.query_device_fused(Qd, k, of[i % 2])
            if i >= 1:
                torch.cuda.synchronize()
                o_ = of[(i - 1) % 2]
                ok = ok and bool(torch.equal(o_["m_rows"], rows)) and bool(torch.equal(o_["m_dist"], dist_)) and bool(torch.equal(o_["m_cnt"], cnt))
        sh.drain()
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(of[0]["m_rows"], rows)) and bool(torch.equal(of[0]["m_dist"], dist_))
        o5 = sh.alloc_out(130, 100)
        sh.lib.b2r_set_path(sh.h, 3)                       # exact scan for everything: K5 emits (and pushes) every list
        sh.query_device_fused(Q2, 100, o5)
        sh.lib.b2r_set_path(sh.h, 0)
        sh.query_device_fused(Qd, k, of[1])
        sh.drain()
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(o5["m_rows"], r2)) and bool(torch.equal(o5["m_dist"], d2)) and bool(torch.equal(of[1]["m_rows"], rows))
        o6 = sh.alloc_out(nq, k)
        r6, d6, c6 = sh.query_device(Qd, k, o6)           # push / merge kernels after fused batches
        torch.cuda.synchronize()
        res["fused_same"] = ok and bool(torch.equal(r6, rows)) and bool(torch.equal(d6, dist_))
        sh.close()
        # (2) the Chroma-shaped collective collection with ids and a where clause
        sc = ShardedCollection("mm", {"hnsw:space": "cosine"}, device=rank)
        ids = [f"doc_{i:06d}_text_{i}" for i in range(n)]
        metas = [{"type": "image" if i % 3 == 0 else "text"} for i in range(n)]
        sc.add(ids=ids, embeddings=X, metadatas=metas)
        r = sc.query(query_embeddings=Q[:4], n_results=k, where={"type": "image"})
        res["ids"] = r["ids"]
        res["count"] = sc.count()
        if rank == 0:
            np.save(out, res, allow_pickle=True)
    finally:
        dist.destroy_process_group()


def test_two_gpu_row_sharded_matches_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import exact_oracle as eo
    n, d, nq, k, world = 30000, 384, 40, 5, 2
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, _free_port(), n, d, nq, k, out), nprocs=world, join=True)
    res = np.load(out, allow_pickle=True).item()
    X, Q = make_unit(n, d, 5), make_unit(nq, d, 6)
    er, ed = eo.topk_exact(eo.normalize_f32(Q), eo.normalize_f32(X), k, "cosine")
    for i in range(nq):
        np.testing.assert_array_equal(res["rows"][i, : res["cnt"][i]], er[i])
        np.testing.assert_allclose(res["dist"][i, : res["cnt"][i]], ed[i], rtol=1e-5, atol=1e-7)
    assert res["pipelined_same"] and res["p2p_same"] and res["fused_same"]
    er3, ed3 = eo.topk_exact(eo.normalize_f32(make_unit(130, d, 8)), eo.normalize_f32(X), 100, "cosine")
    for i in range(130):
        np.testing.assert_array_equal(res["rows100"][i, : res["cnt100"][i]], er3[i])
        np.testing.assert_allclose(res["dist100"][i, : res["cnt100"][i]], ed3[i], rtol=1e-5, atol=1e-7)
    mask = (np.arange(n) % 3 == 0)
    er2, _ = eo.topk_exact(eo.normalize_f32(Q[:4]), eo.normalize_f32(X), k, "cosine", allowed=mask)
    for i in range(4):
        assert res["ids"][i] == [f"doc_{r:06d}_text_{r}" for r in er2[i]]
    assert res["count"] == n


def test_fused_exchange_with_a_world_of_one():
    """b2r_query_push on one GPU (a world of one rank: the mailbox is local): the pushed lists, merged, are the query's own
    answer -- through K3's finalize, through the batch-1 scan, through forced exact scans; merged as a rider of the next call's
    last kernel or by b2r_xchg_merge; 13 batches of three shapes (every slot several times, the acks)."""
    import ctypes
    import torch
    from multimodal_rag_b200 import _lib
    from multimodal_rag_b200.sharded import DeviceShard
    lib = _lib.load()
    n, d = 50_000, 128
    sh = DeviceShard(d, "cosine", capacity=n, row_base=1000, device=0, world=1)
    sh.ingest(torch.from_numpy(make_unit(n, d, 21)).cuda())
    x = ctypes.c_void_p()
    _lib.check(lib.b2r_xchg_create(0, 0, 1, 64, 32, ctypes.byref(x)))
    st = torch.cuda.current_stream().cuda_stream
    pending, checked = [], 0

    def check(o_, want_):
        torch.cuda.synchronize()
        assert torch.equal(o_["m_rows"], want_[0]) and torch.equal(o_["m_dist"], want_[1]) and torch.equal(o_["m_cnt"], want_[2])
        assert torch.equal(o_["rows"], want_[0])                   # the local outputs are written as before

    def merge_oldest():
        nq_, k_, o_, want_ = pending.pop(0)
        _lib.check(lib.b2r_xchg_merge(x, nq_, k_, o_["m_rows"].data_ptr(), o_["m_dist"].data_ptr(), o_["m_cnt"].data_ptr(), st))
        check(o_, want_)

    for i in range(13):
        nq, k = (1, 5) if i % 3 == 0 else (40, 10) if i % 3 == 1 else (64, 32)
        Q = torch.from_numpy(make_unit(nq, d, 30 + i)).cuda()
        o = sh.alloc_out(nq, k)
        sh.query_local(Q, k, o)                                    # the plain answer
        want = (o["rows"].clone(), o["dist"].clone(), o["cnt"].clone())
        o2 = sh.alloc_out(nq, k)
        if i in (4, 8):
            lib.b2r_set_path(sh.h, 3)                              # the exact scan emits (and pushes) every list
        ride = pending[0] if (pending and i % 4 != 2) else None    # most calls carry the oldest unmerged batch as a rider
        m = ride[2] if ride else None
        _lib.check(lib.b2r_query_push(sh.h, x, Q.data_ptr(), nq, k, None, o2["rows"].data_ptr(), o2["dist"].data_ptr(),
                                      o2["cnt"].data_ptr(), m["m_rows"].data_ptr() if m else None, m["m_dist"].data_ptr() if m else None,
                                      m["m_cnt"].data_ptr() if m else None, st), "b2r_query_push")
        lib.b2r_set_path(sh.h, 0)
        if ride:
            pending.pop(0)
            check(ride[2], ride[3])
            checked += 1
        pending.append((nq, k, o2, want))
        if len(pending) == 3:
            merge_oldest()
    while pending:
        merge_oldest()
    assert checked >= 8
    # a wrong shape for the oldest batch is refused; so is a fifth unmerged batch (no deadlock) and a rider when nothing older waits
    Q = torch.from_numpy(make_unit(8, d, 99)).cuda()
    outs = [sh.alloc_out(8, 5) for _ in range(5)]
    push = lambda o_, m=None: lib.b2r_query_push(sh.h, x, Q.data_ptr(), 8, 5, None, o_["rows"].data_ptr(), o_["dist"].data_ptr(),
                                                 o_["cnt"].data_ptr(), m["m_rows"].data_ptr() if m else None,
                                                 m["m_dist"].data_ptr() if m else None, m["m_cnt"].data_ptr() if m else None, st)
    assert push(outs[0], outs[4]) == _lib.B2R_EINVAL               # nothing older to carry
    assert [push(o_) for o_ in outs] == [0, 0, 0, 0, _lib.B2R_EINVAL]
    o_ = outs[0]
    assert lib.b2r_xchg_merge(x, 9, 5, o_["m_rows"].data_ptr(), o_["m_dist"].data_ptr(), o_["m_cnt"].data_ptr(), st) == _lib.B2R_EINVAL
    torch.cuda.synchronize()
    lib.b2r_xchg_destroy(x)
    sh.close()
