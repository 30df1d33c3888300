"""b2r_query_async / b2r_wait: the pipelined form of the query call must deliver exactly what the blocking call does,
with pinned and pageable host arrays, with more calls enqueued than there are slots, and whatever the wait order."""
import ctypes

import numpy as np
import pytest

from conftest import make_unit

pytestmark = pytest.mark.gpu


def test_pipelined_queries_equal_blocking_queries():
    import torch
    from multimodal_rag_b200 import B200Collection, _lib
    lib = _lib.load()
    n, d, k, nq = 120_000, 384, 7, 48
    X = make_unit(n, d, 3)
    c = B200Collection("a", {"hnsw:space": "cosine"}, capacity=n)
    c.add(ids=[f"r{i}" for i in range(n)], embeddings=X, metadatas=[{"type": "image" if i % 4 == 0 else "text"} for i in range(n)])
    batches = [make_unit(nq, d, 100 + i) for i in range(6)]
    want = [c.query_rows(b, k) for b in batches]
    want_img = c.query_rows(batches[0], k, {"type": "image"})
    stream = torch.cuda.current_stream().cuda_stream

    def outs(pinned):
        r, dd, cc = torch.empty((nq, k), dtype=torch.int64), torch.empty((nq, k), dtype=torch.float32), torch.empty((nq,), dtype=torch.int32)
        return (r.pin_memory(), dd.pin_memory(), cc.pin_memory()) if pinned else (r, dd, cc)

    def enqueue(b, o, f=None):
        t = ctypes.c_uint64()
        q = b if isinstance(b, torch.Tensor) else torch.from_numpy(b)
        keep.append(q)
        _lib.check(lib.b2r_query_async(c.handle, q.data_ptr(), nq, k, None if f is None else ctypes.byref(f), o[0].data_ptr(),
                                       o[1].data_ptr(), o[2].data_ptr(), stream, ctypes.byref(t)), "b2r_query_async")
        return t.value

    keep = []
    O = [outs(pinned=(i % 2 == 0)) for i in range(6)]
    # six calls back to back: only two slots exist, so every third call completes the oldest one itself
    tickets = [enqueue(torch.from_numpy(b).pin_memory() if i % 3 == 0 else b, O[i]) for i, b in enumerate(batches)]
    for t in reversed(tickets):                      # any wait order, including tickets that were already delivered
        _lib.check(lib.b2r_wait(c.handle, t), "b2r_wait")
    _lib.check(lib.b2r_wait(c.handle, tickets[0]), "b2r_wait")
    for (r, dd, cc), (wr, wd, wc) in zip(O, want):
        np.testing.assert_array_equal(r.numpy(), wr)
        np.testing.assert_array_equal(dd.numpy().view(np.uint32), wd.view(np.uint32))
        np.testing.assert_array_equal(cc.numpy(), wc)
    # a type-mask filter rides along; clause / host-bitmap filters are refused (they would need host staging)
    f = _lib.B2RFilter(type_mask=c._meta.type_only_mask({"type": "image"}), allow_bits=None, where=None)
    o = outs(True)
    _lib.check(lib.b2r_wait(c.handle, enqueue(batches[0], o, f)), "b2r_wait")
    np.testing.assert_array_equal(o[0].numpy(), want_img[0])
    bits = np.zeros((n + 31) // 32, dtype=np.uint32)
    fbad = _lib.B2RFilter(type_mask=(1 << 64) - 1, allow_bits=bits.ctypes.data, where=None)
    t = ctypes.c_uint64()
    qh = torch.from_numpy(batches[0])
    assert lib.b2r_query_async(c.handle, qh.data_ptr(), nq, k, ctypes.byref(fbad), o[0].data_ptr(), o[1].data_ptr(),
                               o[2].data_ptr(), stream, ctypes.byref(t)) == _lib.B2R_EINVAL
    qd = qh.cuda()
    assert lib.b2r_query_async(c.handle, qd.data_ptr(), nq, k, None, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                               stream, ctypes.byref(t)) == _lib.B2R_EINVAL
    c.close()


def test_collection_pipelined_batches():
    from multimodal_rag_b200 import B200Collection
    n, d = 50_000, 384
    X = make_unit(n, d, 5)
    c = B200Collection("p", {"hnsw:space": "l2"}, capacity=n)
    c.add(ids=[f"r{i}" for i in range(n)], embeddings=X, metadatas=[{"type": "image" if i % 3 else "text", "page": i % 5} for i in range(n)])
    batches = [make_unit(nq, d, 40 + nq) for nq in (1, 17, 130, 64, 5)]
    for where in (None, {"type": "text"}, {"page": {"$gte": 3}}):
        got = c.query_rows_pipelined(batches, 6, where)
        for b, (r, dd, cc) in zip(batches, got):
            wr, wd, wc = c.query_rows(b, 6, where)
            np.testing.assert_array_equal(r, wr)
            np.testing.assert_array_equal(dd.view(np.uint32), wd.view(np.uint32))
            np.testing.assert_array_equal(cc, wc)
    c.close()
