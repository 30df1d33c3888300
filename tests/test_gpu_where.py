"""The device where-engine: compiled clauses (include/b2r.h b2r_where) evaluated by where_bits_kernel against the
dictionary-encoded metadata columns must give, bit for bit, the rows Chroma's grammar selects
(oracle.exact_oracle.where_match, the per-row restatement) -- and queries restricted by them must match the oracle."""
import numpy as np
import pytest

from conftest import make_unit

pytestmark = pytest.mark.gpu

CLAUSES = [
    {"type": {"$ne": "text"}}, {"type": {"$nin": ["image"]}}, {"doc_id": "doc_0007"}, {"page": {"$gte": 5}},
    {"page": {"$lt": 3}}, {"page": 4}, {"page": 4.0}, {"score": {"$gt": 20.5}}, {"flag": True}, {"flag": {"$ne": True}},
    {"page": {"$ne": "4"}}, {"missing": "x"}, {"missing": {"$ne": "x"}}, {"type": "video"},
    {"$and": [{"type": "text"}, {"page": {"$lte": 2}}]},
    {"$or": [{"type": "image"}, {"$and": [{"doc_id": "doc_0003"}, {"page": {"$in": [1, 2, 3]}}]}]},
    {"page": {"$in": [1, "1", True]}},
    {"$and": [{"doc_id": {"$in": ["doc_0001", "doc_0002", "doc_0009"]}}, {"type": {"$ne": "table"}}, {"page": {"$gte": 1}}]},
    {"$or": [{"missing": 1}, {"flag": False}]},
]


def _metas(n, seed=0):
    rng = np.random.default_rng(seed)
    metas = []
    for i in range(n):
        m = {"doc_id": f"doc_{i % 17:04d}", "item_id": f"text_{i}", "type": str(rng.choice(["text", "table", "image"]))}
        if i % 3:
            m["page"] = int(i % 11)
        if i % 5 == 0:
            m["score"] = float(i) / 7
        if i % 7 == 0:
            m["flag"] = bool(i % 2)
        if i % 50 == 0:
            m = None
        metas.append(m)
    return metas


@pytest.fixture(scope="module")
def coll():
    from multimodal_rag_b200 import B200Collection
    n, d = 5000, 384
    X = make_unit(n, d, 4)
    metas = _metas(n)
    c = B200Collection("w", {"hnsw:space": "cosine"})
    for s in range(0, n, 1700):                       # several batches: columns appear and grow over time
        c.add(ids=[f"id{i}" for i in range(s, min(n, s + 1700))], embeddings=X[s:s + 1700], metadatas=metas[s:s + 1700])
    dead = list(range(0, n, 9))
    c.delete(ids=[f"id{i}" for i in dead])
    alive = np.ones(n, dtype=bool)
    alive[dead] = False
    yield c, X, metas, alive
    c.close()


@pytest.mark.parametrize("where", CLAUSES)
def test_device_clause_bitmap_is_bit_exact(coll, where):
    from oracle.exact_oracle import where_match
    c, X, metas, alive = coll
    assert c._meta.compile(where) is not None, "this clause should run on the device"
    want = np.array([where_match(m, where) for m in metas]) & alive
    np.testing.assert_array_equal(c.filter_bits(where), want)
    c.device_where = False                            # the host-evaluated bitmap path gives the same rows
    try:
        np.testing.assert_array_equal(c.filter_bits(where), want)
    finally:
        c.device_where = True


@pytest.mark.parametrize("where", CLAUSES[::3])
@pytest.mark.parametrize("nq", [1, 40])
def test_filtered_queries_match_oracle(coll, where, nq):
    from oracle import exact_oracle as eo
    c, X, metas, alive = coll
    mask = np.array([eo.where_match(m, where) for m in metas]) & alive
    Q = make_unit(nq, 384, 11)
    rows, dist, cnt = c.query_rows(Q, 10, where)
    er, ed = eo.topk_exact(eo.normalize_f32(Q), eo.normalize_f32(X), 10, "cosine", allowed=mask)
    for i in range(nq):
        assert cnt[i] == len(er[i])
        np.testing.assert_array_equal(rows[i, : cnt[i]], er[i])
        np.testing.assert_allclose(dist[i, : cnt[i]], ed[i], rtol=1e-5, atol=1e-7)


def test_keys_beyond_the_device_columns_fall_back_to_a_host_bitmap():
    from multimodal_rag_b200 import B200Collection
    n = 600
    X = make_unit(n, 384, 8)
    metas = [{f"k{j}": int((i + j) % 4) for j in range(20)} for i in range(n)]
    c = B200Collection("wide", {"hnsw:space": "l2"})
    c.add(ids=[f"r{i}" for i in range(n)], embeddings=X, metadatas=metas)
    assert c._meta.compile({"k3": 1}) is not None and c._meta.compile({"k19": 1}) is None
    for where in ({"k3": 1}, {"k19": 1}, {"$and": [{"k2": {"$gte": 2}}, {"k18": {"$ne": 0}}]}):
        want = np.array([all(_leaf(m, k, v) for k, v in _flatten(where)) for m in metas])
        np.testing.assert_array_equal(c.filter_bits(where), want)
    c.close()


def _flatten(where):
    if "$and" in where:
        for w in where["$and"]:
            yield from _flatten(w)
    else:
        yield from where.items()


def _leaf(m, k, cond):
    if not isinstance(cond, dict):
        return m.get(k) == cond
    (op, v), = cond.items()
    return {"$gte": lambda a: a >= v, "$ne": lambda a: a != v}[op](m[k])


def test_malformed_programs_are_rejected(coll):
    import ctypes
    from multimodal_rag_b200 import _lib
    c, *_ = coll
    lib = _lib.load()
    out = np.zeros(200, dtype=np.uint32)

    def run(nodes, lut_words=0):
        w = _lib.B2RWhere(n_nodes=len(nodes), lut=None, lut_words=lut_words)
        for i, nd in enumerate(nodes):
            w.nodes[i] = _lib.B2RWhereNode(*nd)
        f = _lib.B2RFilter(type_mask=(1 << 64) - 1, allow_bits=None, where=ctypes.pointer(w))
        return lib.b2r_filter_eval(c.handle, ctypes.byref(f), out.ctypes.data, 0)

    assert run([(_lib.WHERE_AND, 0, 0, 0)]) == _lib.B2R_EINVAL                                  # operator without operands
    assert run([(_lib.WHERE_LEAF, 0, 0, 0), (_lib.WHERE_LEAF, 0, 0, 0)]) == _lib.B2R_EINVAL     # two results left
    assert run([(_lib.WHERE_LEAF, 99, 0, 0)]) == _lib.B2R_EINVAL                                # column out of range
    assert run([(_lib.WHERE_LEAF, 0, 0, 64)]) == _lib.B2R_EINVAL                                # table outside lut
    assert run([(_lib.WHERE_LEAF, 0, 0, 0)]) == _lib.B2R_OK and not out.any()                   # empty table: nothing passes


def test_repeated_clause_reuses_its_bitmaps(coll):
    """A session that keeps asking with the same filter pays for the clause once: the clause bitmap and the pass bitmap
    are remembered by the clause's hash until rows, tombstones or columns change."""
    from oracle import exact_oracle as eo
    c, X, metas, alive = coll
    where = {"$and": [{"type": "text"}, {"page": {"$lte": 6}}]}
    other = {"$and": [{"type": "text"}, {"page": {"$lte": 5}}]}
    Q = make_unit(12, 384, 21)
    c.set_path(2)

    def launches(w):
        before = c.stats()["launches"]
        out = c.query_rows(Q, 5, w)
        return c.stats()["launches"] - before, out

    n1, r1 = launches(where)
    n2, r2 = launches(where)
    assert n2 == n1 - 2, (n1, n2)                      # no clause kernel, no pass-bitmap kernel the second time
    np.testing.assert_array_equal(r1[0], r2[0])
    n3, r3 = launches(other)                            # a different clause is evaluated again ...
    assert n3 == n1
    mask = np.array([eo.where_match(m, other) for m in metas]) & alive
    er, _ = eo.topk_exact(eo.normalize_f32(Q), eo.normalize_f32(X), 5, "cosine", allowed=mask)
    for i in range(Q.shape[0]):
        np.testing.assert_array_equal(r3[0][i, : r3[2][i]], er[i])
    n4, r4 = launches(where)                            # ... and so is the first one after it
    assert n4 == n1
    np.testing.assert_array_equal(r4[0], r1[0])
    # a mutation invalidates: delete the best hit of query 0, the next answer must not contain it
    victim = int(r1[0][0, 0])
    c.delete(ids=[f"id{victim}"])
    alive[victim] = False
    n5, r5 = launches(where)
    assert n5 >= n1 and victim not in r5[0][0]
    c.set_path(0)


def test_sparse_keys_that_end_before_the_batch_does():
    """A key carried by the first row of a batch only: its host column is shorter than the batch, the device column
    must still be written for the whole batch (-1 = absent) and later batches must not inherit anything."""
    from multimodal_rag_b200 import B200Collection
    n = 3000
    X = make_unit(2 * n, 384, 12)
    c = B200Collection("sparse", {"hnsw:space": "cosine"})
    c.add(ids=[f"a{i}" for i in range(n)], embeddings=X[:n], metadatas=[{"first": 1} if i == 0 else {"other": i % 3} for i in range(n)])
    c.add(ids=[f"b{i}" for i in range(n)], embeddings=X[n:], metadatas=[{"late": True} if i == n - 1 else None for i in range(n)])
    want_first = np.zeros(2 * n, dtype=bool); want_first[0] = True
    want_late = np.zeros(2 * n, dtype=bool); want_late[2 * n - 1] = True
    np.testing.assert_array_equal(c.filter_bits({"first": 1}), want_first)
    np.testing.assert_array_equal(c.filter_bits({"late": True}), want_late)
    np.testing.assert_array_equal(c.filter_bits({"other": {"$in": [0, 1, 2]}}), (np.arange(2 * n) > 0) & (np.arange(2 * n) < n))
    r = c.query(query_embeddings=X[5:6].tolist(), n_results=3, where={"first": 1})
    assert r["ids"] == [["a0"]]
    c.close()


def test_where_document_on_query_get_delete():
    """Chroma's where_document ($contains / $not_contains, alone, beside a type filter, beside a compiled clause): query, get and
    delete select what the oracle's collection selects."""
    from multimodal_rag_b200 import B200Collection
    from oracle import exact_oracle as eo
    n, d = 3000, 128
    X = make_unit(n, d, 14)
    metas = _metas(n, seed=3)
    docs = [None if i % 13 == 0 else f"{('text on page', 'table row', 'image of Figure')[i % 3]} {i % 29} / {i}" for i in range(n)]
    ids = [f"id{i}" for i in range(n)]
    c = B200Collection("wd", {"hnsw:space": "cosine"})
    o = eo.ExactCollection("wd", {"hnsw:space": "cosine"})
    for col in (c, o):
        col.add(ids=ids, embeddings=X, metadatas=metas, documents=docs)
        col.delete(ids=ids[::11])
    Q = make_unit(6, d, 15)
    cases = [({"$contains": "table row"}, None), ({"$not_contains": "page"}, {"type": "image"}),
             ({"$or": [{"$contains": "Figure 3"}, {"$contains": "row 12"}]}, {"$and": [{"page": {"$gte": 2}}, {"type": {"$ne": "text"}}]}),
             ({"$contains": "never there"}, None)]
    for wd, where in cases:
        got = c.query(query_embeddings=Q, n_results=7, where=where, where_document=wd)
        want = o.query(Q, n_results=7, where=where, where_document=wd)
        assert got["ids"] == want["ids"] and got["documents"] == want["documents"]
        for a, b in zip(got["distances"], want["distances"]):
            np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-7)
        assert c.get(where=where, where_document=wd)["ids"] == o.get(where=where, where_document=wd)["ids"]
    gone = c.delete(where_document={"$contains": "image of"})
    before = set(o._row_of)
    o.delete(where_document={"$contains": "image of"})
    assert sorted(gone) == sorted(before - set(o._row_of)) and len(gone) > 500
    assert c.count() == o.count()
    assert c.query(query_embeddings=Q, n_results=5)["ids"] == o.query(Q, n_results=5)["ids"]
    with pytest.raises(ValueError):
        c.query(query_embeddings=Q, n_results=5, where_document={"$contains": ""})
    c.close()
