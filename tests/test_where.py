"""CPU: the product's columnar where-evaluator (multimodal_rag_b200/where.py) against the oracle's
per-row restatement of Chroma's grammar."""
import numpy as np
import pytest

from multimodal_rag_b200.where import MetaTable, pack_bits
from oracle.exact_oracle import where_match


def _table(n=500, seed=0):
    rng = np.random.default_rng(seed)
    t, metas = MetaTable(), []
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
        t.append(m)
    return t, metas


CLAUSES = [
    {"type": "image"}, {"type": {"$eq": "text"}}, {"type": {"$ne": "text"}}, {"type": {"$in": ["image", "table"]}},
    {"type": {"$nin": ["image"]}}, {"doc_id": "doc_0007"}, {"page": {"$gte": 5}}, {"page": {"$lt": 3}},
    {"page": 4}, {"page": 4.0}, {"score": {"$gt": 20.5}}, {"flag": True}, {"flag": {"$ne": True}},
    {"page": {"$ne": "4"}}, {"missing": "x"}, {"missing": {"$ne": "x"}}, {"type": "video"},
    {"$and": [{"type": "text"}, {"page": {"$lte": 2}}]},
    {"$or": [{"type": "image"}, {"$and": [{"doc_id": "doc_0003"}, {"page": {"$in": [1, 2, 3]}}]}]},
    {"page": {"$in": [1, "1", True]}},
]


@pytest.mark.parametrize("where", CLAUSES)
def test_mask_matches_oracle(where):
    t, metas = _table()
    got = t.mask(where)
    want = np.array([where_match(m, where) for m in metas])
    np.testing.assert_array_equal(got, want)


def test_type_fast_path_and_bits():
    t, metas = _table()
    tm = t.type_only_mask({"type": "image"})
    codes = np.array([t.type_code_of(m) for m in metas])
    np.testing.assert_array_equal(((tm >> codes) & 1).astype(bool), t.mask({"type": "image"}))
    tm = t.type_only_mask({"type": {"$in": ["image", "nope"]}})
    np.testing.assert_array_equal(((tm >> codes) & 1).astype(bool), t.mask({"type": "image"}))
    assert t.type_only_mask({"type": "nope"}) == 0
    assert t.type_only_mask({"doc_id": "doc_0001"}) is None and t.type_only_mask({"type": {"$ne": "text"}}) is None
    m = t.mask({"page": {"$gte": 5}})
    bits = pack_bits(m)
    assert bits.dtype == np.uint32 and bits.shape[0] == (m.shape[0] + 31) // 32
    back = np.array([(bits[r >> 5] >> (r & 31)) & 1 for r in range(m.shape[0])], dtype=bool)
    np.testing.assert_array_equal(back, m)


def test_type_code_overflow_falls_back_to_bitmap():
    t = MetaTable()
    for i in range(100):
        t.append({"type": f"kind{i}"})
    assert t.type_overflow and t.type_only_mask({"type": "kind3"}) is None
    assert t.mask({"type": "kind99"}).sum() == 1


@pytest.mark.parametrize("bad", [{"a": 1, "b": 2}, {"$xor": []}, {"a": {"$like": "x"}}, {"a": {"$gt": "x"}},
                                 {"a": {"$in": []}}, {"$and": []}])
def test_bad_clauses_raise(bad):
    t, _ = _table(20)
    with pytest.raises(ValueError):
        t.mask(bad)


def _run_program(t, nodes, lut):
    """Host model of where_bits_kernel: the compiled postfix program over the dictionary codes."""
    cols = list(t.cols.values())
    stack = []
    for op, col, off, nv in nodes:
        if op == 0:
            bit = np.zeros(t.nrows, dtype=bool)
            if nv:
                codes = np.full(t.nrows, -1, dtype=np.int64)
                codes[: cols[col].n] = cols[col].codes[: cols[col].n]
                ok = (codes >= 0) & (codes < nv)
                c = np.where(ok, codes, 0)
                bit = ok & (((lut[off + (c >> 5)] >> (c & 31).astype(np.uint32)) & 1) == 1)
            stack.append(bit)
        else:
            b, a = stack.pop(), stack.pop()
            stack.append(a & b if op == 1 else a | b)
    assert len(stack) == 1
    return stack[0]


@pytest.mark.parametrize("where", CLAUSES)
def test_compiled_program_matches_host_mask(where):
    t, metas = _table()
    prog = t.compile(where)
    assert prog is not None
    nodes, lut = prog
    assert lut.dtype == np.uint32 and len(nodes) <= 32
    np.testing.assert_array_equal(_run_program(t, nodes, lut), t.mask(where))


def test_compile_gives_up_on_host_only_columns_and_long_clauses():
    t = MetaTable()
    for i in range(40):
        t.append({f"k{j}": i % (j + 2) for j in range(20)})
    assert t.compile({"k15": 1}) is not None and t.compile({"k16": 1}) is None
    assert t.compile({"$or": [{"k1": i} for i in range(17)]}) is None          # 17 leaves + 16 ORs = 33 nodes
    assert t.compile({"$or": [{"k1": i} for i in range(16)]}) is not None
    with pytest.raises(ValueError):
        t.compile({"k1": {"$gt": "a"}})


DOC_CLAUSES = [
    {"$contains": "table"}, {"$not_contains": "table"}, {"$contains": "Figure 3"}, {"$contains": "é"},
    {"$and": [{"$contains": "page"}, {"$not_contains": "7"}]},
    {"$or": [{"$contains": "image of"}, {"$and": [{"$contains": "row"}, {"$contains": "12"}]}]},
    {"$contains": "never there"},
]


def _docs(n=400):
    docs = []
    for i in range(n):
        kind = ("text on page", "table row", "image of Figure", "résumé é")[i % 4]
        docs.append(None if i % 13 == 0 else f"{kind} {i % 29} / {i}")
    return docs


@pytest.mark.parametrize("wd", DOC_CLAUSES)
def test_doc_mask_matches_oracle(wd):
    """where_document ($contains / $not_contains / $and / $or): the product's pass over the document list against the oracle's
    per-row restatement; a row without a document matches neither operator."""
    from multimodal_rag_b200.where import doc_mask
    from oracle.exact_oracle import where_document_match
    docs = _docs()
    want = np.asarray([where_document_match(d, wd) for d in docs])
    np.testing.assert_array_equal(doc_mask(docs, wd), want)


def test_doc_mask_rejects_malformed_clauses():
    from multimodal_rag_b200.where import doc_mask
    for bad in ({"$contains": ""}, {"$contains": 3}, {"$like": "x"}, {"$and": []}, {"$contains": "a", "$not_contains": "b"}, "table"):
        with pytest.raises(ValueError):
            doc_mask(["a"], bad)


def test_batch_ingestion_builds_the_same_tables_as_row_by_row():
    """MetaTable.append_batch / validate_batch (what B200Collection.add uses) against append / validate row by row: same
    dictionaries, codes, column order (= device column numbers) and masks -- with absent keys, rows without metadata, values of
    mixed kinds in one column (4 / 4.0 / "4" / True), numpy scalars, several batches."""
    rng = np.random.default_rng(8)
    metas = []
    for i in range(3000):
        m = {"doc_id": f"doc_{i % 41:03d}", "type": str(rng.choice(["text", "table", "image"]))}
        if i % 3:
            m["page"] = [4, 4.0, "4", True, 7, 2.5][i % 6]
        if i % 5 == 0:
            m["score"] = np.float64(i) / 3 if i % 10 else float(i)
        if i > 1500 and i % 7 == 0:
            m["late_key"] = f"v{i % 4}"
        metas.append(None if i % 29 == 0 else ({} if i % 31 == 0 else m))
    a, b = MetaTable(), MetaTable()
    for md in metas:
        MetaTable.validate(md)
        a.append(md)
    for lo in (0, 700, 701, 2000):
        hi = {0: 700, 700: 701, 701: 2000, 2000: 3000}[lo]
        MetaTable.validate_batch(metas[lo:hi])
        b.append_batch(metas[lo:hi])
    assert a.nrows == b.nrows and a.meta == b.meta and list(a.cols) == list(b.cols)
    for k in a.cols:
        ca, cb = a.cols[k], b.cols[k]
        assert ca.values == cb.values and ca.kinds == cb.kinds and ca.n == cb.n, k
        np.testing.assert_array_equal(ca.codes[: ca.n], cb.codes[: cb.n])
    for where in ({"page": 4}, {"page": {"$ne": "4"}}, {"page": True}, {"score": {"$gt": 100.0}}, {"late_key": {"$in": ["v1", "v3"]}},
                  {"$or": [{"type": "image"}, {"doc_id": "doc_007"}]}):
        np.testing.assert_array_equal(a.mask(where), b.mask(where))
        want = np.asarray([where_match(m, where) for m in metas])
        np.testing.assert_array_equal(b.mask(where), want)
    for bad in ([{"k": [1]}], [{3: "x"}], ["not a dict"], [{"k": None}]):
        with pytest.raises(ValueError):
            MetaTable.validate_batch(bad)
