"""CPU: the native id table (b2r_idtab_*, csrc/idtable.cu) against a Python dict model of what Chroma's id index does
for the reference's calls (add / upsert / get(ids) / delete(ids): app/utils/embedder.py:518, 632, 640, 888)."""
import ctypes

import numpy as np
import pytest

from multimodal_rag_b200 import _lib, build as b2r_build
from multimodal_rag_b200.idtable import IdTable, IdsByRow, RowOfId, encode_ids


@pytest.fixture(scope="module", autouse=True)
def _built():
    b2r_build.build()


def test_encode_forms():
    e = encode_ids(["a", "bc", "déf"])
    assert e.gap == 1 and e.n == 3 and e.off.tolist() == [0, 2, 5, 10] and not e.has_empty()
    assert encode_ids(["a", "", "b"]).has_empty()
    assert encode_ids([""]).has_empty()
    z = encode_ids(["a\0b", "c"])                     # an id with a NUL inside: packed form
    assert z.gap == 0 and z.off.tolist() == [0, 3, 4] and z.buf == b"a\0bc"
    assert encode_ids([]).n == 0
    with pytest.raises(TypeError):
        encode_ids(["a", 3])


def test_lookup_append_erase_follow_a_dict_model():
    rng = np.random.default_rng(7)
    t = IdTable()
    model, ids_by_row = {}, []
    universe = [f"doc{j}_text_{i}" for j in range(40) for i in range(60)] + ["ü" * 30, "x\0y", "x", "y" * 300]
    for step in range(60):
        n = int(rng.integers(1, 200))
        batch = [universe[i] for i in rng.choice(len(universe), size=n, replace=False)]
        enc = encode_ids(batch)
        found, dup = t.lookup(enc, want_dup=True)
        assert dup == -1 and found.tolist() == [model.get(i, -1) for i in batch]
        if step % 3 == 2:                              # delete what is there
            rows = found[found >= 0]
            t.erase_rows(rows)
            for i in batch:
                model.pop(i, None)
        else:                                          # upsert
            prev = t.append(enc, len(ids_by_row))
            assert prev.tolist() == found.tolist()
            for j, i in enumerate(batch):
                model[i] = len(ids_by_row) + j
            ids_by_row.extend(batch)
        assert t.live == len(model) and t.rows == len(ids_by_row)
    # full read-back: every id, and every row's bytes (erased rows keep theirs)
    assert t.lookup(encode_ids(universe)).tolist() == [model.get(i, -1) for i in universe]
    assert t.ids_of(np.arange(len(ids_by_row))) == ids_by_row
    assert IdsByRow(t)[5] == ids_by_row[5] and IdsByRow(t)[-1] == ids_by_row[-1] and len(IdsByRow(t)) == len(ids_by_row)
    view = RowOfId(t)
    some = next(iter(model))
    assert some in view and view[some] == model[some] and "never seen" not in view and view.get(3) is None
    t.clear()
    assert t.live == 0 and t.rows == 0 and t.lookup(encode_ids(universe[:5])).tolist() == [-1] * 5


def test_in_batch_repeats_are_reported_and_erase_keeps_a_newer_mapping():
    t = IdTable()
    _, dup = t.lookup(encode_ids(["a", "b", "c", "b", "a"]), want_dup=True)
    assert dup == 3
    _, dup = t.lookup(encode_ids(["ab", "a", "b", "a\0b", "a"]), want_dup=True)
    assert dup == 4
    t.append(encode_ids(["a", "b"]), 0)
    t.append(encode_ids(["a"]), 2)                     # upsert: "a" now lives at row 2
    t.erase_rows([0])                                  # the old version's tombstone must not unmap it
    assert t.lookup(encode_ids(["a", "b"])).tolist() == [2, 1] and t.live == 2
    t.erase_rows([2, 1])
    assert t.lookup(encode_ids(["a", "b"])).tolist() == [-1, -1] and t.live == 0
    t.append(encode_ids(["b"]), 3)                     # delete, then add again
    assert t.lookup(encode_ids(["b"])).tolist() == [3] and t.ids_of([0, 1, 2, 3]) == ["a", "b", "a", "b"]


def test_growth_and_runs_of_colliding_slots():
    """1.5M ids through doubling rehashes; then deletions out of the middle of probe runs (backward shift, no tombstones)
    leave every survivor findable."""
    t = IdTable(reserve=1000)
    n = 1_500_000
    ids = [f"r{i}" for i in range(n)]
    for s in range(0, n, 1 << 18):
        t.append(encode_ids(ids[s: s + (1 << 18)]), s)
    assert t.live == n
    rng = np.random.default_rng(3)
    gone = rng.choice(n, size=n // 2, replace=False)
    t.erase_rows(gone)
    want = np.arange(n)
    want[gone] = -1
    assert np.array_equal(t.lookup(encode_ids(ids)), want) and t.live == n - n // 2


def test_bad_arguments():
    lib = _lib.load()
    t = IdTable()
    with pytest.raises(ValueError):
        t.append(encode_ids(["a"]), 5)                 # first_row must follow the rows appended so far
    with pytest.raises(ValueError):
        t.ids_of([0])
    with pytest.raises(ValueError):
        t.erase_rows([-1])
    off = np.asarray([0, 3, 2], dtype=np.int64)
    rows = np.zeros(2, dtype=np.int64)
    assert lib.b2r_idtab_lookup(t._h, b"abcd", off.ctypes.data, 2, 0, rows.ctypes.data, None) == _lib.B2R_EINVAL
    assert lib.b2r_idtab_live(None) == -1 and lib.b2r_idtab_create(-1, ctypes.byref(ctypes.c_void_p())) == _lib.B2R_EINVAL
    long_ids = ["k" * 5000, "m" * 7000]                # ids_of grows its buffer
    t.append(encode_ids(long_ids), 0)
    assert t.ids_of([1, 0]) == long_ids[::-1]
