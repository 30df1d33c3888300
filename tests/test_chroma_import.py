"""Chroma WAL importer (multimodal_rag_b200/chroma_import.py): parsing on CPU against a sqlite file rebuilt from the
golden fixture with Chroma 0.4.22's schema; replay semantics with a recording stand-in; the real replay on the GPU."""
import json
import sqlite3

import numpy as np
import pytest

SCHEMA = """
CREATE TABLE collections (id TEXT PRIMARY KEY, name TEXT NOT NULL, topic TEXT NOT NULL, dimension INTEGER, database_id TEXT NOT NULL);
CREATE TABLE collection_metadata (collection_id TEXT, key TEXT NOT NULL, str_value TEXT, int_value INTEGER, float_value REAL);
CREATE TABLE embeddings_queue (seq_id INTEGER PRIMARY KEY, created_at TIMESTAMP NOT NULL DEFAULT CURRENT_TIMESTAMP,
    operation INTEGER NOT NULL, topic TEXT NOT NULL, id TEXT NOT NULL, vector BLOB, encoding TEXT, metadata TEXT);
"""


def build_sqlite(path, golden):
    """the reference's committed chroma.sqlite3, rebuilt from tests/golden (140 ops: 70 ADD interleaved with 70 DELETE)"""
    con = sqlite3.connect(path)
    con.executescript(SCHEMA)
    topic = "persistent://default/default/9236d1cb"
    con.execute("insert into collections values ('9236d1cb', ?, ?, 384, '0')", (golden["collection"], topic))
    con.execute("insert into collection_metadata values ('9236d1cb', 'hnsw:space', ?, NULL, NULL)", (golden["space"],))
    row_of = {i: r for r, i in enumerate(golden["ids"])}
    for seq, op, id_ in golden["wal_ops"]:
        if op == 0:
            r = row_of[id_]
            con.execute("insert into embeddings_queue (seq_id, operation, topic, id, vector, encoding, metadata) values (?,?,?,?,?,?,?)",
                        (seq, 0, topic, id_, golden["vectors"][r].astype("<f4").tobytes(), "FLOAT32", json.dumps(golden["metadatas"][r])))
        else:
            con.execute("insert into embeddings_queue (seq_id, operation, topic, id) values (?,?,?,?)", (seq, op, topic, id_))
    con.commit()
    con.close()


def test_read_wal_roundtrip(tmp_path, golden):
    from multimodal_rag_b200.chroma_import import read_chroma_wal
    p = str(tmp_path / "chroma.sqlite3")
    build_sqlite(p, golden)
    wal = read_chroma_wal(p)
    assert (wal["name"], wal["dimension"], wal["space"]) == (golden["collection"], 384, golden["space"])
    assert len(wal["ops"]) == 140 and sum(o[1] == 0 for o in wal["ops"]) == 70 and sum(o[1] == 3 for o in wal["ops"]) == 70
    adds = [o for o in wal["ops"] if o[1] == 0]
    np.testing.assert_array_equal(np.stack([o[3] for o in adds]), golden["vectors"])
    assert [o[2] for o in adds] == golden["ids"]
    assert adds[0][5] == golden["metadatas"][0]["chroma:document"] and "chroma:document" not in adds[0][4]
    assert adds[0][4]["type"] == "text"
    assert len(read_chroma_wal(p, upto_seq=71)["ops"]) == 71
    with pytest.raises(ValueError):
        read_chroma_wal(p, collection_name="nope")


class Recorder:
    """stands in for a collection: records the calls replay() makes"""
    def __init__(self):
        self.calls, self.live = [], {}

    def add(self, ids, embeddings, metadatas, documents):
        self.calls.append(("add", list(ids)))
        for i, e, m, d in zip(ids, embeddings, metadatas, documents):
            self.live.setdefault(i, (e, m, d))

    def upsert(self, ids, embeddings, metadatas, documents):
        self.calls.append(("upsert", list(ids)))
        for i, e, m, d in zip(ids, embeddings, metadatas, documents):
            self.live[i] = (e, m, d)

    def delete(self, ids):
        self.calls.append(("delete", list(ids)))
        for i in ids:
            self.live.pop(i, None)

    def get(self, ids, include):
        ids = [i for i in ids if i in self.live]
        return {"ids": ids, "embeddings": [self.live[i][0].tolist() for i in ids],
                "metadatas": [self.live[i][1] for i in ids], "documents": [self.live[i][2] for i in ids]}


def test_replay_semantics_batches_and_update():
    from multimodal_rag_b200.chroma_import import replay
    v = lambda x: np.full(4, x, dtype=np.float32)
    ops = [(1, 0, "a", v(1), {"type": "text"}, "A"), (2, 0, "b", v(2), {"type": "image"}, "B"),
           (3, 0, "a", v(9), {"type": "text"}, "dup"),                      # ADD of a live id: its own call, skipped downstream
           (4, 1, "b", None, {"page": 3}, None),                            # UPDATE metadata only
           (5, 1, "zz", v(5), None, None),                                  # UPDATE of an unknown id: ignored
           (6, 2, "c", v(3), None, "C"), (7, 3, "a", None, None, None)]
    r = Recorder()
    n = replay(ops, r)
    assert [c[0] for c in r.calls] == ["add", "add", "upsert", "upsert", "delete"]
    assert r.calls[0][1] == ["a", "b"] and r.calls[1][1] == ["a"]
    assert set(r.live) == {"b", "c"}
    assert r.live["b"][1] == {"type": "image", "page": 3} and r.live["b"][2] == "B" and r.live["b"][0][0] == 2.0
    assert n == {"add": 3, "update": 1, "upsert": 1, "delete": 1}


@pytest.mark.gpu
def test_import_reference_wal_on_gpu(tmp_path, golden):
    """state after seq 71 = rows 1..69 live -> the reference's known answers; the full log ends empty"""
    from multimodal_rag_b200 import B200Client
    from multimodal_rag_b200.chroma_import import import_chroma_wal
    p = str(tmp_path / "chroma.sqlite3")
    build_sqlite(p, golden)
    c = import_chroma_wal(p, B200Client(), upto_seq=71)
    assert c.name == golden["collection"] and c.space == golden["space"] and c.count() == 69
    r = c.query(query_embeddings=[golden["vectors"][0].tolist()], n_results=5)
    assert r["ids"][0] == [a["id"] for a in golden["known"]["top5"]]
    np.testing.assert_allclose(r["distances"][0], [a["cosine"] for a in golden["known"]["top5"]], rtol=1e-5)
    assert r["documents"][0][0] == golden["metadatas"][golden["known"]["top5"][0]["row"]]["chroma:document"]
    assert "chroma:document" not in r["metadatas"][0][0]
    c2 = import_chroma_wal(p, B200Client())
    assert c2.count() == 0
