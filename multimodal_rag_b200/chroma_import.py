"""Migration from an existing Chroma persist directory.

The reference persists its collection with Chroma (``CHROMA_PERSIST_DIR=./chroma_db``, ``/root/reference/config.py:58``;
the committed ``chroma_db/chroma.sqlite3`` is such a directory).  Chroma's sqlite file carries a write-ahead log,
table ``embeddings_queue`` (seq_id, operation {0 ADD, 1 UPDATE, 2 UPSERT, 3 DELETE}, id, vector BLOB little-endian
FLOAT32, metadata JSON whose ``chroma:document`` key is the document), plus ``collections`` (name, dimension) and
``collection_metadata`` (``hnsw:space``).  Replaying the log reproduces the collection exactly -- same ids, original
vectors (no re-embedding), metadatas, documents, insertion order -- inside a ``B200Collection``.

Host-side Python over the standard ``sqlite3`` module; the vectors go to the GPU through the normal ``add`` /
``upsert`` / ``delete`` calls, so everything device-side is the engine's own ingest path.
"""
from __future__ import annotations

import json
import sqlite3
from typing import Iterator, Optional

import numpy as np

OP_ADD, OP_UPDATE, OP_UPSERT, OP_DELETE = 0, 1, 2, 3


def read_chroma_wal(sqlite_path: str, collection_name: Optional[str] = None, upto_seq: Optional[int] = None) -> dict:
    """-> {name, dimension, space, ops: [(seq_id, op, id, vector|None, metadata|None, document|None)]}"""
    con = sqlite3.connect(f"file:{sqlite_path}?mode=ro", uri=True)
    try:
        rows = con.execute("select id, name, topic, dimension from collections").fetchall()
        if not rows:
            raise ValueError(f"{sqlite_path}: no collections")
        if collection_name is not None:
            rows = [r for r in rows if r[1] == collection_name]
            if not rows:
                raise ValueError(f"Collection {collection_name} does not exist.")
        cid, name, topic, dim = rows[0]
        sp = con.execute("select str_value from collection_metadata where collection_id=? and key='hnsw:space'", (cid,)).fetchone()
        space = sp[0] if sp and sp[0] else "l2"                     # Chroma's default
        q = "select seq_id, operation, id, vector, encoding, metadata from embeddings_queue where topic=?"
        args = [topic]
        if upto_seq is not None:
            q += " and seq_id<=?"
            args.append(upto_seq)
        ops = []
        for seq, op, id_, blob, enc, meta in con.execute(q + " order by seq_id", args):
            vec = md = doc = None
            if blob is not None:
                if enc == "FLOAT32":
                    vec = np.frombuffer(blob, dtype="<f4").astype(np.float32)
                elif enc == "INT32":
                    vec = np.frombuffer(blob, dtype="<i4").astype(np.float32)
                else:
                    raise ValueError(f"seq {seq}: unknown vector encoding {enc!r}")
                if dim is not None and vec.shape[0] != dim:
                    raise ValueError(f"seq {seq}: vector of length {vec.shape[0]} in a {dim}-d collection")
            if meta:
                md = json.loads(meta)
                doc = md.pop("chroma:document", None)
            ops.append((int(seq), int(op), id_, vec, md, doc))
        return {"name": name, "dimension": dim, "space": space, "ops": ops}
    finally:
        con.close()


def _runs(ops) -> Iterator[list]:
    """maximal runs of consecutive ops of one kind that do not repeat an id (one device call each)"""
    run, seen = [], set()
    for o in ops:
        if run and (o[1] != run[0][1] or o[2] in seen):
            yield run
            run, seen = [], set()
        run.append(o)
        seen.add(o[2])
    if run:
        yield run


def replay(ops, collection) -> dict:
    """Apply WAL ops to a collection with Chroma's semantics: ADD of a live id is skipped, UPDATE touches only live
    ids (fields absent from the record keep their value), UPSERT overwrites or inserts, DELETE removes."""
    n = {"add": 0, "update": 0, "upsert": 0, "delete": 0}
    for run in _runs(ops):
        op = run[0][1]
        ids = [o[2] for o in run]
        if op == OP_DELETE:
            collection.delete(ids=ids)
            n["delete"] += len(ids)
            continue
        if op == OP_UPDATE:
            cur = collection.get(ids=ids, include=["embeddings", "metadatas", "documents"])
            have = {i: j for j, i in enumerate(cur["ids"])}
            run = [o for o in run if o[2] in have]
            if not run:
                continue
            ids = [o[2] for o in run]
            vecs = [o[3] if o[3] is not None else np.asarray(cur["embeddings"][have[o[2]]], dtype=np.float32) for o in run]
            metas = [{**(cur["metadatas"][have[o[2]]] or {}), **(o[4] or {})} or None for o in run]
            docs = [o[5] if o[5] is not None else cur["documents"][have[o[2]]] for o in run]
            collection.upsert(ids=ids, embeddings=np.stack(vecs), metadatas=metas, documents=docs)
            n["update"] += len(ids)
            continue
        if any(o[3] is None for o in run):
            raise ValueError(f"seq {run[0][0]}: ADD/UPSERT without a vector (this engine has no embedding function)")
        vecs = np.stack([o[3] for o in run])
        metas = [o[4] or None for o in run]
        docs = [o[5] for o in run]
        if op == OP_ADD:
            collection.add(ids=ids, embeddings=vecs, metadatas=metas, documents=docs)
            n["add"] += len(ids)
        else:
            collection.upsert(ids=ids, embeddings=vecs, metadatas=metas, documents=docs)
            n["upsert"] += len(ids)
    return n


def import_chroma_wal(sqlite_path: str, client, collection_name: Optional[str] = None, upto_seq: Optional[int] = None,
                      **collection_kw):
    """Create (or reuse) the collection in `client` with the persisted name / space and replay the log into it."""
    wal = read_chroma_wal(sqlite_path, collection_name, upto_seq)
    c = client.create_collection(wal["name"], {"hnsw:space": wal["space"]}, get_or_create=True, **collection_kw)
    replay(wal["ops"], c)
    return c
