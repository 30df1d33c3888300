"""Regenerates tests/golden/chroma_wal.{npz,json} from the reference's committed Chroma DB.

Run in the builder container only (needs /root/reference, which does not exist on the
GPU box):  python tests/golden/make_golden.py

Source: /root/reference/chroma_db/chroma.sqlite3, table ``embeddings_queue`` (Chroma's
write-ahead log): 70 ADD rows (operation=0, encoding FLOAT32, 1536-byte little-endian
vectors) followed by 70 DELETE rows.  The known answers are computed here with a
deliberately naive fp64 loop that shares no code with oracle/ so they can pin it.
"""
import json
import os
import sqlite3

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DB = "/root/reference/chroma_db/chroma.sqlite3"


def main():
    con = sqlite3.connect(f"file:{DB}?mode=ro", uri=True)
    space = con.execute("select str_value from collection_metadata where key='hnsw:space'").fetchone()[0]
    name, dim = con.execute("select name, dimension from collections").fetchone()
    wal = con.execute("select seq_id, operation, id, vector, encoding, metadata from embeddings_queue "
                      "order by seq_id").fetchall()
    adds = [r for r in wal if r[1] == 0]
    assert all(r[4] == "FLOAT32" and len(r[3]) == 4 * dim for r in adds)
    X = np.stack([np.frombuffer(r[3], dtype="<f4") for r in adds]).astype(np.float32)
    ids = [r[2] for r in adds]
    metas = [json.loads(r[5]) for r in adds]
    ops = [[int(r[0]), int(r[1]), r[2]] for r in wal]

    # naive fp64 known answers: query = WAL row 0, corpus = rows 1..69
    def naive(qi, rows, k, pred=lambda m: True):
        out = []
        for r in rows:
            if not pred(metas[r]):
                continue
            dot = 0.0
            l2 = 0.0
            for a, b in zip(X[qi].tolist(), X[r].tolist()):
                dot += a * b
                l2 += (a - b) * (a - b)
            out.append((1.0 - dot, l2, r))
        out.sort(key=lambda t: (t[0], t[2]))
        return [{"id": ids[r], "row": r, "cosine": c, "l2": l} for c, l, r in out[:k]]

    known = {
        "query_row": 0,
        "corpus_rows": [1, 69],
        "top5": naive(0, range(1, 70), 5),
        "top5_image": naive(0, range(1, 70), 5, lambda m: m["type"] == "image"),
        "top10_self_excluded_row7": naive(7, [r for r in range(70) if r != 7], 10),
    }
    np.savez_compressed(os.path.join(HERE, "chroma_wal.npz"), vectors=X)
    with open(os.path.join(HERE, "chroma_wal.json"), "w") as f:
        json.dump({"collection": name, "dimension": dim, "space": space, "ids": ids,
                   "metadatas": metas, "wal_ops": ops, "known": known}, f, indent=1)
    print("wrote", X.shape, space, known["top5"][0])


if __name__ == "__main__":
    main()
