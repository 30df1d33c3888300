"""A deterministic stand-in for the reference's sentence encoder (all-MiniLM-L6-v2 via sentence-transformers,
/root/reference/app/utils/embedder.py:385-405): text -> unit-norm 384-d fp32 vector seeded by the text's MD5.
Shared by the script that runs the reference's host logic to produce golden results and by the tests that replay the
same scenario through this repo's façade."""
import hashlib

import numpy as np

DIM = 384


def fake_embed(texts):
    out = np.empty((len(texts), DIM), dtype=np.float32)
    for i, t in enumerate(texts):
        seed = int.from_bytes(hashlib.md5(t.encode("utf-8")).digest()[:8], "little")
        v = np.random.default_rng(seed).standard_normal(DIM).astype(np.float32)
        out[i] = v / np.linalg.norm(v)
    return out


# the scenario both sides run
DOC_A, DOC_B = "doc_aaaaaaaaaaaa", "doc_bbbbbbbbbbbb"
SUMMARIES_A = ([{"id": f"text_{i}", "summary": f"chunk {i} about pointers and arrays in C, part {i % 3}", "raw": f"raw text {i}", "type": "text"} for i in range(9)]
               + [{"id": "table_1", "summary": "table of operator precedence in C", "raw": "<table>..</table>", "type": "table"}]
               + [{"id": f"Session01_page_{i}_0{i}ab", "summary": f"Image: slide {i} of the C course", "raw": f"b64img{i}", "type": "image", "path": f"figures/p{i}.png"} for i in range(4)])
SUMMARIES_B = [{"id": f"text_{i}", "summary": f"paragraph {i} on loops, for and while statements", "raw": f"raw b {i}", "type": "text"} for i in range(6)]
QUERIES = ["what is a pointer", "loops in C", "", "operator precedence table", "   "]
