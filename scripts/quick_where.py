"""Filtered queries: compiled clause on the device vs host-evaluated bitmap (development aid, not the bench).

    python scripts/quick_where.py [N D]
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multimodal_rag_b200 import B200Collection

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rng = np.random.default_rng(0x7E57)
types = rng.choice(["text", "table", "image"], size=n, p=[0.6, 0.1, 0.3])
c = B200Collection("w", {"hnsw:space": "cosine"}, capacity=n, dimension=d)
g = torch.Generator(device="cuda").manual_seed(1)
step = 1 << 17
t0 = time.perf_counter()
for s in range(0, n, step):
    m = min(step, n - s)
    x = torch.randn(m, d, generator=g, device="cuda")
    c.add(ids=[f"doc_{i // 64:06x}_{i}" for i in range(s, s + m)], embeddings=x,
          metadatas=[{"type": str(types[i]), "doc_id": f"doc_{i // 64:06x}", "page": int(i % 9)} for i in range(s, s + m)])
print(f"ingest {n} rows with metadata: {time.perf_counter() - t0:.1f} s", flush=True)
clauses = [None, {"type": "image"}, {"$and": [{"type": "image"}, {"page": {"$gte": 3}}]},
           {"doc_id": {"$in": [f"doc_{j:06x}" for j in range(100, 140)]}}]
for nq in (1, 256):
    q = torch.nn.functional.normalize(torch.randn(nq, d, generator=g, device="cuda"), dim=1)
    for where in clauses:
        for dev in (True, False):
            if where is None and not dev:
                continue
            c.device_where = dev
            for _ in range(3):
                c.query_rows(q, 10, where)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            it = 20
            for _ in range(it):
                rows, dist, cnt = c.query_rows(q, 10, where)
            dt = (time.perf_counter() - t0) / it
            kind = "none" if where is None else ("type mask" if c._meta.type_only_mask(where) is not None else ("device clause" if dev else "host bitmap"))
            print(f"nq={nq:4d} where={str(where)[:60]:60s} {kind:14s} {dt * 1e6:9.1f} us/call  ({nq / dt:10.0f} q/s)  hits/query={cnt.mean():.1f}", flush=True)
c.close()
