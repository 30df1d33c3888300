"""Regenerates tests/golden/reference_host_logic.json by RUNNING THE REFERENCE'S OWN host code.

Run in the builder container only (needs /root/reference):  python tests/golden/make_reference_golden.py

The reference's `EmbeddingManager` (app/utils/embedder.py) and `MultiVectorRetriever` (app/utils/retriever.py) are
loaded by file path, unmodified.  What they import but this image lacks (chromadb, sentence_transformers, redis) is
replaced by empty stub modules, and the three objects they would get from those packages are injected:
  * the Chroma collection  -> the oracle's ExactCollection (oracle/exact_oracle.py; exact answers in Chroma's result shape)
  * the SentenceTransformer -> tests/fake_encoder.py (deterministic text -> unit vector)
  * the Redis client       -> an in-memory dict with get / pipeline that records the keys each pipeline asked for
Everything else -- id and metadata construction, the MD5-keyed LRU embedding cache, result flattening, batch_query's
error entries, get_similar_documents' self-exclusion, delete_document, the stats dictionaries, the id -> Redis key rule,
the cache-then-one-pipeline fetch plan and the by-type bucketing -- is the reference's code executing.
"""
import asyncio
import gzip
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from fake_encoder import DOC_A, DOC_B, QUERIES, SUMMARIES_A, SUMMARIES_B, fake_embed  # noqa: E402
from oracle.exact_oracle import ExactCollection  # noqa: E402


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    for name in ("redis", "redis.asyncio", "chromadb", "chromadb.config", "sentence_transformers"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["redis"].asyncio = sys.modules["redis.asyncio"]
    sys.modules["chromadb.config"].Settings = object
    sys.modules["sentence_transformers"].SentenceTransformer = object
    E = _load("ref_embedder", "/root/reference/app/utils/embedder.py")
    R = _load("ref_retriever", "/root/reference/app/utils/retriever.py")
    return E, R


class FakeModel:
    def encode(self, texts, **kw):
        assert kw.get("normalize_embeddings") is True
        return fake_embed(list(texts))

    def get_sentence_embedding_dimension(self):
        return 384


class FakeClient:
    def __init__(self):
        self.collections = {}

    def create_collection(self, name, metadata=None):
        self.collections[name] = ExactCollection(name, metadata)
        return self.collections[name]

    def get_collection(self, name):
        return self.collections[name]

    def delete_collection(self, name):
        del self.collections[name]


class FakePipe:
    def __init__(self, store, log):
        self.store, self.log, self.keys = store, log, []

    def get(self, key):
        self.keys.append(key)

    async def execute(self):
        self.log.append(list(self.keys))
        return [self.store.get(k) for k in self.keys]

    async def __aenter__(self):
        return self

    async def __aexit__(self, *a):
        return False


class FakeRedis:
    def __init__(self):
        self.store, self.pipelines = {}, []

    def pipeline(self, transaction=False):
        return FakePipe(self.store, self.pipelines)


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    return o


async def run():
    E, R = load_reference()
    out = {}
    # ---------------- EmbeddingManager ----------------
    m = E.EmbeddingManager(batch_size=4, enable_cache=True, cache_size=8, device="cpu", enable_progress_logging=False)
    m.client = FakeClient()
    m.collection = m.client.create_collection("multimodal_rag", {"description": "Multimodal RAG document embeddings"})
    m.text_model = FakeModel()
    m.is_initialized = True
    out["counts_a"] = await m.embed_and_store(SUMMARIES_A, DOC_A)
    out["counts_b"] = await m.embed_and_store(SUMMARIES_B, DOC_B)
    out["stored"] = m.collection.get(include=["metadatas", "documents"])
    out["query_1"] = await m.query(QUERIES[0], n_results=5)
    out["query_1_again"] = await m.query(QUERIES[0], n_results=5)            # served from the embedding cache
    out["query_image"] = await m.query(QUERIES[0], n_results=3, filter_dict={"type": "image"})
    out["query_doc_b"] = await m.query(QUERIES[1], n_results=20, filter_dict={"doc_id": DOC_B})   # fewer rows than n_results
    try:
        await m.query(QUERIES[4])
        out["empty_error"] = None
    except ValueError as e:
        out["empty_error"] = str(e)
    out["batch"] = await m.batch_query(QUERIES, n_results=4)
    out["similar"] = await m.get_similar_documents(DOC_A, "text_0", n_results=3)
    out["cache_stats"] = await m.get_cache_stats()
    st = await m.get_collection_stats()
    out["collection_stats"] = {k: st[k] for k in ("count", "embedding_dim", "batch_size", "stats", "cache")}
    await m.delete_document(DOC_B)
    out["count_after_delete"] = m.collection.count()
    out["query_after_delete"] = await m.query(QUERIES[1], n_results=5)
    await m.delete_all_documents()
    out["count_after_delete_all"] = m.collection.count()
    out["query_on_empty"] = await m.query(QUERIES[0], n_results=5)
    # ---------------- MultiVectorRetriever: id -> key, cache-then-one-pipeline, by-type buckets ----------------
    r = R.MultiVectorRetriever.__new__(R.MultiVectorRetriever)
    r.is_initialized = True
    r.enable_compression = True
    r.max_retries = 1
    r.cache = R.DocumentCache(maxsize=4) if hasattr(R, "DocumentCache") else None
    r.stats = {"total_stored": 0, "total_retrieved": 0, "cache_hits": 0, "cache_misses": 0, "total_deleted": 0}
    r.redis_client = FakeRedis()
    R.AIOREDIS_AVAILABLE = True
    ids = [f"{DOC_A}_{s['id']}" for s in SUMMARIES_A] + ["doc_zzzzzzzzzzzz_text_0", "plain"]
    for s in SUMMARIES_A:
        key = r._item_id_to_redis_key(f"{DOC_A}_{s['id']}")
        r.redis_client.store[key] = gzip.compress(json.dumps({"id": s["id"], "type": s["type"], "raw": s["raw"]}).encode("utf-8"))
    out["redis_keys"] = {i: r._item_id_to_redis_key(i) for i in ids}
    want = [ids[12], ids[0], ids[9], ids[14], ids[3], ids[15], ids[0]]
    out["fetch_ids"] = want
    out["fetch_1"] = await r.retrieve_raw_documents(want)
    out["fetch_2"] = await r.retrieve_raw_documents([ids[0], ids[3], ids[5]])          # two of three now come from the cache
    out["fetch_pipelines"] = r.redis_client.pipelines
    out["fetch_empty"] = await r.retrieve_raw_documents([])
    with open(os.path.join(HERE, "reference_host_logic.json"), "w") as f:
        json.dump(jsonable(out), f, indent=1, sort_keys=True)
    print("wrote reference_host_logic.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    import logging
    logging.disable(logging.CRITICAL)
    asyncio.run(run())
