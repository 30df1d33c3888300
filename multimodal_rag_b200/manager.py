"""``EmbeddingManager``-shaped façade over the B200 collection.

Mirrors the public surface of the reference's ``EmbeddingManager`` (``/root/reference/app/utils/embedder.py:83-931``)
for the part of it that is the vector hot path -- same method names, argument meaning, result dictionaries and
error behaviour -- so ``app/server/api.py`` can hold this object instead:

=============================  =====================================================  ==============================
reference method               reference file:line                                    here
=============================  =====================================================  ==============================
``initialize``                 ``embedder.py:152-193``                                ``initialize`` (get-or-create)
``embed_and_store``            ``embedder.py:428-500``  (ids ``f"{doc_id}_{item['id']}"``, metadata ``{doc_id,item_id,type}``)
``query``                      ``embedder.py:539-583``  (empty text -> ``ValueError``; flattened result dict)
``batch_query``                ``embedder.py:784-832``  -- ONE device call for the whole batch instead of a gather of singles
``get_similar_documents``      ``embedder.py:861-930``  (stored vector -> query n+1 -> drop self)
``delete_document``            ``embedder.py:619-656``  (``get(where={'doc_id':…})`` -> ``delete(ids)``)
``delete_all_documents``       ``embedder.py:658-688``
``get_collection_stats``       ``embedder.py:690-728``
=============================  =====================================================  ==============================

The text encoder is NOT part of the hot path (the reference uses sentence-transformers' all-MiniLM-L6-v2 with
``normalize_embeddings=True``, ``embedder.py:385-405``); it is injected as ``encoder(texts) -> [n, dim]`` (numpy array
or CUDA torch tensor -- a tensor stays on the device end to end, skipping the reference's ``.tolist()`` round trip).
Store calls run on ``asyncio.to_thread`` workers exactly like the reference's; the C ABI releases the GIL.
"""
from __future__ import annotations

import asyncio
import hashlib
import logging
import time
from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

from .collection import B200Client

logger = logging.getLogger(__name__)

_EMPTY = {"ids": [], "distances": [], "metadatas": [], "documents": []}


class EmbeddingCache:
    """The reference's per-text embedding cache (``LRUCache``, ``embedder.py:26-80``; key = MD5 of the text,
    ``embedder.py:736-742``): least-recently-used eviction at ``maxsize`` entries, a hit refreshes the entry, hit / miss
    counters with the same ``get_stats()`` dictionary."""

    def __init__(self, maxsize: int = 1000):
        self.maxsize = maxsize
        self._rows: "OrderedDict[str, Any]" = OrderedDict()
        self.hits = self.misses = 0

    @staticmethod
    def key(text: str) -> str:
        return hashlib.md5(text.encode("utf-8")).hexdigest()

    def get(self, key: str):
        row = self._rows.get(key)
        if row is None:
            self.misses += 1
            return None
        self._rows.move_to_end(key)
        self.hits += 1
        return row

    def put(self, key: str, row) -> None:
        if key in self._rows:
            self._rows.move_to_end(key)
        elif len(self._rows) >= self.maxsize:
            self._rows.popitem(last=False)
        self._rows[key] = row

    def clear(self) -> None:
        self._rows.clear()
        self.hits = self.misses = 0

    def __len__(self) -> int:
        return len(self._rows)

    def get_stats(self) -> Dict[str, Any]:
        total = self.hits + self.misses
        return {"size": len(self._rows), "maxsize": self.maxsize, "hits": self.hits, "misses": self.misses,
                "hit_rate": round(self.hits / total, 3) if total else 0.0}


class B200EmbeddingManager:
    def __init__(self, encoder: Callable[[List[str]], Any], collection_name: str = "multimodal_rag", *,
                 space: Optional[str] = None, device: int = 0, max_retries: int = 3, client: Optional[B200Client] = None,
                 model_name: str = "injected-encoder", capacity: int = 0, persist_directory: Optional[str] = None,
                 batch_size: int = 32, enable_cache: bool = True, cache_size: int = 1000):
        self.encoder = encoder
        self.batch_size = batch_size            # texts per encoder call (embedder.py:349-383)
        self.cache = EmbeddingCache(cache_size) if enable_cache else None
        self.collection_name = collection_name
        self.space = space                       # None = Chroma's default (l2), as the reference's create_collection
        self.device = device
        self.max_retries = max_retries
        self.model_name = model_name
        self.capacity = capacity
        self.persist_directory = persist_directory     # settings.CHROMA_PERSIST_DIR in the reference (embedder.py:164-168)
        self.client = client
        self.collection = None
        self.is_initialized = False
        self.stats = {"total_embeddings_created": 0, "total_items_stored": 0, "total_queries": 0,
                      "cache_hits": 0, "cache_misses": 0}

    # ---- lifecycle (embedder.py:152-193) ------------------------------------------------------------
    def _metadata(self):
        md = {"description": "Multimodal RAG document embeddings"}
        if self.space is not None:
            md["hnsw:space"] = self.space
        return md

    async def initialize(self):
        if self.is_initialized:
            return
        if self.client is None:
            self.client = B200Client(device=self.device, default_capacity=self.capacity, path=self.persist_directory)
        try:
            self.collection = await asyncio.to_thread(self.client.get_collection, self.collection_name)
        except ValueError:
            self.collection = await asyncio.to_thread(self.client.create_collection, self.collection_name, self._metadata())
        self.is_initialized = True

    async def persist(self):
        """Write the collection to the persist directory (Chroma persists on its own; here it is an explicit call)."""
        await asyncio.to_thread(self.client.persist)

    async def embed_texts_batch(self, texts: List[str]):
        """``embedder.py:266-347``: texts already in the cache are served from it, the others are encoded in
        ``batch_size`` chunks on worker threads and cached; returns one [n, dim] matrix in the order of `texts` (numpy, or
        a torch tensor on the encoder's device -- such rows never leave the device)."""
        if not self.is_initialized:
            await self.initialize()
        if not texts:
            return []
        rows: List[Any] = [None] * len(texts)
        todo: List[int] = []
        for i, t in enumerate(texts):
            hit = self.cache.get(EmbeddingCache.key(t)) if self.cache is not None else None
            if hit is None:
                todo.append(i)
            else:
                rows[i] = hit
        for b0 in range(0, len(todo), self.batch_size):
            idx = todo[b0: b0 + self.batch_size]
            emb = await asyncio.to_thread(self.encoder, [texts[i] for i in idx])
            for j, i in enumerate(idx):
                rows[i] = emb[j]
                if self.cache is not None:
                    self.cache.put(EmbeddingCache.key(texts[i]), emb[j])
        self.stats["total_embeddings_created"] += len(todo)
        if self.cache is not None:
            self.stats["cache_hits"], self.stats["cache_misses"] = self.cache.hits, self.cache.misses
        return _stack(rows)

    _embed = embed_texts_batch

    async def warmup_cache(self, common_queries: List[str]) -> None:
        """``embedder.py:744-762``"""
        if self.cache is not None:
            await self.embed_texts_batch(list(common_queries))

    async def get_cache_stats(self) -> Dict[str, Any]:
        """``embedder.py:764-772``"""
        return {"enabled": False} if self.cache is None else {"enabled": True, **self.cache.get_stats()}

    async def clear_cache(self) -> None:
        if self.cache is not None:
            self.cache.clear()

    async def _retry(self, fn, *args, **kwargs):
        """3 attempts with 1 s / 2 s back-off, then re-raise (embedder.py:514-537, 592-617)."""
        for attempt in range(self.max_retries):
            try:
                return await asyncio.to_thread(fn, *args, **kwargs)
            except ValueError:
                raise                                   # bad arguments do not get better by retrying
            except Exception as e:                      # noqa: BLE001
                if attempt == self.max_retries - 1:
                    raise
                wait = 2 ** attempt
                logger.warning("store call failed (attempt %d): %s; retrying in %ds", attempt + 1, e, wait)
                await asyncio.sleep(wait)

    # ---- write path (embedder.py:428-537) -----------------------------------------------------------
    async def embed_and_store(self, summaries: List[Dict[str, Any]], doc_id: str) -> Dict[str, int]:
        if not self.is_initialized:
            await self.initialize()
        if not summaries:
            logger.warning("No summaries provided for embedding")
            return {"text": 0, "table": 0, "image": 0}
        t0 = time.time()
        embeddings = await self._embed([item["summary"] for item in summaries])
        documents, metadatas, ids = [], [], []
        counts = {"text": 0, "table": 0, "image": 0}
        for item in summaries:
            documents.append(item["summary"])
            metadatas.append({"doc_id": doc_id, "item_id": item["id"], "type": item["type"]})
            ids.append(f"{doc_id}_{item['id']}")
            if item["type"] in counts:
                counts[item["type"]] += 1
        await self._retry(self.collection.add, ids=ids, embeddings=embeddings, metadatas=metadatas, documents=documents)
        self.stats["total_items_stored"] += len(summaries)
        logger.info("Stored %d embeddings for doc %s in %.3fs", len(ids), doc_id, time.time() - t0)
        return counts

    # ---- read path (embedder.py:539-617, 784-832) ---------------------------------------------------
    @staticmethod
    def _flatten(res, i=0) -> Dict[str, Any]:
        return {"ids": res["ids"][i] if res["ids"] else [],
                "distances": res["distances"][i] if res["distances"] else [],
                "metadatas": res["metadatas"][i] if res["metadatas"] else [],
                "documents": res["documents"][i] if res["documents"] else []}

    async def query(self, query_text: str, n_results: int = 5, filter_dict: Optional[Dict] = None) -> Dict[str, Any]:
        if not self.is_initialized:
            await self.initialize()
        if not query_text or not query_text.strip():
            raise ValueError("Query text cannot be empty")
        emb = await self._embed([query_text])
        res = await self._retry(self.collection.query, query_embeddings=emb, n_results=n_results, where=filter_dict,
                                include=["metadatas", "documents", "distances"])
        self.stats["total_queries"] += 1
        return self._flatten(res)

    async def batch_query(self, queries: List[str], n_results: int = 5,
                          filter_dict: Optional[Dict] = None) -> List[Dict[str, Any]]:
        """The reference gathers independent single queries; here every non-empty text of the batch is embedded
        once and scored in ONE device call.  Per-query failure semantics are kept: an empty text yields the
        empty result with an ``error`` string, as the reference's exception branch does."""
        if not queries:
            return []
        if not self.is_initialized:
            await self.initialize()
        ok = [i for i, q in enumerate(queries) if q and q.strip()]
        out: List[Dict[str, Any]] = [dict(_EMPTY, error="Query text cannot be empty") for _ in queries]
        if ok:
            try:
                emb = await self._embed([queries[i] for i in ok])
                res = await self._retry(self.collection.query, query_embeddings=emb, n_results=n_results,
                                        where=filter_dict, include=["metadatas", "documents", "distances"])
                for j, i in enumerate(ok):
                    out[i] = self._flatten(res, j)
                self.stats["total_queries"] += len(ok)
            except Exception as e:                      # noqa: BLE001  (reference: per-query error entries)
                logger.error("batch query failed: %s", e)
                for i in ok:
                    out[i] = dict(_EMPTY, error=str(e))
        return out

    async def get_similar_documents(self, doc_id: str, item_id: str, n_results: int = 5) -> Dict[str, Any]:
        if not self.is_initialized:
            await self.initialize()
        source_id = f"{doc_id}_{item_id}"
        src = await asyncio.to_thread(self.collection.get, ids=[source_id], include=["embeddings", "documents"])
        if not src["ids"]:
            raise ValueError(f"Item not found: {source_id}")
        res = await asyncio.to_thread(self.collection.query, query_embeddings=[src["embeddings"][0]],
                                      n_results=n_results + 1, include=["metadatas", "documents", "distances"])
        flat = self._flatten(res)
        keep = [i for i, id_ in enumerate(flat["ids"]) if id_ != source_id][:n_results]
        return {key: [flat[key][i] for i in keep] for key in ("ids", "distances", "metadatas", "documents")}

    # ---- deletes / stats (embedder.py:619-728) ------------------------------------------------------
    async def delete_document(self, doc_id: str):
        if not self.is_initialized:
            await self.initialize()
        found = await self._retry(self.collection.get, where={"doc_id": doc_id}, include=[])
        if found["ids"]:
            await self._retry(self.collection.delete, ids=found["ids"])
            logger.info("Deleted %d embeddings for doc %s", len(found["ids"]), doc_id)

    async def delete_all_documents(self):
        if not self.is_initialized:
            await self.initialize()
        await asyncio.to_thread(self.client.delete_collection, self.collection_name)
        self.collection = await asyncio.to_thread(self.client.create_collection, self.collection_name, self._metadata())

    async def get_collection_stats(self) -> Dict[str, Any]:
        if not self.is_initialized:
            await self.initialize()
        try:
            count = await asyncio.to_thread(self.collection.count)
            out = {"name": self.collection_name, "count": count, "model": self.model_name, "device": f"cuda:{self.device}",
                   "embedding_dim": self.collection.dimension, "batch_size": self.batch_size,
                   "stats": {k: self.stats[k] for k in ("total_embeddings_created", "total_items_stored", "total_queries")}}
            if self.cache is not None:
                out["cache"] = self.cache.get_stats()
            engine = getattr(self.collection, "stats", None)
            if callable(engine):
                out["engine"] = engine()
            return out
        except Exception as e:                          # noqa: BLE001
            return {"name": self.collection_name, "count": 0, "error": str(e)}


def _stack(rows: List[Any]):
    """rows of one kind (numpy vectors / lists, or torch tensors on one device) -> one [n, dim] matrix"""
    import sys
    torch = sys.modules.get("torch")
    if torch is not None and rows and isinstance(rows[0], torch.Tensor):
        return torch.stack(list(rows))
    import numpy as np
    return np.asarray(rows, dtype=np.float32)


# ---- what the reference does with a result --------------------------------------------
# Not part of the engine, restated here so the contract it imposes on the engine's output is executable: ids must come
# back verbatim (the Redis key is parsed out of them) and distances must be in the collection's space with "smaller is
# closer", because the server turns them into a relevance score.

def sources_from_result(search_results: Dict[str, Any]) -> List[Dict[str, Any]]:
    """The `sources` list `/query` returns (reference: app/server/api.py:384-396): rank from 1, the Chroma id as
    `doc_id`, `relevance_score = round(1 - min(distance, 1), 3)`, the item's `type` (or 'unknown')."""
    out = []
    triples = zip(search_results["ids"], search_results["distances"], search_results["metadatas"])
    for rank, (item_id, distance, meta) in enumerate(triples, start=1):
        score = 1.0 - (distance if distance < 1.0 else 1.0)
        out.append({"rank": rank, "doc_id": item_id, "relevance_score": round(float(score), 3),
                    "type": (meta or {}).get("type", "unknown")})
    return out


def redis_key_for(item_id: str) -> str:
    """Docstore key of a Chroma id (reference: MultiVectorRetriever._item_id_to_redis_key, app/utils/retriever.py:610-637):
    `doc_<hex>_<item...>` -> `doc:doc_<hex>:<item...>`; ids with fewer than three `_`-separated parts -> `doc:<id>`."""
    head, sep, rest = item_id.partition("_")
    second, sep2, item_part = rest.partition("_")
    if not (sep and sep2):
        return f"doc:{item_id}"
    return f"doc:{head}_{second}:{item_part}"


# ---- the consumer's fetch plan (reference: app/utils/retriever.py:428-574) ------------------
# What `/query` does with the ids right after the vector search (api.py:348): look every id up in the retriever's
# document cache, turn the misses into Redis keys IN THE ORDER OF THE RESULT, fetch them with ONE pipeline (no
# de-duplication: an id that occurs twice is asked for twice), then walk the ids again in result order and bucket the raw
# payloads by their `type`.  Restated here as two pure functions around whatever performs the pipeline, so the whole top-k
# of a batched query can be coalesced into a single round trip.

def plan_raw_fetch(ids: Sequence[str], cached: Optional[Dict[str, Any]] = None) -> Tuple[Dict[str, Any], List[Tuple[str, str]]]:
    """(items already at hand, [(item id, redis key)] still to fetch -- one pipeline GET each, in result order)"""
    have: Dict[str, Any] = {}
    fetch: List[Tuple[str, str]] = []
    for item_id in ids:
        hit = cached.get(item_id) if cached is not None else None
        if hit:
            have[item_id] = hit
        else:
            fetch.append((item_id, redis_key_for(item_id)))
    return have, fetch


def bucket_raw_documents(ids: Sequence[str], items: Dict[str, Any]) -> Dict[str, List[Any]]:
    """{'text_chunks', 'table_chunks', 'image_chunks'}: the raw payload of every id that was found, in result order,
    by the item's type (ids that were not found, or of another type, are skipped) -- retriever.py:497-531."""
    out: Dict[str, List[Any]] = {"text_chunks": [], "table_chunks": [], "image_chunks": []}
    for item_id in ids:
        item = items.get(item_id)
        if item and item.get("type") in ("text", "table", "image"):
            out[item["type"] + "_chunks"].append(item["raw"])
    return out


# ---- the add / query(top_k, use_multimodal) surface BASELINE.json names ---------------------------------------------
class B200Retriever:
    """The retrieval half of the reference's `/upload` and `/query` endpoints as one object:

      add(summaries, doc_id)                     what `/upload` does with the summariser's output (api.py:297 ->
                                                 embedder.py:428-500): ids `f"{doc_id}_{item['id']}"`, metadata
                                                 `{doc_id, item_id, type}`, documents = the summaries
      query(query, top_k=5, use_multimodal=False, filter_dict=None)
                                                 `QueryRequest` (api.py:161-164: 1..2000 characters, top_k in 1..20) ->
                                                 `embedder.query(request.query, n_results=request.top_k)` (api.py:338)
                                                 plus the `sources` list the endpoint returns (api.py:384-396) and the
                                                 docstore fetch plan (api.py:348 -> retriever.py:428-574).

    `use_multimodal` does NOT change retrieval in the reference: `/query` passes no filter whatever its value
    (api.py:338) and only uses the flag, after the search, to pick the generator (api.py:356).  It is accepted and echoed
    for that reason; restricting the search to a type is `filter_dict={"type": ...}` (the `where=` pass-through of
    embedder.py:543,599).
    """

    MAX_QUERY_CHARS, MAX_TOP_K = 2000, 20

    def __init__(self, manager: B200EmbeddingManager):
        self.manager = manager

    async def add(self, summaries: List[Dict[str, Any]], doc_id: str) -> Dict[str, int]:
        return await self.manager.embed_and_store(summaries, doc_id)

    def _check(self, query: str, top_k: int):
        if not isinstance(query, str) or not (1 <= len(query) <= self.MAX_QUERY_CHARS):
            raise ValueError(f"query must be a string of 1..{self.MAX_QUERY_CHARS} characters")
        if not isinstance(top_k, int) or isinstance(top_k, bool) or not (1 <= top_k <= self.MAX_TOP_K):
            raise ValueError(f"top_k must be an integer in 1..{self.MAX_TOP_K}")

    @staticmethod
    def _shape(res: Dict[str, Any], use_multimodal: bool) -> Dict[str, Any]:
        have, fetch = plan_raw_fetch(res["ids"])
        return {**res, "sources": sources_from_result(res) if res["ids"] else [],
                "redis_keys": [key for _, key in fetch], "use_multimodal": bool(use_multimodal)}

    async def query(self, query: str, top_k: int = 5, use_multimodal: bool = False,
                    filter_dict: Optional[Dict] = None) -> Dict[str, Any]:
        self._check(query, top_k)
        res = await self.manager.query(query, n_results=top_k, filter_dict=filter_dict)
        return self._shape(res, use_multimodal)

    async def batch_query(self, queries: List[str], top_k: int = 5, use_multimodal: bool = False,
                          filter_dict: Optional[Dict] = None) -> List[Dict[str, Any]]:
        """Many requests at once: ONE device call scores the whole batch (B200EmbeddingManager.batch_query)."""
        for q in queries:
            if q and q.strip():
                self._check(q, top_k)
        out = []
        for res in await self.manager.batch_query(queries, n_results=top_k, filter_dict=filter_dict):
            shaped = self._shape(res, use_multimodal)
            if "error" in res:
                shaped["error"] = res["error"]
            out.append(shaped)
        return out
