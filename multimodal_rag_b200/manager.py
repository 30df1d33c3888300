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
import logging
import time
from typing import Any, Callable, Dict, List, Optional

from .collection import B200Client

logger = logging.getLogger(__name__)

_EMPTY = {"ids": [], "distances": [], "metadatas": [], "documents": []}


class B200EmbeddingManager:
    def __init__(self, encoder: Callable[[List[str]], Any], collection_name: str = "multimodal_rag", *,
                 space: Optional[str] = None, device: int = 0, max_retries: int = 3, client: Optional[B200Client] = None,
                 model_name: str = "injected-encoder", capacity: int = 0, persist_directory: Optional[str] = None):
        self.encoder = encoder
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
        self.stats = {"total_embeddings_created": 0, "total_items_stored": 0, "total_queries": 0}

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

    async def _embed(self, texts: List[str]):
        emb = await asyncio.to_thread(self.encoder, texts)
        self.stats["total_embeddings_created"] += len(texts)
        return emb

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
            return {"name": self.collection_name, "count": count, "model": self.model_name, "device": f"cuda:{self.device}",
                    "embedding_dim": self.collection.dimension, "stats": dict(self.stats), "engine": self.collection.stats()}
        except Exception as e:                          # noqa: BLE001
            return {"name": self.collection_name, "count": 0, "error": str(e)}


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
