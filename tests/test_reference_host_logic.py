"""The façade (multimodal_rag_b200/manager.py) against the REFERENCE'S OWN host logic.

tests/golden/reference_host_logic.json was produced by running the reference's unmodified `EmbeddingManager` and
`MultiVectorRetriever` (tests/golden/make_reference_golden.py: its Chroma collection replaced by the oracle's exact
collection, its encoder by tests/fake_encoder.py, its Redis by an in-memory fake).  The same scenario is replayed here
through `B200EmbeddingManager` / `B200Retriever`:
  * on the CPU with the oracle's collection injected (host logic only: ids, metadata, cache, flattening, error entries,
    self-exclusion, deletes, stats, Redis keys, fetch plan, buckets) -- every value must be identical;
  * on the GPU with the real collection (`-m gpu`): ids / metadatas / documents identical, distances to 1e-5 relative.
"""
import asyncio
import json
import os

import numpy as np
import pytest

from fake_encoder import DOC_A, DOC_B, QUERIES, SUMMARIES_A, SUMMARIES_B, fake_embed

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ref():
    return json.load(open(os.path.join(HERE, "golden", "reference_host_logic.json")))


class OracleClient:
    """chromadb.Client's trio over the oracle's exact collection (test infrastructure)."""

    def __init__(self):
        self.c = {}

    def get_collection(self, name):
        if name not in self.c:
            raise ValueError(f"Collection {name} does not exist.")
        return self.c[name]

    def create_collection(self, name, metadata=None):
        from oracle.exact_oracle import ExactCollection
        self.c[name] = ExactCollection(name, metadata)
        return self.c[name]

    def delete_collection(self, name):
        del self.c[name]


async def _scenario(client):
    from multimodal_rag_b200.manager import B200EmbeddingManager
    m = B200EmbeddingManager(fake_embed, client=client, batch_size=4, enable_cache=True, cache_size=8)
    out = {}
    out["counts_a"] = await m.embed_and_store(SUMMARIES_A, DOC_A)
    out["counts_b"] = await m.embed_and_store(SUMMARIES_B, DOC_B)
    out["stored"] = m.collection.get(include=["metadatas", "documents"])
    out["query_1"] = await m.query(QUERIES[0], n_results=5)
    out["query_1_again"] = await m.query(QUERIES[0], n_results=5)
    out["query_image"] = await m.query(QUERIES[0], n_results=3, filter_dict={"type": "image"})
    out["query_doc_b"] = await m.query(QUERIES[1], n_results=20, filter_dict={"doc_id": DOC_B})
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
    return out, m


def _same_result(got, want, exact_dist):
    assert got["ids"] == want["ids"]
    assert got["metadatas"] == want["metadatas"] and got["documents"] == want["documents"]
    if exact_dist:
        assert got["distances"] == want["distances"]
    else:
        np.testing.assert_allclose(got["distances"], want["distances"], rtol=1e-5, atol=1e-7)
    assert got.get("error") == want.get("error")


def _compare(out, ref, exact_dist):
    assert out["counts_a"] == ref["counts_a"] and out["counts_b"] == ref["counts_b"]
    assert out["stored"]["ids"] == ref["stored"]["ids"]
    assert out["stored"]["metadatas"] == ref["stored"]["metadatas"] and out["stored"]["documents"] == ref["stored"]["documents"]
    for key in ("query_1", "query_1_again", "query_image", "query_doc_b", "similar", "query_after_delete", "query_on_empty"):
        _same_result(out[key], ref[key], exact_dist)
    assert out["empty_error"] == ref["empty_error"]
    assert len(out["batch"]) == len(ref["batch"])
    for g, w in zip(out["batch"], ref["batch"]):
        _same_result(g, w, exact_dist)
    assert out["cache_stats"] == ref["cache_stats"]
    assert out["collection_stats"] == ref["collection_stats"]
    assert out["count_after_delete"] == ref["count_after_delete"] and out["count_after_delete_all"] == ref["count_after_delete_all"]


def test_manager_host_logic_equals_the_reference(ref):
    out, _ = asyncio.run(_scenario(OracleClient()))
    _compare(out, ref, exact_dist=True)


def test_redis_keys_fetch_plan_and_buckets_equal_the_reference(ref):
    from multimodal_rag_b200.manager import bucket_raw_documents, plan_raw_fetch, redis_key_for
    for item_id, key in ref["redis_keys"].items():
        assert redis_key_for(item_id) == key
    # the docstore as the golden script filled it
    store = {redis_key_for(f"{DOC_A}_{s['id']}"): {"id": s["id"], "type": s["type"], "raw": s["raw"]} for s in SUMMARIES_A}
    cache = {}
    pipelines = []
    for want_ids, want_out in ((ref["fetch_ids"], ref["fetch_1"]), ([f"{DOC_A}_text_0", f"{DOC_A}_text_3", f"{DOC_A}_text_5"], ref["fetch_2"])):
        have, fetch = plan_raw_fetch(want_ids, cache)
        pipelines.append([key for _, key in fetch])
        got = {i: store[k] for i, k in fetch if k in store}           # what ONE pipeline returns
        cache.update(got)
        assert bucket_raw_documents(want_ids, {**have, **got}) == want_out
    assert pipelines == ref["fetch_pipelines"]                        # same keys, same order, no de-duplication
    assert bucket_raw_documents([], {}) == ref["fetch_empty"]


def test_retriever_facade_surface(ref):
    from multimodal_rag_b200.manager import B200EmbeddingManager, B200Retriever, redis_key_for

    async def go():
        m = B200EmbeddingManager(fake_embed, client=OracleClient(), batch_size=4, cache_size=8)
        r = B200Retriever(m)
        assert await r.add(SUMMARIES_A, DOC_A) == ref["counts_a"]
        await r.add(SUMMARIES_B, DOC_B)
        a = await r.query(QUERIES[0], top_k=5)
        b = await r.query(QUERIES[0], top_k=5, use_multimodal=True)          # the flag does not touch retrieval (api.py:338,356)
        assert a["ids"] == b["ids"] == ref["query_1"]["ids"] and a["distances"] == ref["query_1"]["distances"]
        assert b["use_multimodal"] is True and a["use_multimodal"] is False
        assert [s["doc_id"] for s in a["sources"]] == a["ids"] and [s["rank"] for s in a["sources"]] == [1, 2, 3, 4, 5]
        assert a["redis_keys"] == [redis_key_for(i) for i in a["ids"]]
        img = await r.query(QUERIES[0], top_k=3, filter_dict={"type": "image"})
        assert img["ids"] == ref["query_image"]["ids"]
        for bad in ({"query": "", "top_k": 5}, {"query": "x" * 2001, "top_k": 5}, {"query": "q", "top_k": 0}, {"query": "q", "top_k": 21}):
            with pytest.raises(ValueError):
                await r.query(**bad)
        many = await r.batch_query(QUERIES, top_k=4)
        assert [x["ids"] for x in many] == [x["ids"] for x in ref["batch"]]
        assert many[2]["error"] == ref["batch"][2]["error"] and many[2]["sources"] == []
    asyncio.run(go())


@pytest.mark.gpu
def test_manager_on_the_gpu_equals_the_reference(ref):
    from multimodal_rag_b200 import B200Client
    out, m = asyncio.run(_scenario(B200Client()))
    _compare(out, ref, exact_dist=False)
