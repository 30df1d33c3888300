"""EmbeddingManager-shaped facade (multimodal_rag_b200/manager.py) against the reference's call patterns
(app/utils/embedder.py:428-617, 784-930; app/server/api.py:338-396) with an injected deterministic encoder."""
import asyncio
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DIM = 384


def fake_encoder(texts):
    """deterministic unit vectors: similar texts (shared words) -> similar vectors"""
    out = np.zeros((len(texts), DIM), dtype=np.float32)
    for i, t in enumerate(texts):
        for w in t.lower().split():
            seed = int.from_bytes(hashlib.md5(w.encode()).digest()[:4], "little")
            out[i] += np.random.default_rng(seed).standard_normal(DIM).astype(np.float32)
    out /= np.linalg.norm(out, axis=1, keepdims=True) + 1e-30
    return out


def test_manager_upload_query_delete_cycle():
    from multimodal_rag_b200 import B200EmbeddingManager
    from oracle import exact_oracle as eo

    async def run():
        m = B200EmbeddingManager(fake_encoder, "multimodal_rag", space="cosine")
        docs = {}
        for d in range(3):
            doc_id = f"doc_{d:012x}"
            summaries = [{"id": f"text_{i}", "summary": f"chapter {d} section {i} about topic{(d * 7 + i) % 11} and c pointers",
                          "raw": "…", "type": "text"} for i in range(40)]
            summaries += [{"id": f"page_{i}", "summary": f"Image: figure {i} of document {d}", "raw": "…", "type": "image",
                           "path": f"figures/{d}_{i}.png"} for i in range(5)]
            counts = await m.embed_and_store(summaries, doc_id)
            assert counts == {"text": 40, "table": 0, "image": 5}
            docs[doc_id] = summaries
        stats = await m.get_collection_stats()
        assert stats["count"] == 135 and stats["embedding_dim"] == DIM and stats["stats"]["total_items_stored"] == 135

        # /query: flattened dict, ids usable by the Redis key mapping ("doc_<hex12>_<item>")
        r = await m.query("chapter 1 section 3 about topic10 and c pointers", n_results=5)
        assert r["ids"][0] == "doc_000000000001_text_3" and abs(r["distances"][0]) < 1e-5
        assert set(r) == {"ids", "distances", "metadatas", "documents"} and len(r["ids"]) == 5
        assert r["metadatas"][0] == {"doc_id": "doc_000000000001", "item_id": "text_3", "type": "text"}
        assert r["distances"] == sorted(r["distances"])
        relevance = [round(1 - min(d, 1), 3) for d in r["distances"]]            # api.py:390
        assert relevance[0] == 1.0
        with pytest.raises(ValueError):
            await m.query("   ")
        # filter_dict pass-through (where=)
        r = await m.query("Image: figure 2 of document 0", n_results=3, filter_dict={"type": "image"})
        assert r["ids"][0] == "doc_000000000000_page_2" and all(md["type"] == "image" for md in r["metadatas"])

        # batch_query: one device call; must equal the single queries and the oracle
        qs = [f"chapter {d} section {i} about topic{(d * 7 + i) % 11} and c pointers" for d in range(3) for i in (0, 17, 39)] + [""]
        before = m.collection.stats()["launches"]
        batch = await m.batch_query(qs, n_results=4)
        assert m.collection.stats()["launches"] - before <= 8                  # not 10 separate scans
        assert batch[-1]["ids"] == [] and "error" in batch[-1]
        allsum = [s for d in docs.values() for s in d]
        X = eo.normalize_f32(fake_encoder([s["summary"] for s in allsum]))
        ids = [f"{d}_{s['id']}" for d, ss in docs.items() for s in ss]
        er, ed = eo.topk_exact(eo.normalize_f32(fake_encoder(qs[:-1])), X, 4, "cosine")
        for i, q in enumerate(qs[:-1]):
            single = await m.query(q, n_results=4)
            assert batch[i]["ids"] == single["ids"] == [ids[j] for j in er[i]]
            np.testing.assert_allclose(batch[i]["distances"], ed[i], rtol=1e-5, atol=1e-6)

        # get_similar_documents: self excluded, n results
        sim = await m.get_similar_documents("doc_000000000002", "text_5", n_results=3)
        assert len(sim["ids"]) == 3 and "doc_000000000002_text_5" not in sim["ids"]
        with pytest.raises(ValueError):
            await m.get_similar_documents("doc_000000000002", "nope")

        # delete_document / delete_all_documents
        await m.delete_document("doc_000000000001")
        assert (await m.get_collection_stats())["count"] == 90
        r = await m.query("chapter 1 section 3 about topic10 and c pointers", n_results=5)
        assert all(not i.startswith("doc_000000000001") for i in r["ids"])
        await m.delete_all_documents()
        assert (await m.get_collection_stats())["count"] == 0
        assert (await m.query("anything", n_results=5))["ids"] == []

    asyncio.run(run())
