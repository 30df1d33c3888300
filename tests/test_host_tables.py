"""Host-side behaviour of the collection that needs no device: argument errors raised before any CUDA call."""
import pytest


def test_delete_without_ids_or_where_raises_instead_of_wiping():
    # chromadb 0.4.22 raises when none of ids / where / where_document is given; the reference empties a collection through
    # client.delete_collection (app/utils/embedder.py:669-678), never through delete()
    from multimodal_rag_b200 import B200Collection
    c = B200Collection("t", {"hnsw:space": "cosine"})
    with pytest.raises(ValueError):
        c.delete()
    with pytest.raises(ValueError):
        c.delete(ids=[])
    with pytest.raises(ValueError):
        c.delete(where={})


def test_embedding_cache_is_lru_with_the_reference_counters():
    from multimodal_rag_b200.manager import EmbeddingCache
    c = EmbeddingCache(2)
    c.put("a", [1.0]); c.put("b", [2.0])
    assert c.get("a") == [1.0]                  # refreshes a
    c.put("c", [3.0])                           # evicts b
    assert c.get("b") is None and c.get("c") == [3.0]
    assert c.get_stats() == {"size": 2, "maxsize": 2, "hits": 2, "misses": 1, "hit_rate": 0.667}
    assert EmbeddingCache.key("hello") == "5d41402abc4b2a76b9719d911017c592"
    c.clear()
    assert c.get_stats() == {"size": 0, "maxsize": 2, "hits": 0, "misses": 0, "hit_rate": 0.0}
