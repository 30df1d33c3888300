"""CPU: the two consumers of a query result in the reference, restated in
multimodal_rag_b200/manager.py, against the reference's committed ids and the fixture's known answers."""
from multimodal_rag_b200.manager import redis_key_for, sources_from_result


def test_redis_keys_of_the_committed_ids(golden):
    # reference: app/utils/retriever.py:610-637 (split on '_', first two parts = doc id, rest = item id)
    assert redis_key_for("doc_d8164983ea8e_text_35") == "doc:doc_d8164983ea8e:text_35"
    assert redis_key_for("doc_d8164983ea8e_Session01_Khai_niem_co_ban_C_page_19_19130f87") == \
        "doc:doc_d8164983ea8e:Session01_Khai_niem_co_ban_C_page_19_19130f87"
    assert redis_key_for("doc_abc123") == "doc:doc_abc123" and redis_key_for("plain") == "doc:plain"
    for item_id, meta in zip(golden["ids"], golden["metadatas"]):
        parts = item_id.split("_")
        assert redis_key_for(item_id) == f"doc:{'_'.join(parts[:2])}:{'_'.join(parts[2:])}"
        assert redis_key_for(item_id).startswith(f"doc:{meta['doc_id']}:")       # the key is built from the id alone


def test_sources_from_the_known_top5(golden):
    # reference: app/server/api.py:384-396
    top5 = golden["known"]["top5"]
    res = {"ids": [a["id"] for a in top5], "distances": [a["cosine"] for a in top5],
           "metadatas": [{"type": "text"}] * 4 + [None]}
    src = sources_from_result(res)
    assert [s["rank"] for s in src] == [1, 2, 3, 4, 5]
    assert [s["doc_id"] for s in src] == res["ids"]
    assert src[0]["relevance_score"] == round(1 - top5[0]["cosine"], 3) == 0.104
    assert src[4]["type"] == "unknown" and src[0]["type"] == "text"
    far = sources_from_result({"ids": ["x"], "distances": [1.79], "metadatas": [{}]})   # l2^2 of unit vectors can exceed 1
    assert far[0]["relevance_score"] == 0.0
