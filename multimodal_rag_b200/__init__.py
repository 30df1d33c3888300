"""B200-native exact vector retrieval behind the reference's Chroma collection seam.

Public surface:
  B200Client / B200Collection   -- chromadb.Client / Collection stand-ins (collection.py)
  B200EmbeddingManager          -- app/utils/embedder.py EmbeddingManager-compatible shim (manager.py)
  B200Retriever                 -- add(summaries, doc_id) / query(query, top_k, use_multimodal, filter_dict): the retrieval
                                   half of the reference's /upload and /query endpoints (manager.py)
  ShardedCollection             -- row-sharded multi-GPU collection over torch.distributed (sharded.py)
The compute lives in libb2r.so (csrc/, C ABI in include/b2r.h); nothing here computes on the CPU.
"""
from .collection import B200Client, B200Collection  # noqa: F401
from .manager import B200EmbeddingManager, B200Retriever  # noqa: F401

__all__ = ["B200Client", "B200Collection", "B200EmbeddingManager", "B200Retriever"]
__version__ = "0.1.0"
