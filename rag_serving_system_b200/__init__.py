"""B200-native backend for the similarity + top-k hot path of ngray0/RAG-Serving-System.

    from rag_serving_system_b200 import CupyRetriever      # drop-in for rag_service.core.retriever

Only the path needs live here: ``csrc/`` (sm_100a kernels + C ABI), ``_cabi`` (ctypes binding),
``retriever`` (the reference's class surface), ``sharded`` (one-process-per-GPU row sharding).
"""
from .retriever import B200Retriever, CupyRetriever, SimpleRetriever  # noqa: F401

__all__ = ["B200Retriever", "CupyRetriever", "SimpleRetriever"]
