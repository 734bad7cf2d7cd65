"""rag_uq_b200 - B200-native retrieval scoring behind the reference's ``rag_uq`` API.

Import as ``rag_uq_b200`` (the directory name carries the reference repository's name and is
not a valid Python identifier; ``rag_uq_b200/__init__.py`` at the repository root points its
``__path__`` here).  The exported names mirror rag_uq/__init__.py:11-24 for the hot path.
Importing this package loads libragb200.so and fails loudly if it has not been built.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from . import ops  # noqa: F401   (registers the torch.library custom ops)
from .confidence import ConfidenceResult, MCDropoutConfidence, RouterUncertainty
from .engine import HybridEngine, exchange_pools, gather_candidates, global_bm25_statistics, shard_rows
from .retrieval import BM25Index, DenseIndex, Document, HybridRetriever, RetrievalResult, StreamingIndex
from .router import RetrievalRouter, RouterConfig
from .shard_io import load_engine, load_shard, save_engine, save_shard
from .sparse import SegmentedIndex, SparseShard, build_shard, build_shard_blocked

__version__ = "0.1.0"
__all__ = [
    "RetrievalRouter", "RouterConfig", "MCDropoutConfidence", "ConfidenceResult", "RouterUncertainty",
    "HybridRetriever", "StreamingIndex", "BM25Index", "DenseIndex", "Document", "RetrievalResult",
    "HybridEngine", "SparseShard", "SegmentedIndex", "build_shard", "build_shard_blocked", "shard_rows", "gather_candidates", "exchange_pools",
    "global_bm25_statistics", "save_shard", "load_shard", "save_engine", "load_engine",
]
