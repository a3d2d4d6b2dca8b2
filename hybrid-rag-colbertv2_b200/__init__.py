"""hybrid-rag-colbertv2_b200 — B200-native MaxSim late-interaction scoring, top-k and RRF behind the
JinaColBERTRetriever / DualIndexer / HybridRetriever API of techmum21p/hybrid-rag-ColBERTv2.

Import as `hybrid_rag_colbertv2_b200` (the repo-root shim makes the hyphenated directory importable).
"""
from . import _lib
from ._lib import HrcError, PATH_AUTO, PATH_SIMT, PATH_TC, PATH_TC_DM
from .chunks import ChunkIdMap, SqliteChunkFetcher
from .encoder import ColBERTEncoder, SyntheticEncoder
from .retriever import DualIndexer, GraphedRerank, HybridRetriever, JinaColBERTRetriever, RAGConfig, install
from .sharded import ShardedSearcher, all_gather_keys
from .store import PackedStore, lengths_to_offsets, shard_doc_ranges

__all__ = [
    "DualIndexer", "HybridRetriever", "JinaColBERTRetriever", "RAGConfig", "PackedStore", "SyntheticEncoder", "ColBERTEncoder", "ChunkIdMap", "SqliteChunkFetcher",
    "ShardedSearcher", "all_gather_keys", "lengths_to_offsets", "shard_doc_ranges", "HrcError",
    "PATH_AUTO", "PATH_SIMT", "PATH_TC", "PATH_TC_DM", "install", "GraphedRerank",
]
__version__ = "0.2.0"
