"""Chunk-id mapping and storage fetch for the candidates the retrievers return.

The reference uses the 0-based corpus list index as `chunk_id` (local_rag_complete.py:956, :948) but looks chunks up by
the SQLite autoincrement primary key `Chunk.id` (1-based, :118, :984), one query per id (:983-984), and silently drops
ids the table does not hold (:985).  Here the mapping is explicit and the fetch is ONE `IN (...)` query:

    fetch = SqliteChunkFetcher("rag_local.db", ChunkIdMap.autoincrement(n_chunks))       # corpus index i <-> Chunk.id i + 1
    hybrid = HybridRetriever(config, indexer, None, chunk_fetcher=fetch)

Storage is outside the hot path: plain `sqlite3`, no ORM.
"""
from __future__ import annotations

import json
import sqlite3
from typing import Dict, Iterable, List, Optional, Sequence


class ChunkIdMap:
    """corpus index (what the retrievers rank) <-> external chunk id (what the storage keys on)."""

    def __init__(self, external_ids: Sequence[int]):
        self.external = [int(x) for x in external_ids]
        self.index = {e: i for i, e in enumerate(self.external)}
        if len(self.index) != len(self.external):
            raise ValueError("external chunk ids must be unique")

    @classmethod
    def autoincrement(cls, n: int, first: int = 1) -> "ChunkIdMap":
        """The reference's situation: chunks inserted in corpus order into a table with an autoincrement key."""
        return cls(range(first, first + n))

    def __len__(self) -> int:
        return len(self.external)

    def to_external(self, indices: Iterable[int]) -> List[Optional[int]]:
        return [self.external[i] if 0 <= i < len(self.external) else None for i in indices]

    def to_index(self, external_ids: Iterable[int]) -> List[Optional[int]]:
        return [self.index.get(int(e)) for e in external_ids]


class SqliteChunkFetcher:
    """`chunk_fetcher` for HybridRetriever: corpus indices in, the reference's chunk dicts out (:986-993), in the
    requested order, ids the table does not hold dropped (:985) — with one query for the whole candidate list."""

    COLUMNS = "id, text, document_id, heading_path, has_images, metadata"

    def __init__(self, db_path: str, id_map: Optional[ChunkIdMap] = None, table: str = "chunks"):
        if not table.replace("_", "").isalnum():
            raise ValueError("bad table name")
        self.db_path, self.id_map, self.table = db_path, id_map, table

    def __call__(self, chunk_ids: List[int]) -> List[Dict]:
        if not chunk_ids:
            return []
        ext = self.id_map.to_external(chunk_ids) if self.id_map is not None else [int(c) for c in chunk_ids]
        wanted = [e for e in ext if e is not None]
        rows = {}
        if wanted:
            con = sqlite3.connect(self.db_path)
            try:
                marks = ",".join("?" * len(wanted))
                for r in con.execute(f"SELECT {self.COLUMNS} FROM {self.table} WHERE id IN ({marks})", wanted):
                    rows[int(r[0])] = r
            finally:
                con.close()
        out = []
        for cid, e in zip(chunk_ids, ext):
            r = rows.get(e) if e is not None else None
            if r is None:
                continue
            out.append({'chunk_id': int(cid), 'text': r[1], 'document_id': r[2], 'heading_path': r[3],
                        'has_images': bool(r[4]), 'metadata': json.loads(r[5]) if r[5] else {}})
        return out
