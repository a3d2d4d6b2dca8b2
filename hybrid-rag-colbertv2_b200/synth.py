"""Synthetic corpora / queries of the shapes BASELINE.json names (encoder weights are unavailable offline).

Everything is a pure function of (seed, GLOBAL document / token index), so any shard layout
reproduces the same global corpus.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from .store import DIM, PackedStore, shard_doc_ranges

BASE_SEED = 20260101


def doc_lengths(n_docs: int, min_len: int, max_len: int, seed: int) -> np.ndarray:
    """Per-document token counts, U{min_len..max_len} from a counter hash of the global doc index."""
    if min_len == max_len:
        return np.full(n_docs, min_len, dtype=np.int64)
    z = (np.arange(n_docs, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (np.int64(min_len) + (z % np.uint64(max_len - min_len + 1)).astype(np.int64))


def synth_store(n_docs: int, min_len: int, max_len: int, seed: int = BASE_SEED, device="cuda", rank: int = 0,
                world_size: int = 1, chunk_tokens: int = 1 << 24) -> PackedStore:
    """This rank's document shard of the global synthetic corpus, generated on the device in chunks."""
    lens = doc_lengths(n_docs, min_len, max_len, seed)
    off = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    d0, d1 = shard_doc_ranges(off, world_size)[rank]
    t0, t1 = int(off[d0]), int(off[d1])
    tokens = torch.empty((t1 - t0, DIM), dtype=torch.bfloat16, device=device)
    for b in range(0, t1 - t0, chunk_tokens):
        e = min(b + chunk_tokens, t1 - t0)
        _lib.synth_tokens(tokens[b:e], t0 + b, seed)
    offsets = torch.from_numpy(off[d0:d1 + 1] - off[d0]).to(device)
    return PackedStore(tokens, offsets, doc_id_base=d0)


def synth_queries(n_queries: int, lq: int = 32, seed: int = BASE_SEED + 7, device="cpu") -> torch.Tensor:
    """Random-init L2-normalised query token embeddings, bf16 [n_queries, lq, 128]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn((n_queries, lq, DIM), generator=g), dim=-1)
    return q.to(torch.bfloat16).to(device)


def plant(store: PackedStore, queries: torch.Tensor, n_planted: int = 200, seed: int = BASE_SEED + 13,
          n_docs_global: Optional[int] = None) -> torch.Tensor:
    """Overwrite tokens of `n_planted` documents per query with noisy copies of that query's tokens, so
    the top-k is well separated ("planted" distribution).  Deterministic in the GLOBAL
    doc id; only documents of this shard are touched.  Returns the planted global ids [n_queries, n_planted].
    """
    n_global = n_docs_global if n_docs_global is not None else store.doc_id_base + store.n_docs
    nq, lq, _ = queries.shape
    g = torch.Generator(device="cpu").manual_seed(seed)
    ids = torch.stack([torch.randperm(n_global, generator=g)[:n_planted] for _ in range(nq)])
    lens = store.lengths().cpu()
    off = store.offsets.cpu()
    qf = queries.float().cpu()
    for qi in range(nq):
        for j, gid in enumerate(ids[qi].tolist()):
            d = gid - store.doc_id_base
            if not (0 <= d < store.n_docs):
                continue
            gg = torch.Generator(device="cpu").manual_seed(seed * 1000003 + gid)
            n_tok = min(int(lens[d]), lq)
            noise_norm = 0.5 + 1.5 * (j / max(n_planted - 1, 1))  # cos 0.89 .. 0.45: graded, well above background
            rows = qf[qi, :n_tok] + noise_norm * torch.randn((n_tok, DIM), generator=gg) / (DIM ** 0.5)
            rows = torch.nn.functional.normalize(rows, dim=-1).to(torch.bfloat16)
            t0 = int(off[d])
            store.tokens[t0:t0 + n_tok] = rows.to(store.device)
    return ids
