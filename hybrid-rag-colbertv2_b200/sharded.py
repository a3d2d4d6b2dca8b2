"""Document-sharded search: one process per GPU, local top-k, all-gather of k keys, on-device merge.

New design (the reference is single-process): the corpus shards by document, each
rank scores its shard and emits k 64-bit (score, GLOBAL doc id) keys; one all-gather of k*8 bytes per
rank crosses NVLink (latency-bound: 800 B at k=100), and every rank merges the world_size*k keys with
hrc_topk_merge.  Keys are totally ordered, so the merged list equals the single-GPU list exactly.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

from . import _lib
from .retriever import JinaColBERTRetriever


def all_gather_keys(keys: torch.Tensor, k: int, group=None) -> torch.Tensor:
    """keys: int64 [Bq, k_local] (k_local <= k) -> int64 [Bq, world * k], empty slots = 0.

    Works on any backend (NCCL for CUDA tensors, gloo for CPU tensors in the host-logic tests).
    """
    world = dist.get_world_size(group)
    bq, k_local = keys.shape
    if k_local < k:  # a shard with fewer than k documents pads with the empty-slot sentinel
        pad = torch.zeros((bq, k - k_local), dtype=keys.dtype, device=keys.device)
        keys = torch.cat([keys, pad], 1)
    keys = keys.contiguous()
    out = torch.empty((world, bq, k), dtype=keys.dtype, device=keys.device)
    dist.all_gather_into_tensor(out, keys, group=group) if keys.is_cuda else dist.all_gather(
        list(out.unbind(0)), keys, group=group)
    return out.permute(1, 0, 2).reshape(bq, world * k).contiguous()


class ShardedSearcher:
    """search over a corpus sharded across the ranks of a torch.distributed group (one rank per GPU)."""

    def __init__(self, retriever: JinaColBERTRetriever, group=None):
        self.retriever = retriever          # holds THIS rank's shard; store.doc_id_base makes ids global
        self.group = group
        self._pinned = None

    def search_keys(self, query_embeddings: torch.Tensor, k: int) -> torch.Tensor:
        local = self.retriever.search_keys(query_embeddings, k)       # [Bq, min(k, n_local)]
        gathered = all_gather_keys(local, k, self.group)              # [Bq, world * k]
        return _lib.topk_merge(gathered, k)                           # [Bq, k] sorted, 0 = empty

    def _lq(self, query_embeddings: torch.Tensor) -> int:
        return int(query_embeddings.shape[-2])

    def search_embeddings(self, query_embeddings: torch.Tensor, k: int = 10) -> Tuple[torch.Tensor, torch.Tensor]:
        ids, scores = _lib.keys_unpack(self.search_keys(query_embeddings, k))
        return ids, self.retriever._finish_scores(scores, self._lq(query_embeddings))   # honours score_reduction="mean"

    def search_host(self, query_embeddings: torch.Tensor, k: int = 10, copy: bool = True
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Host query embedding (fp32 CPU [Bq, Lq, 128], ideally pinned) -> host (ids, scores): asynchronous H2D,
        the sharded search, asynchronous D2H into pinned buffers, ONE stream synchronisation.  The returned tensors
        belong to the caller; copy=False returns the pinned staging buffers, which the next call overwrites."""
        dev = self.retriever.device
        q = query_embeddings if query_embeddings.dim() == 3 else query_embeddings.unsqueeze(0)
        ids, scores = _lib.keys_unpack(self.search_keys(q.to(dev, non_blocking=True), k))
        if self._pinned is None or self._pinned[0].shape != ids.shape:
            self._pinned = (torch.empty(ids.shape, dtype=torch.int32).pin_memory(),
                            torch.empty(scores.shape, dtype=torch.float32).pin_memory())
        self._pinned[0].copy_(ids, non_blocking=True)
        self._pinned[1].copy_(scores, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        out_ids, out_sc = (self._pinned[0].clone(), self._pinned[1].clone()) if copy else self._pinned
        return out_ids, self.retriever._finish_scores(out_sc, self._lq(q))
