"""Document-sharded search: one process per GPU, local top-k, exchange of k keys per rank, on-device merge.

New design (the reference is single-process): the corpus shards by document, each
rank scores its shard and emits k 64-bit (score, GLOBAL doc id) keys; the ranks exchange k*8 bytes each
over NVLink (latency-bound: 800 B at k=100) and every rank merges the world_size*k keys.  Keys are
totally ordered, so the merged list equals the single-GPU list exactly.

Transports of the exchange:
  "nccl"   (default) ncclAllGather inside libhrc.so — the whole sharded step is ONE C call (hrc_sharded_search)
  "p2p"    direct peer stores into CUDA-IPC-mapped receive buffers + sequence flags, also inside libhrc.so: the
           exchange is fused into the producer / merge kernels, no collective launch
  "torch"  torch.distributed all_gather + hrc_topk_merge — backend-agnostic (gloo on CPU tensors in the host-logic
           tests); the cross-check of the two library transports
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .retriever import JinaColBERTRetriever, _knob

_TRANSPORTS = {"nccl": _lib.TRANSPORT_NCCL, "p2p": _lib.TRANSPORT_P2P}


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


class PendingKeys:
    """The result of ShardedSearcher.search_keys_async: merged keys that become valid on the exchange stream."""

    def __init__(self, keys: torch.Tensor, done: torch.cuda.Event):
        self._keys, self._done = keys, done

    def result(self) -> torch.Tensor:
        """The merged keys, ordered after the exchange on the CURRENT stream (no host synchronisation)."""
        torch.cuda.current_stream(self._keys.device).wait_event(self._done)
        return self._keys


class ShardedSearcher:
    """search / hybrid retrieve over a corpus sharded across the ranks of a torch.distributed group (one rank per GPU).

    transport: "auto" (default: the exchange runs over peer memory inside libhrc when every rank can map its peers, else
    over ncclAllGather inside libhrc), "p2p", "nccl", or "torch" (torch.distributed collectives; the only one that works
    without CUDA, used by the CPU tests and as the cross-check)."""

    def __init__(self, retriever: JinaColBERTRetriever, group=None, transport: str = "auto", p2p_max_keys: int = 1 << 16):
        if transport not in ("auto", "nccl", "p2p", "torch"):
            raise ValueError(f"transport must be 'auto', 'nccl', 'p2p' or 'torch', got {transport!r}")
        self.retriever = retriever          # holds THIS rank's shard; store.doc_id_base makes ids global
        self.group = group
        self.transport = transport
        self.comm: Optional[_lib.Comm] = None
        self._ws = _lib.Workspace()
        self._host: Optional[_lib.ShardedHostSearch] = None
        self._pinned = None
        self._n_global: Optional[int] = None
        self._side: Optional[torch.cuda.Stream] = None       # exchange stream of search_keys_async
        self._async_done: Optional[torch.cuda.Event] = None  # the last exchange issued on it
        self._side_ws = _lib.Workspace()
        if transport == "auto":
            # peer memory over NVLink when EVERY rank can map every other rank's buffer (hrc_comm_enable_p2p fails on all
            # ranks or on none), else NCCL: same results, two launches per single-query search instead of three
            self.comm = _lib.Comm(retriever.device, group)
            try:
                self.comm.enable_p2p(p2p_max_keys)
                self.transport = "p2p"
            except _lib.HrcError:
                self.transport = "nccl"
            self._host = _lib.ShardedHostSearch(self.comm)
        elif transport != "torch":
            self.comm = _lib.Comm(retriever.device, group, p2p_max_keys=p2p_max_keys if transport == "p2p" else 0)
            self._host = _lib.ShardedHostSearch(self.comm)

    # ---- helpers ---------------------------------------------------------------------------------
    def _lq(self, query_embeddings: torch.Tensor) -> int:
        return int(query_embeddings.shape[-2])

    def _path(self) -> int:
        return _knob(self.retriever.config, "maxsim_path")

    def _join_async(self) -> None:
        """Exchanges of one communicator must run in issue order on every rank (the peer-memory slots are double-buffered
        by step parity, NCCL wants one order per communicator): an exchange issued on the CURRENT stream first waits for
        the exchanges search_keys_async put on the side stream."""
        if self._async_done is not None:
            torch.cuda.current_stream(self.retriever.device).wait_event(self._async_done)
            self._async_done = None

    def n_docs_global(self) -> int:
        """Documents of the whole corpus (sum over the ranks' shards), cached."""
        if self._n_global is None:
            t = torch.tensor([self.retriever.store.n_docs], dtype=torch.int64,
                             device=self.retriever.device if dist.get_backend(self.group) == "nccl" else "cpu")
            dist.all_reduce(t, group=self.group)
            self._n_global = int(t[0])
        return self._n_global

    # ---- search ----------------------------------------------------------------------------------
    def search_keys(self, query_embeddings: torch.Tensor, k: int) -> torch.Tensor:
        """Merged, sorted (score, GLOBAL id) keys: int64 [Bq, k] (0 = empty), identical on every rank."""
        r = self.retriever
        if self.transport == "torch" or r._literal():
            local = r.search_keys(query_embeddings, k)                    # [Bq, min(k, n_local)]
            gathered = all_gather_keys(local, k, self.group)              # [Bq, world * k]
            return _lib.topk_merge(gathered, k)                           # [Bq, k] sorted, 0 = empty
        q = r._prep_queries(query_embeddings)
        s = r.store
        self._join_async()
        return _lib.sharded_search(self.comm, s.tokens, s.offsets, q, int(k), id_base=s.doc_id_base, path=self._path(),
                                   transport=_TRANSPORTS[self.transport], workspace=self._ws, unpack=False)[0]

    def search_keys_async(self, query_embeddings: torch.Tensor, k: int) -> PendingKeys:
        """Pipelined form of search_keys for a stream of independent queries: the local MaxSim + top-k runs on the
        current stream, the exchange + merge on a dedicated stream, so the next query's scan of the shard does not wait
        for this query's collective (nor for the slowest rank of this step).  Every rank must issue the same sequence of
        calls.  `.result()` orders the merged keys after the exchange on the current stream."""
        if self.transport == "torch":
            raise ValueError("search_keys_async needs a library transport ('nccl' or 'p2p')")
        r = self.retriever
        dev = r.device
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        q = r._prep_queries(query_embeddings)
        s = r.store
        k_local = min(int(k), s.n_docs)
        local = torch.zeros((q.shape[0], int(k)), dtype=torch.int64, device=dev) if k_local < k else None
        if k_local > 0:
            got = _lib.search(s.tokens, s.offsets, q, k_local, id_base=s.doc_id_base, path=self._path(),
                              workspace=self._ws, unpack=False)[0]
            if local is None:
                local = got
            else:
                local[:, :k_local] = got
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        local.record_stream(self._side)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            merged = _lib.allgather_merge_topk(self.comm, local, int(k), transport=_TRANSPORTS[self.transport],
                                               workspace=self._side_ws)[0]
            done = torch.cuda.Event()
            done.record(self._side)
        self._async_done = done
        merged.record_stream(torch.cuda.current_stream(dev))
        return PendingKeys(merged, done)

    def search_embeddings(self, query_embeddings: torch.Tensor, k: int = 10) -> Tuple[torch.Tensor, torch.Tensor]:
        r = self.retriever
        if self.transport == "torch" or r._literal():
            ids, scores = _lib.keys_unpack(self.search_keys(query_embeddings, k))
        else:
            q = r._prep_queries(query_embeddings)
            s = r.store
            self._join_async()
            _, ids, scores = _lib.sharded_search(self.comm, s.tokens, s.offsets, q, int(k), id_base=s.doc_id_base,
                                                 path=self._path(), transport=_TRANSPORTS[self.transport],
                                                 workspace=self._ws, unpack=True)
        return ids, r._finish_scores(scores, self._lq(query_embeddings))   # honours score_reduction="mean"

    def search_host(self, query_embeddings: torch.Tensor, k: int = 10, copy: bool = True
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Host query embedding (fp32 CPU [Bq, Lq, 128], ideally pinned) -> host (ids, scores), ONE stream
        synchronisation.  With the library transports this is one C call (hrc_sharded_search_host: H2D, fp32 -> bf16,
        local search, exchange, merge + unpack, D2H).  The returned tensors belong to the caller; copy=False returns
        the pinned staging buffers, which the next call overwrites."""
        r = self.retriever
        dev = r.device
        q = query_embeddings if query_embeddings.dim() == 3 else query_embeddings.unsqueeze(0)
        if self.transport != "torch" and not r._literal() and not q.is_cuda:
            s = r.store
            self._join_async()
            ids, sc = self._host(s.tokens, s.offsets, q.to(torch.float32).contiguous(), int(k), id_base=s.doc_id_base,
                                 path=self._path(), transport=_TRANSPORTS[self.transport], copy=copy)
            return ids, r._finish_scores(sc, self._lq(q))
        ids, scores = _lib.keys_unpack(self.search_keys(q.to(dev, non_blocking=True), k))
        if self._pinned is None or self._pinned[0].shape != ids.shape:
            self._pinned = (torch.empty(ids.shape, dtype=torch.int32).pin_memory(),
                            torch.empty(scores.shape, dtype=torch.float32).pin_memory())
        self._pinned[0].copy_(ids, non_blocking=True)
        self._pinned[1].copy_(scores, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        out_ids, out_sc = (self._pinned[0].clone(), self._pinned[1].clone()) if copy else self._pinned
        return out_ids, r._finish_scores(out_sc, self._lq(q))

    # ---- hybrid pipeline (local_rag_complete.py:894-935) over the sharded corpus ---------------------
    def retrieve_batch(self, query_embeddings: torch.Tensor, bm25_ids: torch.Tensor, top_k_final: Optional[int] = None
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
        """ColBERT top-k over ALL shards -> RRF with the given BM25 lists (global ids) -> each rank scores the
        candidates it owns -> merged rerank.  Same result on every rank, bit-equal to HybridRetriever.retrieve_batch
        over the unsharded store.  bm25_ids: int [Bq, n_bm25] ranked GLOBAL doc ids (negative = absent)."""
        r = self.retriever
        cfg = r.config
        k_final = int(cfg.final_top_k if top_k_final is None else top_k_final)
        n_cand = int(_knob(cfg, "rerank_candidates"))
        colbert_k = min(int(cfg.colbert_top_k), self.n_docs_global())
        q = r._prep_queries(query_embeddings)
        a = bm25_ids.to(r.device, torch.int32).contiguous()
        s = r.store
        if self.transport != "torch" and not r._literal() and k_final <= n_cand:
            self._join_async()
            ids, scores = _lib.sharded_hybrid_retrieve(
                self.comm, s.tokens, s.offsets, self.n_docs_global(), q, a, colbert_k=colbert_k, rrf_k=_knob(cfg, "rrf_k"),
                n_candidates=n_cand, final_k=k_final, id_base=s.doc_id_base, path=self._path(),
                transport=_TRANSPORTS[self.transport], workspace=self._ws)
            return ids, r._finish_scores(scores, q.shape[1])
        # the same stages with torch.distributed collectives (cross-check; also score_mode="reference_literal")
        col_ids, _ = _lib.keys_unpack(self.search_keys(q, colbert_k))                       # global ColBERT list
        fused, _, _ = _lib.rrf_fuse(a, col_ids.contiguous(), _knob(cfg, "rrf_k"), n_cand)   # global ids, every rank
        local = fused - s.doc_id_base
        owned = (fused >= 0) & (local >= 0) & (local < s.n_docs)
        local = torch.where(owned, local, torch.full_like(local, -1)).contiguous()
        if r._literal():
            full = r._score_store(s, q)
            cs = torch.gather(full, 1, local.clamp_min(0).long())
        else:
            cs = _lib.maxsim_scores_ids(s.tokens, s.offsets, local, q, path=self._path())
        pos = torch.arange(n_cand, device=r.device, dtype=torch.int32).unsqueeze(0).expand_as(fused).contiguous()
        nobody = (fused < 0) | (fused >= self.n_docs_global())
        mine = owned | (nobody & (dist.get_rank(self.group) == 0))
        cs = torch.where(owned, cs, torch.full_like(cs, float("-inf"))).contiguous()
        keys = _lib.topk(cs, n_cand, ids=pos)                                               # sorted (score, pos) keys
        # drop the keys of candidates this rank does not answer for (their pos would collide with the owner's key)
        kpos, _ = _lib.keys_unpack(keys)
        keep = torch.gather(mine, 1, kpos.clamp_min(0).long())
        keys = torch.where(keep, keys, torch.zeros_like(keys)).contiguous()
        merged = _lib.topk_merge(all_gather_keys(keys, n_cand, self.group), k_final)
        mpos, scores = _lib.keys_unpack(merged)
        ids = torch.where(mpos >= 0, torch.gather(fused, 1, mpos.clamp_min(0).long()), mpos)
        if r._literal():
            return ids, scores
        return ids, r._finish_scores(scores, q.shape[1])

    def close(self) -> None:
        if self.comm is not None:
            self.comm.close()
            self.comm = None
