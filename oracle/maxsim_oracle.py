"""CPU oracle for the MaxSim hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker (or the timed CPU baseline), never as a product path.  The product
(hybrid-rag-colbertv2_b200/) never imports it and has no CPU fallback.

What is restated, and how it is pinned
  * literal_reference(q, D)      — line-by-line restatement of what the reference's `_maxsim_score`
    really computes (mean-pool cosine), local_rag_complete.py:813-831.  PINNED: bit-equal to the
    imported reference function on the committed fixtures (tests/golden/, made by make_golden.py).
  * rrf_reference(a, b, k)       — restatement of `_reciprocal_rank_fusion`, :960-978.  PINNED the same way
    (ids, fp64 scores and tie order, including the known answer in SURVEY.md §A.3).
  * search_reference / rerank_reference — result-dict shapes and k clamping of :755-800.  PINNED on fixtures.
  * maxsim_scores / maxsim_dense — TRUE MaxSim as the reference's docstring (:807-812) and
    BASELINE.json's north_star define it: per query token, max over the document's tokens, summed over
    query tokens, fp32 arithmetic on the same bf16-rounded inputs the kernels see.
    PARITY ONLY PARTLY PINNED BY THE REFERENCE: the reference's own function does not compute MaxSim in
    general and it ships no tests or golden vectors, so no reference output exists for
    MaxSim on general inputs.  What IS pinned: on the degenerate shapes where its mean-pool cosine equals
    MaxSim (identical query tokens, identical document tokens, exactly unit-norm dyadic rows) the
    unmodified reference's outputs (tests/golden/maxsim_pin.npz, made by make_golden.py) are reproduced
    bit for bit — the dot product, the max over document tokens and the sum over query tokens.  General
    inputs are anchored on float64 known answers computed with plain numpy loops
    (tests/golden/maxsim_kat_f64.npz, make_kat.py), an independent pure-loop implementation
    (maxsim_naive), the plain-C restatement (maxsim_oracle.c) and the dense einsum form.
All arithmetic the reference delegates to torch (unpinned `torch>=2.0.0`, requirements.txt:8) is done
here with the installed torch 2.11 CPU kernels, TF32 off.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

DIM = 128


# ----------------------------------------------------------------------------------------------
# what the reference literally computes                                  local_rag_complete.py:813-831
# ----------------------------------------------------------------------------------------------
def literal_reference(query_embedding: torch.Tensor, doc_embeddings: torch.Tensor) -> torch.Tensor:
    if query_embedding.dim() == 2:                       # :814-815
        query_embedding = query_embedding.unsqueeze(0)
    if doc_embeddings.dim() == 2:                        # :816-817  (a 2-D corpus becomes ONE document)
        doc_embeddings = doc_embeddings.unsqueeze(0)
    query_vec = query_embedding.mean(dim=1)              # :821
    doc_vec = doc_embeddings.mean(dim=1)                 # :822
    scores = torch.nn.functional.cosine_similarity(      # :825-829
        query_vec.unsqueeze(1), doc_vec.unsqueeze(0), dim=2)
    return scores.squeeze()                              # :831


# ----------------------------------------------------------------------------------------------
# true MaxSim (the parity target)                 docstring :807-812, shapes :813-817, north_star (sum)
# ----------------------------------------------------------------------------------------------
def round_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 view of the bf16-rounded values — the inputs the kernels actually see."""
    return x.to(torch.bfloat16).to(torch.float32)


def maxsim_dense(query_embedding: torch.Tensor, doc_embeddings: torch.Tensor,
                 lengths: Optional[Sequence[int]] = None) -> torch.Tensor:
    """einsum -> max over (masked) doc tokens -> sum over query tokens; shapes/squeeze as the reference."""
    q = query_embedding.float()
    d = doc_embeddings.float()
    if q.dim() == 2:
        q = q.unsqueeze(0)
    if d.dim() == 2:
        d = d.unsqueeze(0)
    sim = torch.einsum('bqd,nld->bnql', q, d)
    if lengths is not None:
        lens = torch.as_tensor(lengths, dtype=torch.int64)
        pad = torch.arange(d.shape[1]).unsqueeze(0) >= lens.unsqueeze(1)           # [N, Ld]
        sim = sim.masked_fill(pad[None, :, None, :], float('-inf'))
    return sim.max(dim=-1).values.sum(dim=-1).squeeze()


def maxsim_scores(queries: torch.Tensor, tokens: torch.Tensor, offsets: torch.Tensor,
                  doc_chunk_tokens: int = 1 << 18) -> torch.Tensor:
    """Packed form: fp32 [Bq, N]; score[b, i] = sum_q max_{t in doc i} <Q[b, q], tokens[t]>.

    Inputs are used as given (callers pass bf16-rounded values); an empty document scores -inf.
    """
    q = queries.float()
    if q.dim() == 2:
        q = q.unsqueeze(0)
    bq, lq, _ = q.shape
    tok = tokens.float()
    off = offsets.to(torch.int64).cpu()
    n = off.numel() - 1
    out = torch.empty((bq, n), dtype=torch.float32)
    qm = q.reshape(bq * lq, -1)
    d0 = 0
    while d0 < n:
        # take whole documents up to ~doc_chunk_tokens tokens
        limit = int(off[d0]) + doc_chunk_tokens
        d1 = int(torch.searchsorted(off, torch.tensor(limit), right=True)) - 1
        d1 = min(max(d1, d0 + 1), n)
        t0, t1 = int(off[d0]), int(off[d1])
        lens = off[d0 + 1:d1 + 1] - off[d0:d1]
        sim = tok[t0:t1] @ qm.T                                              # [tokens, Bq*Lq]
        if t1 > t0:
            seg = torch.segment_reduce(sim, 'max', lengths=lens, axis=0, unsafe=True)   # [docs, Bq*Lq]
        else:
            seg = torch.full((d1 - d0, bq * lq), float('-inf'))
        seg = torch.where((lens == 0).unsqueeze(1), torch.full_like(seg, float('-inf')), seg)
        out[:, d0:d1] = seg.reshape(d1 - d0, bq, lq).sum(-1).T
        d0 = d1
    return out


def maxsim_naive(queries: np.ndarray, tokens: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """Independent pure-loop MaxSim (fp32 accumulation in index order) for small cases."""
    q = np.asarray(queries, dtype=np.float32)
    if q.ndim == 2:
        q = q[None]
    tok = np.asarray(tokens, dtype=np.float32)
    off = np.asarray(offsets, dtype=np.int64)
    out = np.empty((q.shape[0], off.shape[0] - 1), dtype=np.float32)
    for b in range(q.shape[0]):
        for i in range(off.shape[0] - 1):
            total = np.float32(0.0)
            for qi in range(q.shape[1]):
                best = np.float32(-np.inf)
                for t in range(off[i], off[i + 1]):
                    dot = np.float32(0.0)
                    for k in range(tok.shape[1]):
                        dot = np.float32(dot + q[b, qi, k] * tok[t, k])
                    if dot > best:
                        best = dot
                total = np.float32(total + best)
            out[b, i] = total
    return out


# ----------------------------------------------------------------------------------------------
# top-k / sort                                                             :767, :789-792
# ----------------------------------------------------------------------------------------------
def topk_reference(scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """torch.topk(scores, k=min(k, len(scores))) as at :767 -> (indices, values)."""
    r = torch.topk(scores, k=min(k, scores.shape[-1]))
    return r.indices, r.values


def topk_deterministic(scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """The kernels' documented order: score descending, then id ascending (one of torch's allowed outcomes)."""
    s = scores.float()
    order = torch.argsort(s, dim=-1, descending=True, stable=True)
    k = min(k, s.shape[-1])
    idx = order[..., :k]
    return idx, torch.gather(s, -1, idx)


def float_to_orderable(s: np.ndarray) -> np.ndarray:
    s = np.asarray(s, dtype=np.float32).copy()
    s[np.isnan(s)] = -np.inf
    u = s.view(np.uint32)
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def make_keys(scores: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """uint64 keys as include/hrc.h defines them: orderable(score) << 32 | ~id."""
    o = float_to_orderable(scores).astype(np.uint64)
    return (o << np.uint64(32)) | (~np.asarray(ids, dtype=np.uint32)).astype(np.uint64)


def unpack_keys(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    keys = np.asarray(keys).astype(np.uint64)
    ids = (~(keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)).astype(np.int32)
    o = (keys >> np.uint64(32)).astype(np.uint32)
    u = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o).astype(np.uint32)
    scores = u.view(np.float32).copy()
    ids = np.where(keys == 0, -1, ids)
    scores = np.where(keys == 0, -np.inf, scores).astype(np.float32)
    return ids, scores


def merge_keys(keys: np.ndarray, k: int) -> np.ndarray:
    """Row-wise k largest keys, descending; fewer than k non-empty keys are padded with 0."""
    keys = np.asarray(keys).astype(np.uint64)
    srt = np.sort(keys, axis=-1)[..., ::-1]
    out = np.zeros(keys.shape[:-1] + (k,), dtype=np.uint64)
    m = min(k, keys.shape[-1])
    out[..., :m] = srt[..., :m]
    return out


# ----------------------------------------------------------------------------------------------
# RRF                                                                       :960-978, slice :916
# ----------------------------------------------------------------------------------------------
def rrf_reference(bm25_results: List[Dict], colbert_results: List[Dict], k: int = 60) -> List[Dict]:
    scores = {}
    for rank, result in enumerate(bm25_results, 1):                         # :969-971
        chunk_id = result['chunk_id']
        scores[chunk_id] = scores.get(chunk_id, 0) + (1 / (k + rank))
    for rank, result in enumerate(colbert_results, 1):                      # :973-975
        chunk_id = result['chunk_id']
        scores[chunk_id] = scores.get(chunk_id, 0) + (1 / (k + rank))
    sorted_results = sorted(scores.items(), key=lambda x: x[1], reverse=True)   # :977 (stable)
    return [{'chunk_id': cid, 'rrf_score': score} for cid, score in sorted_results]


def rrf_ids(a_ids: Sequence[int], b_ids: Sequence[int], k: int = 60) -> Tuple[List[int], List[float]]:
    fused = rrf_reference([{'chunk_id': int(i)} for i in a_ids if i >= 0],
                          [{'chunk_id': int(i)} for i in b_ids if i >= 0], k)
    return [r['chunk_id'] for r in fused], [r['rrf_score'] for r in fused]


# ----------------------------------------------------------------------------------------------
# search / rerank result shapes                                            :755-800
# ----------------------------------------------------------------------------------------------
def search_reference(scores: torch.Tensor, corpus: Optional[List[str]], k: int) -> List[Dict]:
    idx, _ = topk_reference(scores, k)
    return [{'document_id': int(i), 'score': float(scores[i]), 'text': corpus[i] if corpus else None} for i in idx]


def rerank_reference(scores: torch.Tensor, documents: List[str], k: int) -> List[Dict]:
    order = torch.argsort(scores, descending=True)
    return [{'result_index': int(i), 'score': float(scores[i]), 'rank': r + 1, 'text': documents[i]}
            for r, i in enumerate(order[:k])]


# ----------------------------------------------------------------------------------------------
# gap-aware ranking comparison ("ids and order bit-exact wherever gaps exceed tolerance")
# ----------------------------------------------------------------------------------------------
def check_ranking(test_ids: Sequence[int], test_scores: Sequence[float], oracle_scores: torch.Tensor, k: int,
                  rtol: float = 1e-3) -> Optional[str]:
    """None if (ids, scores) is a valid top-k of `oracle_scores` up to `rtol`; else a description.

    Rules: scores within rtol (relative) of the oracle's score of the same id; position i must hold the
    oracle's i-th id unless the two ids' oracle scores differ by less than the tolerance; every returned
    id must belong to the oracle top-k unless it ties (within tolerance) with the k-th oracle score.
    """
    s = oracle_scores.float()
    n = s.numel()
    k = min(k, n)
    if len(test_ids) != k:
        return f"length {len(test_ids)} != {k}"
    if len(set(int(i) for i in test_ids)) != k:
        return "duplicate ids"
    o_idx, o_val = topk_deterministic(s, k)
    scale = max(float(s[torch.isfinite(s)].abs().max()) if torch.isfinite(s).any() else 1.0, 1e-6)
    tol = rtol * scale
    kth = float(o_val[-1])
    prev = float('inf')
    for pos, (tid, ts) in enumerate(zip(test_ids, test_scores)):
        tid = int(tid)
        if not (0 <= tid < n):
            return f"pos {pos}: id {tid} out of range"
        ref = float(s[tid])
        if abs(ts - ref) > tol and not (ts == ref):
            return f"pos {pos}: score {ts} vs oracle {ref} for id {tid} (tol {tol})"
        if ts > prev + 1e-12:
            return f"pos {pos}: scores not descending"
        prev = ts
        if tid != int(o_idx[pos]) and abs(ref - float(o_val[pos])) > 2 * tol:
            return f"pos {pos}: id {tid} (oracle score {ref}) vs oracle id {int(o_idx[pos])} ({float(o_val[pos])})"
        if ref < kth - 2 * tol:
            return f"pos {pos}: id {tid} score {ref} is below the k-th oracle score {kth}"
    return None
