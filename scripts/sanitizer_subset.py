#!/usr/bin/env python
"""A <= 60 s subset of the parity suite for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
one small case per kernel organisation — single-CTA tensor-core kernel with 1, 2 and 4 queries, the doc-major single-query kernel,
the single-CTA batched kernel, CTA pairs (cta_group::2) with and without an odd query group, candidate (rerank)
mode, queries longer than 32 tokens (slot partials in the caller's workspace), the CUDA-core kernel, mean-pool
cosine, radix top-k (one and several chunks), merge, RRF and the fused search / rerank / hybrid calls.
Every result is checked against the CPU oracle, so a sanitizer-clean run is also a correct run.

    compute-sanitizer --tool memcheck python scripts/sanitizer_subset.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import hybrid_rag_colbertv2_b200 as hrc  # noqa: E402
from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from oracle import maxsim_oracle as o  # noqa: E402

dev = torch.device("cuda:0")
L.set_watchdog_ms(0)            # kernels run 10-100x slower under the sanitizer: no wall-clock watchdog


def case(seed, lens, bq, lq):
    g = torch.Generator().manual_seed(seed)
    lens = torch.as_tensor(lens, dtype=torch.int64)
    off = torch.zeros(lens.numel() + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(lens, 0)
    tok = torch.nn.functional.normalize(torch.randn((int(off[-1]), 128), generator=g), dim=-1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn((bq, lq, 128), generator=g), dim=-1).to(torch.bfloat16)
    return q, tok, off


def close(got, exp, what):
    got = got.float().cpu()
    fin = torch.isfinite(exp)
    assert torch.equal(torch.isfinite(got), fin), what
    err = float((got[fin] - exp[fin]).abs().max() / exp[fin].abs().max().clamp_min(1e-6)) if fin.any() else 0.0
    assert err < 1e-4, f"{what}: {err}"
    print(f"ok {what}: rel err {err:.1e}", flush=True)


rng = np.random.default_rng(3)
lens = np.concatenate([rng.integers(1, 200, 300), [0, 1, 31, 32, 33, 127, 128, 129, 400]]).tolist()   # ~31k tokens: every SM
for bq, lq, path, name in [(1, 32, L.PATH_TC, "tc 1 query"), (2, 32, L.PATH_TC, "tc 2 queries"), (4, 17, L.PATH_TC, "tc 4 queries lq=17"),
                           (1, 32, L.PATH_TC_DM, "doc-major 1 query"), (1, 9, L.PATH_TC_DM, "doc-major lq=9"),
                           (7, 32, L.PATH_TC, "single-CTA batched, 7 queries"), (16, 32, L.PATH_TC, "CTA pairs, 16 queries"),
                           (24, 32, L.PATH_TC, "CTA pairs + odd group, 24 queries"), (2, 70, L.PATH_TC, "lq=70 (3 slots)"),
                           (9, 40, L.PATH_TC, "lq=40 x 9 queries (pairs of slots)"), (2, 32, L.PATH_SIMT, "simt")]:
    q, tok, off = case(11, lens, bq, lq)
    exp = o.maxsim_scores(q.float(), tok.float(), off)
    got = L.maxsim_scores(tok.to(dev), off.to(dev), q.to(dev), path=path)
    torch.cuda.synchronize()
    close(got, exp, name)

q, tok, off = case(12, lens, 3, 32)
tok_d, off_d, q_d = tok.to(dev), off.to(dev), q.to(dev)
full = o.maxsim_scores(q.float(), tok.float(), off)
n = len(lens)
cand = torch.from_numpy(rng.integers(0, n, (3, 50)).astype(np.int32))
cand[1, 3] = -1
cand[2, 4] = n
expc = torch.gather(full, 1, cand.clamp(0, n - 1).long())
expc[1, 3] = float("-inf")
expc[2, 4] = float("-inf")
for path, name in ((L.PATH_TC, "candidates tc"), (L.PATH_SIMT, "candidates simt")):
    close(L.maxsim_scores_ids(tok_d, off_d, cand.to(dev), q_d, path=path), expc, name)
q70 = case(13, lens, 2, 70)[0]
exp70 = torch.gather(o.maxsim_scores(q70.float(), tok.float(), off), 1, cand[:2].clamp(0, n - 1).long())
exp70[1, 3] = float("-inf")
close(L.maxsim_scores_ids(tok_d, off_d, cand[:2].contiguous().to(dev), q70.to(dev)), exp70, "candidates lq=70")

lit = L.meanpool_cosine_scores(tok_d, off_d, q_d).cpu()
for d in (1, 5, 301):
    e = o.literal_reference(q.float(), tok[int(off[d]):int(off[d + 1])].float())
    assert float((lit[:, d] - e).abs().max()) < 1e-4
print("ok meanpool cosine", flush=True)

for nn, k, rows in ((50, 10, 2), (20_000, 100, 2), (8192, 64, 1)):
    s = torch.from_numpy(rng.standard_normal((rows, nn)).astype(np.float32))
    s[:, : nn // 3] = torch.round(s[:, : nn // 3] * 4) / 4
    keys = L.topk(s.to(dev), k, id_base=3)
    ref = o.merge_keys(o.make_keys(s.numpy(), np.broadcast_to(np.arange(nn) + 3, s.shape)), k)
    assert (keys.cpu().numpy().view(np.uint64) == ref).all()
print("ok radix top-k", flush=True)
allk = np.concatenate([o.make_keys(rng.standard_normal((2, 100)).astype(np.float32), np.arange(100)[None] + 1000 * r) for r in range(8)], 1)
got = L.topk_merge(torch.from_numpy(allk.view(np.int64).copy()).to(dev), 100)
assert (got.cpu().numpy().view(np.uint64) == o.merge_keys(allk, 100)).all()
print("ok merge", flush=True)
a = rng.integers(0, 300, (4, 100)).astype(np.int32)
b = rng.integers(0, 300, (4, 100)).astype(np.int32)
ids, sc, cnt = L.rrf_fuse(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev), 60, 200)
for r in range(4):
    ri, rs = o.rrf_ids(a[r].tolist(), b[r].tolist(), 60)
    assert ids[r, :len(ri)].tolist() == ri and sc[r, :len(ri)].tolist() == rs
print("ok rrf", flush=True)

# the fused calls, through the reference-shaped classes
r = hrc.JinaColBERTRetriever(hrc.RAGConfig(colbert_top_k=40, rerank_candidates=20, final_top_k=5), encoder=hrc.SyntheticEncoder())
r.store = hrc.PackedStore.from_packed(tok, off, device=dev, allow_empty=True)
ids, sc = r.search_embeddings(q, 40)
for i in range(3):
    assert o.check_ranking(ids[i].tolist(), sc[i].tolist(), full[i], 40, 1e-3) is None
hi, hs = r.search_host(q.float(), 40)
assert torch.equal(hi, ids.cpu()) and torch.equal(hs, sc.cpu())
idx = hrc.DualIndexer(r.config, encoder=hrc.SyntheticEncoder())
idx.colbert_retriever = r
h = hrc.HybridRetriever(r.config, idx, None, verbose=False)
bm = torch.from_numpy(rng.integers(0, n, (3, 30)).astype(np.int32)).to(dev)
f_ids, f_sc = h.retrieve_batch(q, bm)
s_ids, s_sc = h._retrieve_batch_staged(q, bm, 5)
assert torch.equal(f_ids, s_ids) and torch.equal(f_sc, s_sc)
torch.cuda.synchronize()
print("ok fused search / search_host / hybrid retrieve", flush=True)
print("sanitizer_subset ok")
