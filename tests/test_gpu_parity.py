"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the
same bf16-rounded inputs.  Tolerance (north_star): fp32-accumulated scores within 1e-3 relative;
returned ids and their order exact wherever oracle score gaps exceed that tolerance; integer work
(keys, RRF ids, fp64 RRF scores) bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as o

pytestmark = pytest.mark.gpu

RTOL = 1e-3          # the API contract (north_star): fp32-accumulated scores within 1e-3 relative
# What the kernels are actually held to: <= 10x the largest error observed on the B200 (every check below records its
# error; tests/conftest.py writes the maxima to gpurun_out/parity_errors.json; profiles/r02_summary.md quotes them).
# A dropped or doubled document token moves a score by ~1e-3..1e-1 relative, far outside this.
TIGHT = 4e-6
OBSERVED = {}        # what -> largest relative error seen (dumped at session end)


def _lib():
    from hybrid_rag_colbertv2_b200 import _lib
    return _lib


def _case(seed, n_docs, min_len, max_len, bq, lq, lens=None):
    g = torch.Generator().manual_seed(seed)
    if lens is None:
        lens = torch.randint(min_len, max_len + 1, (n_docs,), generator=g)
    else:
        lens = torch.as_tensor(lens, dtype=torch.int64)
    off = torch.zeros(lens.numel() + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(lens, 0)
    tok = torch.nn.functional.normalize(torch.randn((int(off[-1]), 128), generator=g), dim=-1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn((bq, lq, 128), generator=g), dim=-1).to(torch.bfloat16)
    return q, tok, off


def _assert_scores(got: torch.Tensor, exp: torch.Tensor, what="", tol=TIGHT, bucket="maxsim"):
    got = got.float().cpu()
    fin = torch.isfinite(exp)
    assert torch.equal(torch.isfinite(got), fin), f"{what}: finite pattern differs"
    assert torch.equal(got[~fin], exp[~fin]), f"{what}: non-finite values differ"
    scale = exp[fin].abs().max().clamp_min(1e-6) if fin.any() else 1.0
    err = ((got[fin] - exp[fin]).abs().max() / scale).item() if fin.any() else 0.0
    OBSERVED[bucket] = max(OBSERVED.get(bucket, 0.0), err)
    assert err <= tol, f"{what}: max relative error {err:.3e} > {tol}"


@pytest.fixture(scope="module", autouse=True)
def _dump_observed_errors():
    yield
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "parity_errors.json"), "w") as f:
        json.dump({"tight_tolerance": TIGHT, "api_tolerance": RTOL, "max_relative_error_observed": OBSERVED}, f, indent=1)


def _path(L, name):
    return {"tc": L.PATH_TC, "simt": L.PATH_SIMT, "tc_dm": L.PATH_TC_DM, "auto": L.PATH_AUTO}[name]


SHAPES = [
    # (n_docs, min_len, max_len, bq, lq)           what it exercises
    (50, 32, 512, 1, 32),      # C1 rerank shape
    (300, 128, 128, 1, 32),    # C2 shape, tile-aligned documents
    (257, 1, 1, 1, 32),        # one-token documents: a boundary at every column
    (1, 700, 700, 1, 32),      # N == 1 (the reference crashes here), document spanning 6 tiles
    (40, 1, 300, 3, 32),       # ragged, several queries in one M tile
    (64, 32, 512, 9, 32),      # > 8 queries: two query groups on the MT=2 kernel, the second nearly empty
    (33, 127, 129, 5, 32),     # documents straddling tile boundaries by one token, MT=2
    (120, 1, 200, 21, 32),     # 21 queries: three 8-query groups on the batched kernel, the last partial
    (20, 5, 90, 2, 17),        # lq < 32: query rows zero-filled by TMA
    (10, 3, 40, 1, 1),         # single query token
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("path", ["tc", "simt", "tc_dm", "auto"])
def test_maxsim_scores_match_oracle(cuda_dev, shape, path):
    """Every scoring path: tcgen05 query-major (HRC_PATH_TC), CUDA cores, the doc-major kernel for one query
    (HRC_PATH_TC_DM; with more queries it is the query-major kernel) and what HRC_PATH_AUTO picks."""
    L = _lib()
    q, tok, off = _case(hash(shape) % 10_000, *shape)
    exp = o.maxsim_scores(q.float(), tok.float(), off)
    got = L.maxsim_scores(tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev), path=_path(L, path))
    torch.cuda.synchronize()
    _assert_scores(got, exp, f"{path} {shape}", bucket=path)


@pytest.mark.parametrize("shape", [(64, 32, 512, 9, 32), (33, 127, 129, 5, 32), (120, 1, 200, 21, 32), (300, 1, 40, 16, 20),
                                   (3000, 16, 200, 40, 32), (1, 700, 700, 17, 32), (500, 1, 300, 2, 32), (77, 30, 34, 1, 32),
                                   (5000, 1, 70, 1, 32), (900, 100, 600, 1, 20), (3, 1, 2, 1, 32), (4, 3000, 9000, 1, 32),
                                   (2, 40_000, 60_000, 1, 7)])
@pytest.mark.parametrize("path", ["tc", "tc_dm"])
def test_batched_and_single_query_kernels_more_shapes(cuda_dev, shape, path):
    """Batched (MT=2) kernels — CTA pairs (cta_group::2, from two query groups up, with an odd last group on the
    single-CTA kernel) and single CTAs: 40 queries over many segments, 17 queries on one 6-tile document, short
    documents, partial query groups — and the one-query shapes (very long documents, thousands of short ones, fewer
    documents than token streams) on both single-query kernels (query-major / doc-major)."""
    L = _lib()
    q, tok, off = _case(78, *shape)
    exp = o.maxsim_scores(q.float(), tok.float(), off)
    got = L.maxsim_scores(tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev), path=_path(L, path))
    torch.cuda.synchronize()
    _assert_scores(got, exp, f"batched {path} {shape}", bucket=path)


def test_reference_literal_path_matches_reference_outputs(cuda_dev, golden_dir):
    """PINNED parity: the mean-pool-cosine kernel (what the reference's `_maxsim_score` literally computes,
    local_rag_complete.py:821-829) and the retriever in score_mode="reference_literal" against vectors the
    UNMODIFIED reference produced (tests/golden/literal_bf16.npz, inputs exactly representable in bf16)."""
    import hybrid_rag_colbertv2_b200 as hrc
    L = _lib()
    z = np.load(os.path.join(golden_dir, "literal_bf16.npz"))
    q, qb, D = (torch.from_numpy(z[k]) for k in ("q", "qb", "D"))
    n, ld, _ = D.shape
    tok = D.reshape(n * ld, 128).to(torch.bfloat16).to(cuda_dev)
    off = torch.arange(0, (n + 1) * ld, ld, dtype=torch.int64, device=cuda_dev)
    got = L.meanpool_cosine_scores(tok, off, q.unsqueeze(0).to(torch.bfloat16).to(cuda_dev))
    _assert_scores(got[0], torch.from_numpy(z["out_q_D"]), "literal q", tol=RTOL, bucket="literal")
    gotb = L.meanpool_cosine_scores(tok, off, qb.to(torch.bfloat16).to(cuda_dev))
    _assert_scores(gotb, torch.from_numpy(z["out_qb_D"]), "literal qb", tol=RTOL, bucket="literal")

    class FixedEncoder:                       # the encoder make_golden.py gave the reference
        def encode(self, x, **kw):
            return q if isinstance(x, str) else torch.stack([D[int(t[1:])] for t in x])

    r = hrc.JinaColBERTRetriever(hrc.RAGConfig(score_mode="reference_literal"), encoder=FixedEncoder())
    r.index_embeddings(D, corpus=[f"d{i}" for i in range(n)])
    res = r.search(query="query", k=10)
    exp_scores = torch.from_numpy(z["out_q_D"])
    assert o.check_ranking([x["document_id"] for x in res], [x["score"] for x in res], exp_scores, 10, RTOL) is None
    assert [x["document_id"] for x in res][:5] == z["search_ids"].tolist()[:5]      # gaps there are > tolerance
    np.testing.assert_allclose([x["score"] for x in res], z["search_scores"], rtol=0, atol=RTOL * 0.22)
    cand = z["rerank_cand"].tolist()
    rr = r.rerank(query="query", documents=[f"d{i}" for i in cand], k=5)
    assert [x["result_index"] for x in rr] == z["rerank_index"].tolist()
    np.testing.assert_allclose([x["score"] for x in rr], z["rerank_scores"], rtol=0, atol=RTOL * 0.22)
    assert [x["rank"] for x in rr] == [1, 2, 3, 4, 5]
    # _maxsim_score(mode=...) keeps the reference's shapes (:813-817, :831)
    s = r._maxsim_score(q, D)
    assert s.shape == (n,)
    _assert_scores(s, exp_scores, "_maxsim_score literal", tol=RTOL, bucket="literal")
    assert r._maxsim_score(q, D, mode="maxsim").shape == (n,) and float(r._maxsim_score(q, D, mode="maxsim").min()) > 2.0
    # ragged store + empty document: NaN like torch's mean over an empty axis
    q3, tok3, off3 = _case(5, 0, 0, 0, 2, 32, lens=[3, 0, 130, 1, 77])
    g3 = L.meanpool_cosine_scores(tok3.to(cuda_dev), off3.to(cuda_dev), q3.to(cuda_dev)).cpu()
    assert bool(torch.isnan(g3[:, 1]).all())
    for d in (0, 2, 3, 4):
        e = o.literal_reference(q3.float(), tok3[int(off3[d]):int(off3[d + 1])].float())
        assert float((g3[:, d] - e).abs().max()) < 1e-4


def test_fused_search_and_rerank_calls(cuda_dev):
    """hrc_search / hrc_rerank (one C call each) equal the staged calls bit for bit."""
    L = _lib()
    q, tok, off = _case(31, 20_000, 8, 64, 3, 32)
    tok_d, off_d, q_d = tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev)
    keys, ids, sc = L.search(tok_d, off_d, q_d, 100, id_base=5)
    staged = L.topk(L.maxsim_scores(tok_d, off_d, q_d), 100, id_base=5)
    assert torch.equal(keys, staged)
    si, ss = L.keys_unpack(staged)
    assert torch.equal(ids, si) and torch.equal(sc, ss)
    g = torch.Generator().manual_seed(2)
    cand = torch.randint(0, 20_000, (3, 50), generator=g, dtype=torch.int32).to(cuda_dev)
    cand[1, 4] = -1
    pos, dids, rs, cs = L.rerank(tok_d, off_d, cand, q_d, 10, want_cand_scores=True)
    assert L.rerank(tok_d, off_d, cand, q_d, 10)[3] is None
    cs2 = L.maxsim_scores_ids(tok_d, off_d, cand, q_d)
    assert torch.equal(cs, cs2)
    p2, s2 = L.keys_unpack(L.topk(cs2, 10))
    assert torch.equal(pos, p2) and torch.equal(rs, s2)
    assert torch.equal(dids, torch.gather(cand, 1, pos.long()))
    with pytest.raises(Exception):
        L.search(tok_d, off_d, q_d, 20_001)


def test_search_host_equals_device_search(cuda_dev):
    """hrc_search_host (host fp32 queries in, host ids/scores out, one C call) == the device-tensor API."""
    import hybrid_rag_colbertv2_b200 as hrc
    q, tok, off = _case(41, 5000, 8, 90, 3, 32)
    r = hrc.JinaColBERTRetriever(hrc.RAGConfig())
    r.index_embeddings(tok, off, packed=True)
    qf = q.float()                                            # bf16-representable fp32, as an encoder would hand over
    ids_d, sc_d = r.search_embeddings(qf, 50)
    for host_q in (qf, qf.pin_memory()):
        ids_h, sc_h = r.search_host(host_q, 50)
        assert not ids_h.is_cuda and ids_h.dtype == torch.int32 and ids_h.shape == (3, 50)
        assert torch.equal(ids_h, ids_d.cpu()) and torch.equal(sc_h, sc_d.cpu())
    one_i, one_s = r.search_host(qf[1], 7)                    # [Lq, 128] -> one query
    assert torch.equal(one_i[0], ids_d[1, :7].cpu()) and torch.equal(one_s[0], sc_d[1, :7].cpu())
    small = hrc.JinaColBERTRetriever(hrc.RAGConfig())
    small.index_embeddings(tok[: int(off[30])], off[:31], packed=True)
    assert small.search_host(qf, 1000)[0].shape == (3, 30)    # k clamps to N (:767)


@pytest.mark.parametrize("shape", [(25, 1, 70, 2, 77), (300, 20, 300, 1, 64), (40, 1, 200, 5, 33), (64, 32, 512, 9, 96),
                                   (10, 3, 40, 1, 256)])
def test_long_queries_on_the_tensor_core_path(cuda_dev, shape):
    """lq > 32: the query is scored as ceil(lq / 32) slots of <= 32 tokens (zero-filled rows add 0) whose partial
    scores are summed in slot order — single-query, few-query and CTA-pair kernels, and the candidate entry point."""
    L = _lib()
    q, tok, off = _case(11, *shape)
    exp = o.maxsim_scores(q.float(), tok.float(), off)
    tok_d, off_d, q_d = tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev)
    _assert_scores(L.maxsim_scores(tok_d, off_d, q_d, path=L.PATH_TC), exp, f"tc lq={shape[4]}")
    _assert_scores(L.maxsim_scores(tok_d, off_d, q_d), exp, f"auto lq={shape[4]}")
    n = shape[0]
    g = torch.Generator().manual_seed(4)
    cand = torch.randint(0, n, (shape[3], 7), generator=g, dtype=torch.int32)
    cand[0, 3] = -1
    expc = torch.gather(exp, 1, cand.clamp(0, n - 1).to(torch.int64))
    expc[0, 3] = float("-inf")
    _assert_scores(L.maxsim_scores_ids(tok_d, off_d, cand.to(cuda_dev), q_d, path=L.PATH_TC), expc, "candidates")


def test_very_long_queries_take_the_simt_path(cuda_dev):
    L = _lib()
    q, tok, off = _case(11, 25, 1, 70, 2, 300)
    exp = o.maxsim_scores(q.float(), tok.float(), off)
    _assert_scores(L.maxsim_scores(tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev)), exp, "auto lq=300")
    with pytest.raises(Exception):
        L.maxsim_scores(tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev), path=L.PATH_TC)


def test_empty_documents_score_minus_inf(cuda_dev):
    L = _lib()
    q, tok, off = _case(12, 0, 0, 0, 2, 32, lens=[0, 5, 0, 0, 130, 1, 0])
    exp = o.maxsim_scores(q.float(), tok.float(), off)
    for path in (L.PATH_TC, L.PATH_SIMT):
        _assert_scores(L.maxsim_scores(tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev), path=path), exp, str(path))


def test_many_segments_cover_every_document_once(cuda_dev):
    """More tiles than SMs: every CTA owns a run of whole documents; none may be dropped or doubled."""
    L = _lib()
    q, tok, off = _case(13, 3000, 16, 200, 1, 32)
    exp = o.maxsim_scores(q.float(), tok.float(), off)
    out = torch.full((1, 3000), float("nan"), device=cuda_dev)
    L.maxsim_scores(tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev), path=L.PATH_TC, out=out)
    _assert_scores(out, exp, "3000 ragged docs")


def test_tc_and_simt_agree_on_candidates(cuda_dev):
    L = _lib()
    q, tok, off = _case(14, 400, 1, 512, 4, 32)
    g = torch.Generator().manual_seed(3)
    cand = torch.randint(0, 400, (4, 50), generator=g, dtype=torch.int32)
    cand[1, 7] = -1
    cand[2, 9] = 400                                      # out of range -> -inf
    full = o.maxsim_scores(q.float(), tok.float(), off)
    exp = torch.gather(full, 1, cand.clamp(0, 399).to(torch.int64))
    exp[1, 7] = float("-inf")
    exp[2, 9] = float("-inf")
    for path in (L.PATH_TC, L.PATH_SIMT):
        got = L.maxsim_scores_ids(tok.to(cuda_dev), off.to(cuda_dev), cand.to(cuda_dev), q.to(cuda_dev), path=path)
        _assert_scores(got, exp, f"candidates path {path}")


@pytest.mark.parametrize("n,k,rows", [(50, 10, 1), (50, 100, 2), (8192, 100, 3), (8193, 100, 1), (100_000, 100, 4),
                                      (300_000, 1000, 1), (70_000, 2048, 2), (5, 1, 1)])
def test_topk_is_exact(cuda_dev, n, k, rows):
    L = _lib()
    g = torch.Generator().manual_seed(n + k)
    s = torch.randn((rows, n), generator=g)
    s[:, : n // 3] = torch.round(s[:, : n // 3] * 4) / 4          # many exact ties
    if n > 10:
        s[0, 3] = float("nan")
        s[0, 4] = float("-inf")
        s[0, 5] = float("inf")
    keys = L.topk(s.to(cuda_dev), k, id_base=7)
    ids, sc = L.keys_unpack(keys)
    ref = o.merge_keys(o.make_keys(s.numpy(), np.broadcast_to(np.arange(n) + 7, s.shape)), k)
    assert (keys.cpu().numpy().view(np.uint64) == ref).all()
    ri, rs = o.unpack_keys(ref)
    assert (ids.cpu().numpy() == ri).all() and (sc.cpu().numpy() == rs).all()
    # against torch.topk: same values; ids may differ only inside ties
    kk = min(k, n)
    clean = torch.where(torch.isnan(s), torch.full_like(s, float("-inf")), s)
    tv = torch.topk(clean, kk).values
    assert torch.equal(sc.cpu()[:, :kk], tv)


@pytest.mark.parametrize("pattern", ["ascending", "descending", "constant", "random"])
def test_streaming_topk_adversarial_rows(cuda_dev, pattern):
    """The streaming top-k (k <= 128): rows in which EVERY score beats the running threshold (ascending: a compaction
    every 156 appends), none does, all tie (id order decides), lengths that are not multiples of the 32-lane / 256-score
    granules, many rows (several chunks per row and one chunk per row), explicit ids — bit-exact against the key sort."""
    L = _lib()
    rng = np.random.default_rng(5)
    for n, rows, k in ((100_003, 3, 100), (70_001, 300, 128), (5_000, 2, 100), (257, 5, 128), (31, 2, 31), (1_000_000, 1, 100)):
        if pattern == "ascending":
            s = np.tile(np.linspace(-3, 3, n, dtype=np.float32), (rows, 1))
        elif pattern == "descending":
            s = np.tile(np.linspace(3, -3, n, dtype=np.float32), (rows, 1))
        elif pattern == "constant":
            s = np.full((rows, n), 0.25, dtype=np.float32)
        else:
            s = rng.standard_normal((rows, n)).astype(np.float32)
            s[:, ::7] = np.round(s[:, ::7])
        st = torch.from_numpy(s)
        keys = L.topk(st.to(cuda_dev), k, id_base=11)
        ref = o.merge_keys(o.make_keys(s, np.broadcast_to(np.arange(n) + 11, s.shape)), k)
        assert (keys.cpu().numpy().view(np.uint64) == ref).all(), f"{pattern} n={n} rows={rows} k={k}"
    ids = torch.from_numpy(rng.permutation(2_000_000)[:300_000].astype(np.int32).reshape(3, 100_000))
    s = torch.from_numpy(rng.standard_normal((3, 100_000)).astype(np.float32))
    keys = L.topk(s.to(cuda_dev), 64, ids=ids.to(cuda_dev))
    assert (keys.cpu().numpy().view(np.uint64) == o.merge_keys(o.make_keys(s.numpy(), ids.numpy()), 64)).all()


def test_topk_with_explicit_ids_and_merge(cuda_dev):
    L = _lib()
    g = torch.Generator().manual_seed(9)
    s = torch.randn((3, 500), generator=g)
    ids = torch.stack([torch.randperm(10_000, generator=g)[:500] for _ in range(3)]).to(torch.int32)
    keys = L.topk(s.to(cuda_dev), 40, ids=ids.to(cuda_dev))
    ref = o.merge_keys(o.make_keys(s.numpy(), ids.numpy()), 40)
    assert (keys.cpu().numpy().view(np.uint64) == ref).all()
    # merge of 8 "ranks" of 100 keys, some empty
    parts = []
    for r in range(8):
        sr = torch.randn((3, 100), generator=g)
        kr = o.make_keys(sr.numpy(), np.broadcast_to(np.arange(100) + 1000 * r, sr.shape))
        if r == 5:
            kr[:, 50:] = 0
        parts.append(kr)
    allk = np.concatenate(parts, 1)
    got = L.topk_merge(torch.from_numpy(allk.view(np.int64).copy()).to(cuda_dev), 100)
    assert (got.cpu().numpy().view(np.uint64) == o.merge_keys(allk, 100)).all()
    few = L.topk_merge(torch.from_numpy(allk[:, :30].view(np.int64).copy()).to(cuda_dev), 100)
    assert (few.cpu().numpy().view(np.uint64) == o.merge_keys(allk[:, :30], 100)).all()


def test_rrf_bit_exact_against_reference_fixtures(cuda_dev, golden_dir):
    L = _lib()
    cases = json.load(open(os.path.join(golden_dir, "rrf.json")))
    for c in cases:
        a = torch.tensor([c["a"]], dtype=torch.int32, device=cuda_dev).reshape(1, -1)
        b = torch.tensor([c["b"]], dtype=torch.int32, device=cuda_dev).reshape(1, -1)
        n = len(c["a"]) + len(c["b"])
        ids, scores, counts = L.rrf_fuse(a, b, c["k"], n)
        cnt = int(counts[0])
        assert cnt == len(c["ids"])
        assert ids[0, :cnt].tolist() == c["ids"]
        assert [repr(x) for x in scores[0, :cnt].tolist()] == c["scores"]      # fp64 bit-exact
        assert (ids[0, cnt:] == -1).all()
        top = L.rrf_fuse(a, b, c["k"], 50)[0]                                 # the [:50] slice at :916
        assert top[0, : min(50, cnt)].tolist() == c["ids"][:50]


def test_rrf_batched_random_against_oracle(cuda_dev):
    L = _lib()
    rng = np.random.default_rng(4)
    a = rng.integers(0, 300, (64, 100)).astype(np.int32)           # repeats inside a list are legal
    b = rng.integers(0, 300, (64, 100)).astype(np.int32)
    a[3, 90:] = -1                                                   # absent entries pad the tail of a list
    ids, scores, counts = L.rrf_fuse(torch.from_numpy(a).to(cuda_dev), torch.from_numpy(b).to(cuda_dev), 60, 200)
    for r in range(64):
        ri, rs = o.rrf_ids(a[r].tolist(), b[r].tolist(), 60)
        assert int(counts[r]) == len(ri)
        assert ids[r, : len(ri)].tolist() == ri
        assert scores[r, : len(ri)].tolist() == rs


def test_synthetic_corpus_is_shard_invariant_and_normalised(cuda_dev):
    from hybrid_rag_colbertv2_b200.synth import synth_store
    full = synth_store(1000, 32, 512, seed=5, device=cuda_dev)
    assert full.n_docs == 1000
    n = full.tokens.float().norm(dim=-1)
    assert float((n - 1).abs().max()) < 2e-2
    parts = [synth_store(1000, 32, 512, seed=5, device=cuda_dev, rank=r, world_size=4) for r in range(4)]
    assert sum(p.n_docs for p in parts) == 1000
    assert torch.equal(torch.cat([p.tokens for p in parts]), full.tokens)
    assert [p.doc_id_base for p in parts] == [0] + list(np.cumsum([p.n_docs for p in parts])[:-1])
    other = synth_store(1000, 32, 512, seed=6, device=cuda_dev)
    assert not torch.equal(other.tokens[:100], full.tokens[:100])
    assert abs(float(full.tokens.float().mean())) < 1e-2


def test_sharded_search_equals_single_gpu(cuda_dev):
    """G shards simulated sequentially on one GPU: merged top-k == unsharded top-k, bit-exact."""
    import hybrid_rag_colbertv2_b200 as hrc
    from hybrid_rag_colbertv2_b200.synth import plant, synth_queries, synth_store
    L = _lib()
    full = synth_store(20_000, 32, 200, seed=21, device=cuda_dev)
    q = synth_queries(3, 32, device=cuda_dev)
    plant(full, q, n_planted=40)
    cfg = hrc.RAGConfig()
    one = hrc.JinaColBERTRetriever(cfg)
    one.store = full
    ref = one.search_keys(q, 100)
    for world in (2, 4, 8):
        gathered = []
        for r in range(world):
            rr = hrc.JinaColBERTRetriever(cfg)
            rr.store = full.shard(r, world)
            gathered.append(rr.search_keys(q, 100))
        merged = L.topk_merge(torch.cat(gathered, 1).contiguous(), 100)
        assert torch.equal(merged, ref), f"world={world}"


def test_retriever_api_shapes_and_ranking(cuda_dev, tmp_path):
    """search / rerank / _maxsim_score through the reference-shaped API, checked against the oracle."""
    import hybrid_rag_colbertv2_b200 as hrc
    cfg = hrc.RAGConfig(colbert_index_path=str(tmp_path / "colbert"))
    r = hrc.JinaColBERTRetriever(cfg)
    corpus = [f"document number {i} about topic {i % 7} and late interaction {i * 3}" for i in range(120)]
    r.index(corpus)
    assert os.path.exists(os.path.join(cfg.colbert_index_path, "index.hrc.pt"))        # packed store: its own file name
    assert not os.path.exists(os.path.join(cfg.colbert_index_path, "index.pt"))        # never a fake reference index
    res = r.search(query="late interaction topic 3", k=10)
    assert len(res) == 10 and all(sorted(x) == ["document_id", "score", "text"] for x in res)
    assert all(isinstance(x["document_id"], int) and isinstance(x["score"], float) for x in res)
    assert all(x["text"] == corpus[x["document_id"]] for x in res)
    qe = r.model.encode("late interaction topic 3", convert_to_tensor=True)
    exp = o.maxsim_scores(o.round_bf16(qe), r.store.tokens.float().cpu(), r.store.offsets.cpu())[0]
    assert o.check_ranking([x["document_id"] for x in res], [x["score"] for x in res], exp, 10, RTOL) is None
    assert len(r.search("anything", k=1000)) == 120                                  # k > N clamps (:767)
    # rerank: result_index indexes the input list
    docs = [corpus[i] for i in (5, 80, 33, 3, 17, 110)]
    rr = r.rerank(query="late interaction topic 3", documents=docs, k=4)
    assert len(rr) == 4 and [x["rank"] for x in rr] == [1, 2, 3, 4]
    assert all(sorted(x) == ["rank", "result_index", "score", "text"] for x in rr)
    assert all(x["text"] == docs[x["result_index"]] for x in rr)
    sub = exp[[5, 80, 33, 3, 17, 110]]
    assert o.check_ranking([x["result_index"] for x in rr], [x["score"] for x in rr], sub, 4, RTOL) is None
    assert len(r.rerank("q", docs[:3], k=10)) == 3
    # load() round trip, and the reference's dense index.pt layout
    r2 = hrc.JinaColBERTRetriever(cfg)
    r2.load()
    assert torch.equal(r2.store.tokens, r.store.tokens) and r2.corpus == corpus
    dense = torch.nn.functional.normalize(torch.randn(30, 16, 128), dim=-1)
    os.makedirs(tmp_path / "ref", exist_ok=True)
    torch.save({"embeddings": dense, "corpus": [f"doc {i}" for i in range(30)]}, tmp_path / "ref" / "index.pt")
    r3 = hrc.JinaColBERTRetriever(hrc.RAGConfig(colbert_index_path=str(tmp_path / "ref")))
    r3.load()
    assert r3.store.n_docs == 30 and r3.store.total_tokens == 480
    # _maxsim_score: shapes and squeeze of :813-831, values = true MaxSim
    q1 = torch.nn.functional.normalize(torch.randn(32, 128), dim=-1)
    s = r3._maxsim_score(q1, dense)
    assert s.shape == (30,)
    _assert_scores(s, o.maxsim_dense(o.round_bf16(q1), o.round_bf16(dense)), "_maxsim_score")
    assert r3._maxsim_score(torch.stack([q1, q1]), dense).shape == (2, 30)
    assert r3._maxsim_score(q1, dense[0]).dim() == 0                                 # 2-D docs = one document
    with pytest.raises(IndexError):
        r3._maxsim_score(torch.randn(128), dense)
    # N == 1 works here (the reference raises TypeError)
    r3.index_embeddings(dense[:1])
    assert len(r3.search("x", k=5)) == 1


def test_hybrid_retrieve_matches_stagewise_oracle(cuda_dev, tmp_path):
    import hybrid_rag_colbertv2_b200 as hrc
    cfg = hrc.RAGConfig(colbert_index_path=str(tmp_path / "colbert"), colbert_top_k=30, bm25_top_k=30,
                        rerank_candidates=20, final_top_k=5)
    idx = hrc.DualIndexer(cfg)
    corpus = [f"chunk {i} alpha beta {i % 11} gamma {i % 5}" for i in range(200)]
    idx.build_colbert_index(corpus)
    rng = np.random.default_rng(8)
    bm25_ids = rng.permutation(200)[:30].tolist()

    def bm25(query, k):
        return [{"chunk_id": int(i), "score": 1.0 / (j + 1), "source": "bm25"} for j, i in enumerate(bm25_ids[:k])]

    h = hrc.HybridRetriever(cfg, idx, None, bm25_search=bm25, verbose=False)
    out = h.retrieve("alpha gamma 3")
    assert len(out) == 5 and [x["rank"] for x in out] == [1, 2, 3, 4, 5]
    assert all(sorted(x) == sorted(["chunk_id", "text", "document_id", "heading_path", "has_images", "metadata",
                                    "score", "rank"]) for x in out)
    assert set(h.last_timings) == {"bm25", "colbert", "fusion", "fetch", "rerank", "total"}
    # stage-wise oracle
    retr = idx.colbert_retriever
    qe = o.round_bf16(retr.model.encode("alpha gamma 3", convert_to_tensor=True))
    scores = o.maxsim_scores(qe, retr.store.tokens.float().cpu(), retr.store.offsets.cpu())[0]
    col = h._colbert_search("alpha gamma 3", 30)
    assert o.check_ranking([c["chunk_id"] for c in col], [c["score"] for c in col], scores, 30, RTOL) is None
    fused = h._reciprocal_rank_fusion(bm25("q", 30), col)
    assert fused == o.rrf_reference(bm25("q", 30), col)                          # ids, fp64 scores, order
    cand = [f["chunk_id"] for f in fused[:20]]
    assert o.check_ranking([cand.index(x["chunk_id"]) for x in out], [x["score"] for x in out],
                           scores[cand], 5, RTOL) is None
    # batched device pipeline gives the same ids
    ids, sc = h.retrieve_batch(qe.unsqueeze(0), torch.tensor([bm25_ids], dtype=torch.int32), top_k_final=5)
    assert ids[0].tolist() == [x["chunk_id"] for x in out]


def test_fused_hybrid_retrieve_equals_staged_pipeline(cuda_dev):
    """hrc_hybrid_retrieve (search -> RRF -> rerank in one C call) == the staged calls, bit for bit, also on a
    document shard with a non-zero id base."""
    import hybrid_rag_colbertv2_b200 as hrc
    from hybrid_rag_colbertv2_b200.synth import plant, synth_queries, synth_store
    full = synth_store(30_000, 16, 200, seed=31, device=cuda_dev)
    q = synth_queries(5, 32, device=cuda_dev)
    plant(full, q, n_planted=40)
    g = torch.Generator().manual_seed(6)
    for store in (full, full.shard(1, 3)):
        cfg = hrc.RAGConfig(colbert_top_k=100, rerank_candidates=50, final_top_k=10)
        idx = hrc.DualIndexer(cfg)
        idx.colbert_retriever.store = store
        h = hrc.HybridRetriever(cfg, idx, None, verbose=False)
        lo, n = store.doc_id_base, store.n_docs
        bm25 = (torch.randint(0, n, (5, 100), generator=g, dtype=torch.int32) + lo).to(cuda_dev)
        col_ids, _ = idx.colbert_retriever.search_embeddings(q, 100)
        bm25[:, :20] = col_ids[:, torch.randperm(100, generator=g)[:20].to(cuda_dev)]   # overlap with the ColBERT list
        bm25[2, 90:] = -1
        ids_f, sc_f = h.retrieve_batch(q, bm25)
        ids_s, sc_s = h._retrieve_batch_staged(q, bm25, 10)
        assert torch.equal(ids_f, ids_s) and torch.equal(sc_f, sc_s)
        assert bool(((ids_f >= lo) & (ids_f < lo + n)).all())
        assert bool((sc_f[:, :-1] >= sc_f[:, 1:]).all())


def test_full_size_properties_c2(cuda_dev):
    """BASELINE config C2 at full size (1M docs x 128 tokens, 32.8 GB): size-independent checks."""
    import hybrid_rag_colbertv2_b200 as hrc
    from hybrid_rag_colbertv2_b200.synth import plant, synth_queries, synth_store
    L = _lib()
    free, _ = torch.cuda.mem_get_info()
    n_docs = 1_000_000 if free > 60e9 else 100_000
    store = synth_store(n_docs, 128, 128, seed=20260102, device=cuda_dev)
    q = synth_queries(1, 32, device=cuda_dev)
    planted = plant(store, q, n_planted=200)
    r = hrc.JinaColBERTRetriever(hrc.RAGConfig())
    r.store = store
    scores = r.score_embeddings(q)
    assert scores.shape == (1, n_docs) and bool(torch.isfinite(scores).all())
    # (1) a random sample of documents + all planted ones re-scored by the oracle
    g = torch.Generator().manual_seed(1)
    sample = torch.cat([torch.randint(0, n_docs, (3000,), generator=g), planted[0]])
    sub_tok = store.tokens.view(n_docs, 128, 128)[sample.to(cuda_dev)].reshape(-1, 128).float().cpu()
    exp = o.maxsim_scores(q.float().cpu(), sub_tok, torch.arange(0, sample.numel() * 128 + 1, 128))[0]
    _assert_scores(scores[0, sample.to(cuda_dev)], exp, "C2 sample")
    # (2) SIMT path agrees on the same sample through the candidate entry point
    simt = L.maxsim_scores_ids(store.tokens, store.offsets, sample.to(torch.int32).unsqueeze(0).to(cuda_dev), q,
                               path=L.PATH_SIMT)
    _assert_scores(simt[0], exp, "C2 sample simt")
    # (3) top-100: sorted, unique, scores equal the score array, nothing outside beats the k-th
    ids, sc = r.search_embeddings(q, 100)
    ids_c, sc_c = ids[0].cpu(), sc[0].cpu()
    assert len(set(ids_c.tolist())) == 100 and bool((sc_c[:-1] >= sc_c[1:]).all())
    assert torch.equal(scores[0, ids[0].to(torch.int64)].cpu(), sc_c)
    assert int((scores[0] > sc_c[-1]).sum()) <= 99
    tv, ti = torch.topk(scores[0], 100)
    assert torch.equal(tv.cpu(), sc_c)
    # (4) the planted documents dominate the ranking, in the oracle's order
    top_planted = [i for i in ids_c.tolist() if i in set(planted[0].tolist())]
    assert len(top_planted) >= 90
    full_exp = torch.full((n_docs,), float('-inf'))
    full_exp[sample] = exp
    assert o.check_ranking(ids_c.tolist()[:50], sc_c.tolist()[:50], full_exp, 50, RTOL) is None
    # (5) shard invariance: scores of a 4-way document split are bit-identical
    parts = [L.maxsim_scores(s.tokens, s.offsets, q) for s in (store.shard(rk, 4) for rk in range(4))]
    assert torch.equal(torch.cat(parts, 1), scores)


def test_full_size_properties_c3(cuda_dev):
    """BASELINE config C3 at FULL size — 256 queries x 32 tokens over 1M documents of 32..512 tokens (~70 GB) —
    plus a 24-query run of the same corpus (a CTA-pair launch plus the odd group on the single-CTA kernel):
    size-independent checks."""
    import hybrid_rag_colbertv2_b200 as hrc
    from hybrid_rag_colbertv2_b200.synth import plant, synth_queries, synth_store
    L = _lib()
    free, _ = torch.cuda.mem_get_info()
    n_docs = 1_000_000 if free > 100e9 else 100_000
    store = synth_store(n_docs, 32, 512, seed=20260103, device=cuda_dev)
    nq = 256 if n_docs == 1_000_000 else 24
    q = synth_queries(nq, 32, device=cuda_dev)
    planted = plant(store, q[:2], n_planted=60)
    r = hrc.JinaColBERTRetriever(hrc.RAGConfig())
    r.store = store
    scores = r.score_embeddings(q)
    assert scores.shape == (nq, n_docs) and bool(torch.isfinite(scores).all())
    # (1) sampled documents (+ the planted ones) re-scored by the oracle, every query
    g = torch.Generator().manual_seed(2)
    sample = torch.unique(torch.cat([torch.randint(0, n_docs, (1500,), generator=g), planted.reshape(-1),
                                     torch.tensor([0, n_docs - 1])]))
    off = store.offsets.cpu()
    lens = (off[1:] - off[:-1])[sample]
    rows = torch.cat([torch.arange(int(off[d]), int(off[d + 1])) for d in sample.tolist()])
    sub_tok = store.tokens[rows.to(cuda_dev)].float().cpu()
    sub_off = torch.zeros(sample.numel() + 1, dtype=torch.int64)
    sub_off[1:] = torch.cumsum(lens, 0)
    exp = o.maxsim_scores(q.float().cpu(), sub_tok, sub_off)
    _assert_scores(scores[:, sample.to(cuda_dev)], exp, "C3 sample")
    # (2) every query of the batch equals the single-query kernel on the same corpus (different kernels,
    #     same fp32 accumulation): within tolerance everywhere
    for qi in (0, 7, 8, 15, 16, 23, nq - 1):
        single = L.maxsim_scores(store.tokens, store.offsets, q[qi:qi + 1])
        scale = float(single.abs().max())
        assert float((single[0] - scores[qi]).abs().max()) <= TIGHT * scale, f"query {qi}"
    # (3) top-100 per query: sorted, unique, consistent with the score matrix and with torch.topk's values
    ids, sc = r.search_embeddings(q, 100)
    for qi in range(nq):
        ids_c, sc_c = ids[qi].cpu(), sc[qi].cpu()
        assert len(set(ids_c.tolist())) == 100 and bool((sc_c[:-1] >= sc_c[1:]).all())
        assert torch.equal(scores[qi, ids[qi].to(torch.int64)].cpu(), sc_c)
        assert torch.equal(torch.topk(scores[qi], 100).values.cpu(), sc_c)
    # (4) planted documents lead the two planted queries, in the oracle's order
    for qi in range(2):
        full_exp = torch.full((n_docs,), float('-inf'))
        full_exp[sample] = exp[qi]
        assert o.check_ranking(ids[qi].cpu().tolist()[:30], sc[qi].cpu().tolist()[:30], full_exp, 30, RTOL) is None
    # (5) shard invariance: a 3-way document split gives bit-identical scores; and the first 24 queries alone
    #     (CTA pairs + the odd group on the single-CTA kernel) equal their rows of the full batch bit for bit
    q24 = q[:24].contiguous()
    s24 = L.maxsim_scores(store.tokens, store.offsets, q24)
    assert torch.equal(s24, scores[:24])
    parts = [L.maxsim_scores(s.tokens, s.offsets, q24) for s in (store.shard(rk, 3) for rk in range(3))]
    assert torch.equal(torch.cat(parts, 1), s24)


def test_randomised_shapes_against_oracle(cuda_dev):
    """Seeded fuzz through the C ABI: document-length mixes (empty, 1-3 tokens, around the 32-column chunk and the
    128-token tile, several tiles long), query counts that hit every kernel (1-4: HBM-bound variants, 5-8: single-CTA
    batched, 9+: CTA pairs with and without an odd group) and query lengths on both sides of the 32-token slot."""
    L = _lib()
    rng = np.random.default_rng(20260118)
    modes = {
        "tiny": lambda n: rng.integers(0, 4, n),
        "chunk": lambda n: rng.integers(28, 37, n),
        "tile": lambda n: rng.integers(120, 137, n),
        "mixed": lambda n: np.where(rng.random(n) < 0.3, rng.integers(0, 6, n), rng.integers(1, 300, n)),
        "long": lambda n: rng.integers(200, 900, n),
    }
    nqs = [1, 2, 3, 4, 5, 8, 9, 16, 17, 24, 33]
    lqs = [1, 7, 16, 31, 32, 33, 64, 100]
    for case in range(44):
        mode = list(modes)[case % len(modes)]
        n_docs = int(rng.integers(1, 260 if mode != "long" else 40))
        lens = modes[mode](n_docs).astype(np.int64)
        if lens.sum() == 0:
            lens[0] = 5
        nq, lq = nqs[case % len(nqs)], lqs[(case * 3) % len(lqs)]
        q, tok, off = _case(1000 + case, 0, 0, 0, nq, lq, lens=lens)
        exp = o.maxsim_scores(q.float(), tok.float(), off)
        got = L.maxsim_scores(tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev), path=L.PATH_TC)
        torch.cuda.synchronize()
        _assert_scores(got, exp, f"case {case}: {mode} docs={n_docs} nq={nq} lq={lq}")
        if case % 4 == 0:
            simt = L.maxsim_scores(tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev), path=L.PATH_SIMT)
            _assert_scores(simt, exp, f"case {case} simt")


def test_plain_c_client_of_the_abi(cuda_dev, tmp_path):
    """tests/c_abi/client.c: a C99 program (no Python, no torch, no C++) drives libhrc.so through include/hrc.h and
    cross-checks the tensor-core path against the CUDA-core path and a host loop."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    libdir = os.path.join(root, "hybrid-rag-colbertv2_b200")
    exe = str(tmp_path / "client")
    cc = shutil.which("gcc") or shutil.which("cc")
    assert cc is not None
    subprocess.run([cc, "-O2", "-std=c99", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
                    os.path.join(root, "tests", "c_abi", "client.c"), "-o", exe, "-L", libdir, "-l:libhrc.so",
                    "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{libdir}",
                    f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c_abi_client ok" in out.stdout


# ======================================================================================================
# Round 2 hardening (VERDICT r1 "Next" #2): adversarial boundaries, reference pin, float64 known answers,
# full-size C4, ownership of returned buffers, large k.
# ======================================================================================================
BOUNDARY_POS = [0, 1, 31, 32, 33, 63, 64, 65, 95, 96, 97, 126, 127]      # positions inside a 128-token tile
DOC_POS = [0, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257]      # positions inside a document


def _boundary_corpus(nq, lq, seed=5):
    """A corpus in which every document holds exactly ONE decisive token — an exact copy of one query token, cosine
    ~1 against a background <= ~0.35 — placed at a chosen position relative to the document start AND to the
    128-token tile (filler documents shift the alignment), so that losing or doubling a single accumulator column at
    a 32-column chunk edge, the pulled-back last load, the CTA-pair half-tile seam (column 64) or a tile edge moves
    the score by > 0.5.  Returns q, tok, off and, per document, (query, query token, token row)."""
    g = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn((nq, lq, 128), generator=g), dim=-1).to(torch.bfloat16)
    lens, plant_at = [], []
    cur = 0
    for tpos in BOUNDARY_POS:
        for dpos in DOC_POS:
            fill = (tpos - dpos - cur) % 128              # filler document so that (cur + fill + dpos) % 128 == tpos
            if fill:
                lens.append(fill)
                plant_at.append(fill - 1)                 # fillers are planted too: at their LAST token
                cur += fill
            extra = [0, 1, 31, 40, 130][(tpos + dpos) % 5]   # tokens after the decisive one (0: it is the last token)
            lens.append(dpos + 1 + extra)
            plant_at.append(dpos)
            cur += dpos + 1 + extra
    off = torch.zeros(len(lens) + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(torch.tensor(lens), 0)
    tok = torch.nn.functional.normalize(torch.randn((int(off[-1]), 128), generator=g), dim=-1).to(torch.bfloat16)
    planted = []
    for d, p in enumerate(plant_at):
        qi, ti = d % nq, (d * 7) % lq
        row = int(off[d]) + p
        tok[row] = q[qi, ti]
        planted.append((qi, ti, row))
    return q, tok, off, planted


@pytest.mark.parametrize("nq,path", [(1, "tc"), (2, "tc"), (3, "tc"), (1, "tc_dm"), (1, "auto"), (8, "tc"),
                                     (16, "tc"), (24, "tc"), (2, "simt")])
def test_decisive_token_at_every_chunk_and_tile_boundary(cuda_dev, nq, path):
    L = _lib()
    q, tok, off, planted = _boundary_corpus(nq, 32)
    n_docs = off.numel() - 1
    assert int(off[-1]) > 148 * 128 * 2                    # several tiles per CTA segment, every SM busy
    covered = {(row % 128) for _, _, row in planted}
    assert set(BOUNDARY_POS) <= covered
    exp = o.maxsim_scores(q.float(), tok.float(), off)
    # premise: for the query token it copies, the decisive token IS the document's max by > 0.5 — losing its
    # accumulator column (or reading a neighbour's instead) moves the document's score by more than 0.5
    tf = tok.float()
    for d, (qi, ti, row) in enumerate(planted):
        sims = tf[int(off[d]):int(off[d + 1])] @ q[qi, ti].float()
        p = row - int(off[d])
        rest = torch.cat([sims[:p], sims[p + 1:]])
        assert float(sims[p]) > 0.95 and (rest.numel() == 0 or float(sims[p] - rest.max()) > 0.5)
    tok_d, off_d, q_d = tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev)
    got = L.maxsim_scores(tok_d, off_d, q_d, path=_path(L, path))
    _assert_scores(got, exp, f"boundary nq={nq} {path}", bucket="boundary")
    worst = float((got.float().cpu() - exp).abs().max())
    assert worst < 1e-3, f"a boundary column was lost or doubled: {worst}"
    if path == "tc":       # the candidate (rerank) entry point walks ONE document per CTA: same positions, other code path
        cand = torch.arange(n_docs, dtype=torch.int32).flip(0).unsqueeze(0).repeat(nq, 1).contiguous()
        gotc = L.maxsim_scores_ids(tok_d, off_d, cand.to(cuda_dev), q_d, path=L.PATH_TC)
        _assert_scores(gotc, exp.flip(1), f"boundary candidates nq={nq}", bucket="boundary")


@pytest.mark.parametrize("path", ["tc_dm", "auto"])
def test_boundary_suite_through_the_dynamic_work_units(cuda_dev, path):
    """The doc-major kernel hands the last eighth of a large corpus out in small shared units claimed at run time (which
    CTA scores which documents then differs from launch to launch).  The boundary corpus, repeated (on the device) until
    the dynamic route engages, must still match the oracle on every document — the oracle's scores of the base corpus,
    tiled — 5 launches in a row bit for bit, through the score matrix AND through the fused top-k search, and must equal
    the static distribution (no workspace -> no claim counter) bit for bit."""
    L = _lib()
    q, tok, off, planted = _boundary_corpus(1, 32)
    # the dynamic route needs >= 8 shared units of >= 16 average documents per CTA in the last eighth of the corpus
    mean_len = int(off[-1]) // (off.numel() - 1) + 1
    need = 148 * 8 * (8 * 16 * mean_len)
    reps = int(need * 1.25 / int(off[-1])) + 1
    tok_d = tok.to(cuda_dev).repeat(reps, 1)
    off_r = torch.cat([off[:1]] + [off[1:] + r * int(off[-1]) for r in range(reps)])
    assert int(off_r[-1]) >= need and tok_d.shape[0] == int(off_r[-1])
    off_d, q_d = off_r.to(cuda_dev), q.to(cuda_dev)
    exp = o.maxsim_scores(q.float(), tok.float(), off).repeat(1, reps)
    first = None
    for _ in range(5):
        got = L.maxsim_scores(tok_d, off_d, q_d, path=_path(L, path))
        keys = L.search(tok_d, off_d, q_d, 100, path=_path(L, path))[0]
        if first is None:
            _assert_scores(got, exp, f"dynamic units {path}", bucket="boundary")
            first = (got.clone(), keys.clone())
        assert torch.equal(got, first[0]) and torch.equal(keys, first[1]), "results differ between launches"
    assert torch.equal(first[1], L.topk(first[0], 100)), "fused search over dynamic units != top-k of the score matrix"
    static = torch.empty_like(first[0])
    rc = L.load().hrc_maxsim_scores(tok_d.data_ptr(), off_d.data_ptr(), off_d.numel() - 1, tok_d.shape[0], q_d.data_ptr(), 1, 32,
                                    static.data_ptr(), _path(L, path), None, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert torch.equal(static, first[0]), "dynamic and static distributions differ"


@pytest.mark.parametrize("path", ["tc", "simt", "tc_dm"])
def test_maxsim_kernels_reproduce_the_reference_where_it_computes_maxsim(cuda_dev, golden_dir, path):
    """REFERENCE PIN for the MaxSim kernels: tests/golden/maxsim_pin.npz holds outputs of the UNMODIFIED reference
    `_maxsim_score` (local_rag_complete.py:821-829) on inputs where its mean-pool cosine equals MaxSim (identical
    query tokens, identical document tokens, exactly unit-norm dyadic rows; make_golden.py asserts the premise).
    hrc_maxsim_scores / Lq must reproduce them BIT FOR BIT (every partial sum is exactly representable)."""
    import hybrid_rag_colbertv2_b200 as hrc
    L = _lib()
    z = np.load(os.path.join(golden_dir, "maxsim_pin.npz"))
    qrows, drows = torch.from_numpy(z["q_rows"]), torch.from_numpy(z["d_rows"])
    for lq in (1, 2, 4):
        for ld in (1, 2, 4):
            ref = torch.from_numpy(z[f"out_lq{lq}_ld{ld}"])
            q = qrows[:, None, :].expand(-1, lq, -1).contiguous().to(torch.bfloat16)
            tok = drows[:, None, :].expand(-1, ld, -1).reshape(-1, 128).contiguous().to(torch.bfloat16)
            off = torch.arange(0, drows.shape[0] * ld + 1, ld, dtype=torch.int64)
            for qs in (q, q[:1].contiguous(), q[:2].contiguous()):          # batched (5) and the 1- / 2-query kernels
                got = L.maxsim_scores(tok.to(cuda_dev), off.to(cuda_dev), qs.to(cuda_dev), path=_path(L, path))
                assert torch.equal(got.cpu() / lq, ref[: qs.shape[0]]), (path, lq, ld, qs.shape[0])
    # through the reference-shaped API: score_reduction="mean" is the reference's scale on these inputs
    r = hrc.JinaColBERTRetriever(hrc.RAGConfig(score_reduction="mean", maxsim_path=_path(L, path)))
    dense = drows[:, None, :].expand(-1, 4, -1).contiguous()
    r.index_embeddings(dense)
    ref = torch.from_numpy(z["out_lq2_ld4"])
    s = r._maxsim_score(qrows[:, None, :].expand(-1, 2, -1).contiguous(), dense)           # [Bq, N], as :831
    assert torch.equal(s.cpu(), ref)
    ids, sc = r.search_embeddings(qrows[0][None, :].expand(2, -1).contiguous(), 7)
    assert torch.equal(sc[0].cpu(), torch.topk(ref[0], 7).values)                          # the reference's :767
    assert ids[0, 0].item() == 0 and sc[0, 0].item() == 1.0                                # the query's own copy


@pytest.mark.parametrize("path", ["tc", "simt", "tc_dm"])
def test_maxsim_kernels_match_float64_known_answers(cuda_dev, golden_dir, path):
    L = _lib()
    z = np.load(os.path.join(golden_dir, "maxsim_kat_f64.npz"))
    for name in ("a", "b"):
        q, tok, off = (torch.from_numpy(z[f"{name}_{k}"]) for k in ("q", "tok", "off"))
        got = L.maxsim_scores(tok.to(torch.bfloat16).to(cuda_dev), off.to(cuda_dev), q.to(torch.bfloat16).to(cuda_dev),
                              path=_path(L, path))
        _assert_scores(got, torch.from_numpy(z[f"{name}_scores_f64"]).float(), f"kat {name} {path}", bucket="kat_f64")


@pytest.mark.parametrize("path", ["tc", "simt", "tc_dm", "auto"])
def test_maxsim_kernels_match_vllm_outputs(cuda_dev, golden_dir, path):
    """The CUDA kernels against scores vLLM 0.22.0's MaxSim functions produced (tests/golden/make_vllm_pin.py): an
    implementation neither this repo's oracle nor its kernels had a hand in.  Also through the one-call search: the
    top-10 it returns are vLLM's top-10 in vLLM's order (gaps permitting)."""
    from golden.make_vllm_pin import load_pin
    L = _lib()
    z = load_pin(os.path.join(golden_dir, "maxsim_vllm_pin.npz"))
    tok_d = torch.from_numpy(z["tok"]).to(torch.bfloat16).to(cuda_dev)
    off_d = torch.from_numpy(z["off"]).to(cuda_dev)
    for name in ("q32", "q7", "q1"):
        q, pair, _ = z[name]
        q_d = torch.from_numpy(q).to(torch.bfloat16).to(cuda_dev)
        got = L.maxsim_scores(tok_d, off_d, q_d, path=_path(L, path))
        _assert_scores(got, torch.from_numpy(pair), f"vllm pin {name} {path}", bucket="vllm_pin")
        _, ids, sc = L.search(tok_d, off_d, q_d, 10, path=_path(L, path))
        for b in range(q.shape[0]):
            err = o.check_ranking(ids[b].tolist(), sc[b].tolist(), torch.from_numpy(pair[b]), 10, TIGHT)
            assert err is None, f"{name} {path} query {b}: {err}"


def test_search_host_results_belong_to_the_caller(cuda_dev):
    """two consecutive search_host results held at once must not alias (pinned staging is reused)."""
    import hybrid_rag_colbertv2_b200 as hrc
    q, tok, off = _case(43, 3000, 8, 60, 2, 32)
    r = hrc.JinaColBERTRetriever(hrc.RAGConfig())
    r.index_embeddings(tok, off, packed=True)
    a_ids, a_sc = r.search_host(q[0].float(), 20)
    keep_ids, keep_sc = a_ids.clone(), a_sc.clone()
    b_ids, b_sc = r.search_host(q[1].float(), 20)
    assert torch.equal(a_ids, keep_ids) and torch.equal(a_sc, keep_sc)
    assert not torch.equal(a_ids, b_ids)
    z_ids, _ = r.search_host(q[0].float(), 20, copy=False)                 # explicit zero-copy: the staging buffer
    assert z_ids.is_pinned() and torch.equal(z_ids, keep_ids)


def test_k_beyond_the_selection_limit_and_absent_candidates(cuda_dev):
    """the reference's torch.topk / argsort take any k; k > HRC_MAX_TOPK goes through a device sort with the
    kernels' order.  Candidates this store does not hold score -inf and are dropped from _colbert_rerank."""
    import hybrid_rag_colbertv2_b200 as hrc
    q, tok, off = _case(47, 4000, 4, 30, 1, 32)
    r = hrc.JinaColBERTRetriever(hrc.RAGConfig())
    r.index_embeddings(tok, off, packed=True, corpus=[f"t{i}" for i in range(4000)])
    ids_big, sc_big = r.search_embeddings(q, 3000)
    ids_small, sc_small = r.search_embeddings(q, 2048)
    assert ids_big.shape == (1, 3000) and torch.equal(ids_big[:, :2048], ids_small) and torch.equal(sc_big[:, :2048], sc_small)
    assert len(r.search("anything", k=2500)) == 2500
    cand = torch.arange(4000, dtype=torch.int32).unsqueeze(0)
    pos, dids, sc = r.rerank_ids(q, cand, k=2100)
    assert torch.equal(dids[:, :2048], ids_small) and pos.shape == (1, 2100)
    idx = hrc.DualIndexer(hrc.RAGConfig(final_top_k=5))
    idx.colbert_retriever = r
    h = hrc.HybridRetriever(idx.config, idx, None, verbose=False)
    chunks = [{"chunk_id": c, "text": f"t{c}", "document_id": c, "metadata": {}} for c in (5, 999_999, 7, -3)]
    out = h._colbert_rerank("a query", chunks, top_k=4)
    assert [x["chunk_id"] for x in out] and all(x["chunk_id"] in (5, 7) for x in out) and len(out) == 2
    assert [x["rank"] for x in out] == [1, 2] and all(np.isfinite(x["score"]) for x in out)


def test_index_files_and_reference_compatible_export(cuda_dev, tmp_path):
    """the packed store is saved as index.hrc.pt (never as a file the reference would mis-read), doc_id_base
    is persisted, and reference_compatible_index=True additionally writes the reference's own dense layout."""
    import hybrid_rag_colbertv2_b200 as hrc
    cfg = hrc.RAGConfig(colbert_index_path=str(tmp_path / "ix"), reference_compatible_index=True)
    r = hrc.JinaColBERTRetriever(cfg, encoder=hrc.SyntheticEncoder(doc_tokens=8))
    corpus = [f"text {i} of the corpus" for i in range(40)]
    r.index(corpus)
    ref_file = torch.load(os.path.join(cfg.colbert_index_path, "index.pt"))
    assert sorted(ref_file) == ["corpus", "embeddings", "lengths"]
    assert ref_file["embeddings"].shape == (40, 8, 128) and ref_file["embeddings"].dtype == torch.float32   # :735-746
    r.store.doc_id_base = 1234
    r._save_index()
    back = hrc.JinaColBERTRetriever(cfg, encoder=hrc.SyntheticEncoder(doc_tokens=8))
    back.load()
    assert back.store.doc_id_base == 1234 and torch.equal(back.store.tokens, r.store.tokens) and back.corpus == corpus
    os.remove(os.path.join(cfg.colbert_index_path, "index.hrc.pt"))
    dense_only = hrc.JinaColBERTRetriever(cfg, encoder=hrc.SyntheticEncoder(doc_tokens=8))
    dense_only.load()                                                      # falls back to the reference's file
    assert torch.equal(dense_only.store.tokens, r.store.tokens) and dense_only.store.n_docs == 40


def test_full_size_c4_fused_equals_staged_equals_stagewise_oracle(cuda_dev):
    """BASELINE config C4 at full size: 1,000 queries through the hybrid pipeline over 1M passages x 128 tokens.
    fused (hrc_hybrid_retrieve, batches of 64) == staged calls bit for bit for all 1,000 queries; for sampled queries
    every stage is re-derived with the oracle: the ColBERT top-100 against oracle re-scores, RRF against the restated
    reference (:960-978, exact), and the final top-10 against oracle re-scores of the 50 candidates."""
    import hybrid_rag_colbertv2_b200 as hrc
    from hybrid_rag_colbertv2_b200.synth import plant, synth_queries, synth_store
    free, _ = torch.cuda.mem_get_info()
    n_docs = 1_000_000 if free > 60e9 else 50_000
    n_queries = 1000 if n_docs == 1_000_000 else 128
    store = synth_store(n_docs, 128, 128, seed=20260102, device=cuda_dev)
    queries = synth_queries(n_queries, 32, seed=99, device=cuda_dev)
    sample_q = [0, 63, 64, n_queries // 2, n_queries - 1]
    plant(store, queries[sample_q], n_planted=120, seed=17)
    cfg = hrc.RAGConfig(colbert_top_k=100, bm25_top_k=100, rerank_candidates=50, final_top_k=10)
    idx = hrc.DualIndexer(cfg)
    idx.colbert_retriever.store = store
    h = hrc.HybridRetriever(cfg, idx, None, verbose=False)
    g = torch.Generator().manual_seed(4)
    bm25 = torch.randint(0, n_docs, (n_queries, 100), generator=g, dtype=torch.int32).to(cuda_dev)
    fused_ids, fused_sc, staged_ids, staged_sc, col = [], [], [], [], []
    for b in range(0, n_queries, 64):
        qb = queries[b:b + 64]
        col_ids, _ = idx.colbert_retriever.search_embeddings(qb, 100)
        bm = bm25[b:b + 64].clone()
        bm[:, :30] = col_ids[:, torch.randperm(100, generator=g)[:30].to(cuda_dev)]       # 30 % overlap
        bm25[b:b + 64] = bm
        col.append(col_ids)
        i_f, s_f = h.retrieve_batch(qb, bm)
        i_s, s_s = h._retrieve_batch_staged(qb, bm, 10)
        fused_ids.append(i_f); fused_sc.append(s_f); staged_ids.append(i_s); staged_sc.append(s_s)
    fused_ids, fused_sc = torch.cat(fused_ids), torch.cat(fused_sc)
    assert torch.equal(fused_ids, torch.cat(staged_ids)) and torch.equal(fused_sc, torch.cat(staged_sc))
    assert fused_ids.shape == (n_queries, 10) and bool((fused_ids >= 0).all())
    assert bool((fused_sc[:, :-1] >= fused_sc[:, 1:]).all())
    col = torch.cat(col)
    view = store.tokens.view(n_docs, 128, 128)
    for qi in sample_q:
        qf = queries[qi:qi + 1].float().cpu()
        # stage 2: the ColBERT list re-scored by the oracle (+ 2000 random documents: nothing outside beats the k-th)
        extra = torch.randint(0, n_docs, (2000,), generator=g)
        docs = torch.unique(torch.cat([col[qi].cpu().long(), extra]))
        sub = view[docs.to(cuda_dev)].reshape(-1, 128).float().cpu()
        exp = o.maxsim_scores(qf, sub, torch.arange(0, docs.numel() * 128 + 1, 128))[0]
        full = torch.full((n_docs,), float("-inf"))
        full[docs] = exp
        _, sc100 = idx.colbert_retriever.search_embeddings(queries[qi:qi + 1], 100)
        assert o.check_ranking(col[qi].cpu().tolist(), sc100[0].cpu().tolist(), full, 100, RTOL) is None, f"query {qi}"
        # stage 3: RRF, exact
        ri, _ = o.rrf_ids(bm25[qi].cpu().tolist(), col[qi].cpu().tolist(), 60)
        cand = ri[:50]
        # stage 5: rerank of the 50 candidates against oracle re-scores
        cdocs = torch.tensor(cand)
        csub = view[cdocs.to(cuda_dev)].reshape(-1, 128).float().cpu()
        cexp = o.maxsim_scores(qf, csub, torch.arange(0, 50 * 128 + 1, 128))[0]
        got_pos = [cand.index(i) for i in fused_ids[qi].cpu().tolist()]
        assert o.check_ranking(got_pos, fused_sc[qi].cpu().tolist(), cexp, 10, RTOL) is None, f"query {qi}"


# ======================================================================================================
# Fused paths: top-k inside the MaxSim epilogue (hrc_search) and the one-launch rerank (hrc_rerank)
# ======================================================================================================
def _ascending_corpus(n_docs, lq, descending=False, seed=3):
    """One-token documents whose score against the query grows (or falls) with the document id: every document beats
    the running k-th best, so the per-warp key lists of the fused kernel fill and compact as often as they can."""
    g = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn((1, lq, 128), generator=g), dim=-1)
    qm = torch.nn.functional.normalize(q[0].mean(0), dim=-1)
    r = torch.nn.functional.normalize(torch.randn(128, generator=g), dim=-1)
    r = torch.nn.functional.normalize(r - (r @ qm) * qm, dim=-1)
    c = torch.linspace(-0.9, 0.9, n_docs)
    if descending:
        c = c.flip(0)
    tok = c[:, None] * qm[None, :] + (1 - c * c).sqrt()[:, None] * r[None, :]
    off = torch.arange(0, n_docs + 1, dtype=torch.int64)
    return q.to(torch.bfloat16), tok.to(torch.bfloat16), off


@pytest.mark.parametrize("nq", [1, 2, 3, 4, 7, 16, 24])
@pytest.mark.parametrize("corpus", ["random_short", "ascending", "descending"])
def test_fused_topk_equals_score_matrix_topk(cuda_dev, nq, corpus):
    """hrc_search's fused route (per-warp key lists in the MaxSim epilogue -> in-CTA merge -> one merge launch) against
    the staged route (score matrix -> radix top-k), bit for bit, on corpora with thousands of documents per warp: many
    list compactions, every kernel organisation (1 / 2 / 4 queries, single-CTA batched, CTA pairs, pairs + odd group),
    several k, a non-zero id base."""
    L = _lib()
    if corpus == "random_short":
        q, tok, off = _case(100 + nq, 300_000, 1, 3, nq, 32)
    else:
        q1, tok, off = _ascending_corpus(200_000, 32, descending=(corpus == "descending"))
        g = torch.Generator().manual_seed(nq)
        q = torch.cat([q1, torch.nn.functional.normalize(torch.randn((nq - 1, 32, 128), generator=g), dim=-1).to(torch.bfloat16)]) \
            if nq > 1 else q1
    tok_d, off_d, q_d = tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev)
    scores = L.maxsim_scores(tok_d, off_d, q_d, path=L.PATH_TC)
    for k, base in ((100, 0), (128, 7), (1, 0), (37, 1_000_000)):
        keys, ids, sc = L.search(tok_d, off_d, q_d, k, id_base=base)
        ref = L.topk(scores, k, id_base=base)
        assert torch.equal(keys, ref), f"{corpus} nq={nq} k={k}"
        ri, rs = L.keys_unpack(ref)
        assert torch.equal(ids, ri) and torch.equal(sc, rs)
    if nq == 1:                                                          # the doc-major kernel's fused top-k
        dm_scores = L.maxsim_scores(tok_d, off_d, q_d, path=L.PATH_TC_DM)
        for k, base in ((100, 0), (128, 7), (1, 3)):
            assert torch.equal(L.search(tok_d, off_d, q_d, k, id_base=base, path=L.PATH_TC_DM)[0], L.topk(dm_scores, k, id_base=base))
    keys129 = L.search(tok_d, off_d, q_d, 129)[0]                       # k > 128: the staged route
    assert torch.equal(keys129, L.topk(scores, 129))
    if nq == 1:                                                          # oracle on the adversarial corpora
        exp = o.maxsim_scores(q.float(), tok.float(), off)
        _, ids, sc = L.search(tok_d, off_d, q_d, 100)
        assert o.check_ranking(ids[0].tolist(), sc[0].tolist(), exp[0], 100, RTOL) is None


def test_fused_topk_with_empty_and_tiny_inputs(cuda_dev):
    """Fewer documents than list slots, empty documents (-inf keys), k == n_docs, one tile."""
    L = _lib()
    q, tok, off = _case(12, 0, 0, 0, 3, 32, lens=[0, 5, 0, 0, 130, 1, 0, 2, 9])
    tok_d, off_d, q_d = tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev)
    scores = L.maxsim_scores(tok_d, off_d, q_d)
    for k in (1, 5, 9):
        assert torch.equal(L.search(tok_d, off_d, q_d, k)[0], L.topk(scores, k))
    q, tok, off = _case(13, 3, 4, 9, 1, 7)
    tok_d, off_d, q_d = tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev)
    assert torch.equal(L.search(tok_d, off_d, q_d, 3)[0], L.topk(L.maxsim_scores(tok_d, off_d, q_d), 3))
    # long runs of EMPTY documents (every one a -inf key, none bounded by the tile size): mid-corpus and trailing
    for nq in (1, 3, 9):
        lens = [3] * 10 + [0] * 3000 + [5] * 40 + [0] * 700 + [130] + [0] * 2000
        q, tok, off = _case(14, 0, 0, 0, nq, 32, lens=lens)
        tok_d, off_d, q_d = tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev)
        scores = L.maxsim_scores(tok_d, off_d, q_d)
        for k in (40, 100, 128):
            assert torch.equal(L.search(tok_d, off_d, q_d, k)[0], L.topk(scores, k)), f"empty runs nq={nq} k={k}"
            if nq == 1:
                dm = L.maxsim_scores(tok_d, off_d, q_d, path=L.PATH_TC_DM)
                assert torch.equal(L.search(tok_d, off_d, q_d, k, path=L.PATH_TC_DM)[0], L.topk(dm, k)), f"dm empty runs k={k}"


@pytest.mark.parametrize("n_cand,lq,nq", [(50, 32, 1), (50, 32, 5), (1024, 32, 2), (1025, 32, 2), (7, 40, 3), (1, 32, 1), (300, 9, 4)])
def test_one_launch_rerank_equals_staged_rerank(cuda_dev, n_cand, lq, nq):
    """hrc_rerank: candidate MaxSim + last-CTA ranking in ONE launch (lq <= 32, n_cand <= 1024) against the staged
    score -> radix top-k -> unpack, bit for bit; 1025 candidates and lq = 40 take the staged route inside hrc_rerank."""
    L = _lib()
    q, tok, off = _case(21, 5000, 1, 200, nq, lq)
    tok_d, off_d, q_d = tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev)
    g = torch.Generator().manual_seed(n_cand)
    cand = torch.randint(0, 5000, (nq, n_cand), generator=g, dtype=torch.int32)
    if n_cand > 3:
        cand[0, 1] = -1
        cand[nq - 1, 2] = 5000
    cand_d = cand.to(cuda_dev)
    launches = L.launch_count()
    for k in sorted({1, min(10, n_cand), n_cand}):
        pos, ids, sc, cs = L.rerank(tok_d, off_d, cand_d, q_d, k, want_cand_scores=True)
        cs2 = L.maxsim_scores_ids(tok_d, off_d, cand_d, q_d)
        assert torch.equal(cs, cs2)
        p2, s2 = L.keys_unpack(L.topk(cs2, k))
        assert torch.equal(pos, p2) and torch.equal(sc, s2), f"k={k}"
        assert torch.equal(ids, torch.gather(cand_d, 1, pos.long()))
    if n_cand <= 1024 and lq <= 32:
        before = L.launch_count()
        L.rerank(tok_d, off_d, cand_d, q_d, 1)
        assert L.launch_count() - before == 1, "fused rerank must be a single kernel launch"


def test_results_are_bitwise_repeatable(cuda_dev):
    """A stand-in for racecheck (compute-sanitizer is closed on this GPU pool, profiles/r02_sanitizer_closed_on_this_pool.txt):
    a shared-memory or barrier race shows up as run-to-run differences, so every kernel organisation is run 25 times on
    the same inputs and must reproduce its first result bit for bit."""
    L = _lib()
    for nq, lq in ((1, 32), (2, 32), (4, 17), (8, 32), (24, 32), (3, 70)):
        q, tok, off = _case(31 + nq, 4000, 1, 300, nq, lq)
        tok_d, off_d, q_d = tok.to(cuda_dev), off.to(cuda_dev), q.to(cuda_dev)
        cand = torch.randint(0, 4000, (nq, 64), dtype=torch.int32).to(cuda_dev)
        first = None
        for _ in range(25):
            res = [L.maxsim_scores(tok_d, off_d, q_d), L.maxsim_scores_ids(tok_d, off_d, cand, q_d)]
            res += list(L.rerank(tok_d, off_d, cand, q_d, 10)[:3])
            if lq <= 32:
                res += [L.search(tok_d, off_d, q_d, 100)[0], L.maxsim_scores(tok_d, off_d, q_d, path=L.PATH_TC),
                        L.maxsim_scores(tok_d, off_d, q_d, path=L.PATH_TC_DM), L.search(tok_d, off_d, q_d, 100, path=L.PATH_TC_DM)[0]]
            if first is None:
                first = [r.clone() for r in res]
            else:
                for a, b in zip(first, res):
                    assert torch.equal(a, b), f"nq={nq} lq={lq}: results differ between runs"


def test_entry_points_are_cuda_graph_capturable(cuda_dev):
    """The device-pointer entry points neither allocate nor synchronise, so a caller can capture them in a CUDA graph
     and replay it on new query contents written in place: search, one-launch rerank and the
    whole hybrid pipeline, replayed 3 times each, must equal the direct calls bit for bit."""
    L = _lib()
    q, tok, off = _case(55, 30_000, 8, 200, 5, 32)
    tok_d, off_d = tok.to(cuda_dev), off.to(cuda_dev)
    g = torch.Generator().manual_seed(2)
    cands = torch.randint(0, 30_000, (5, 50), generator=g, dtype=torch.int32).to(cuda_dev)
    bm25s = torch.randint(0, 30_000, (5, 100), generator=g, dtype=torch.int32).to(cuda_dev)
    q_all = q.to(cuda_dev)
    for nq in (1, 2):
        q_in = q_all[:nq].clone()                                   # static inputs of the graph
        cand_in, bm_in = cands[:nq].clone(), bm25s[:nq].clone()
        ws = [L.Workspace() for _ in range(3)]

        def calls():
            keys, ids, sc = L.search(tok_d, off_d, q_in, 100, workspace=ws[0])
            pos, rid, rsc, _ = L.rerank(tok_d, off_d, cand_in, q_in, 10, workspace=ws[1])
            hid, hsc = L.hybrid_retrieve(tok_d, off_d, q_in, bm_in, colbert_k=100, rrf_k=60, n_candidates=50, final_k=10,
                                         workspace=ws[2])
            return [keys, ids, sc, pos, rid, rsc, hid, hsc]

        side = torch.cuda.Stream(cuda_dev)
        with torch.cuda.stream(side):
            calls()                                                  # warm-up: descriptors encoded, scratch grown
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                outs = calls()
        torch.cuda.synchronize()
        for shift in (0, 1, 3):
            q_in.copy_(q_all[shift:shift + nq])
            cand_in.copy_(cands[shift:shift + nq])
            bm_in.copy_(bm25s[shift:shift + nq])
            torch.cuda.synchronize()
            graph.replay()
            torch.cuda.synchronize()
            got = [t.clone() for t in outs]
            want = calls()
            torch.cuda.synchronize()
            for a, b in zip(got, want):
                assert torch.equal(a, b), f"nq={nq} shift={shift}: graph replay differs from the direct call"


def test_graphed_rerank_equals_rerank_ids(cuda_dev):
    """hrc.GraphedRerank: the fixed-shape, CUDA-graph form of rerank_ids returns what rerank_ids returns for every new
    input written in place, and refuses to run against a replaced store."""
    import hybrid_rag_colbertv2_b200 as hrc
    q, tok, off = _case(56, 20_000, 8, 300, 6, 32)
    r = hrc.JinaColBERTRetriever(hrc.RAGConfig(device=str(cuda_dev)))
    r.store = hrc.PackedStore(tok.to(cuda_dev), off.to(cuda_dev))
    g = torch.Generator().manual_seed(8)
    cands = torch.randint(-1, 20_000, (6, 50), generator=g, dtype=torch.int32).to(cuda_dev)
    q_d = q.to(cuda_dev)
    for nq in (1, 3):
        plan = hrc.GraphedRerank(r, nq, 50, k=10)
        for shift in range(3):
            plan.queries.copy_(q_d[shift:shift + nq])
            plan.candidates.copy_(cands[shift:shift + nq])
            got = [t.clone() for t in plan.run()]
            want = r.rerank_ids(q_d[shift:shift + nq], cands[shift:shift + nq], 10)
            for a, b in zip(got, want):
                assert torch.equal(a, b), f"nq={nq} shift={shift}"
    r.store = hrc.PackedStore(tok.to(cuda_dev), off.to(cuda_dev))
    with pytest.raises(RuntimeError, match="store was replaced"):
        plan.run()


def test_sharded_entry_points_on_a_one_rank_communicator(cuda_dev):
    """The multi-GPU entry points (csrc/comm.cu) inside the single-GPU suite: a ONE-rank NCCL communicator on this GPU
    drives hrc_comm_*, hrc_sharded_search[_host] and hrc_sharded_hybrid_retrieve over every transport — NCCL all-gather
    + merge kernel, peer-memory push + merge kernels (5 queries), and the exchange fused into the search's final
    selection kernel (1 query) — and must return exactly what the unsharded calls return.  (N = 2, 4, 8 over NVLink:
    scripts/check_sharded_nccl.py, bench.py's parity_check.)"""
    import socket

    import torch.distributed as dist

    import hybrid_rag_colbertv2_b200 as hrc
    L = _lib()
    if dist.is_initialized():
        pytest.skip("a process group already exists in this process")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    try:
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=cuda_dev)
    except Exception as exc:  # noqa: BLE001  (torch's own rendezvous, not the library under test)
        pytest.skip(f"cannot create a one-rank torch.distributed group here: {exc!r}")
    try:
        q, tok, off = _case(61, 40_000, 8, 200, 5, 32)
        cfg = hrc.RAGConfig(device=str(cuda_dev), colbert_top_k=100, rerank_candidates=50, final_top_k=10)
        r = hrc.JinaColBERTRetriever(cfg)
        r.store = hrc.PackedStore(tok.to(cuda_dev), off.to(cuda_dev), doc_id_base=1000)      # a shard with a non-zero base
        idx = hrc.DualIndexer(cfg)
        idx.colbert_retriever = r
        h = hrc.HybridRetriever(cfg, idx, None, verbose=False)
        q_d = q.to(cuda_dev)
        g = torch.Generator().manual_seed(3)
        bm25 = torch.randint(1000, 41_000, (5, 100), generator=g, dtype=torch.int32).to(cuda_dev)
        want_1, want_5 = r.search_keys(q_d[:1], 100), r.search_keys(q_d, 100)
        want_ids, want_sc = r.search_embeddings(q_d, 100)
        hy_ids, hy_sc = h.retrieve_batch(q_d, bm25)
        hy1_ids, hy1_sc = h.retrieve_batch(q_d[:1], bm25[:1])
        for transport in ("nccl", "p2p", "auto", "torch"):
            s = hrc.ShardedSearcher(r, transport=transport)
            if transport == "auto":
                assert s.transport == "p2p"
            l0 = L.launch_count()
            assert torch.equal(s.search_keys(q_d[:1], 100), want_1), transport
            if s.transport == "p2p":
                assert L.launch_count() - l0 == 2, "one query over peer memory: the exchange rides in the search's 2 launches"
            assert torch.equal(s.search_keys(q_d, 100), want_5), transport
            ids, sc = s.search_embeddings(q_d, 100)
            assert torch.equal(ids, want_ids) and torch.equal(sc, want_sc), transport
            hi, hs = s.search_host(q.float(), 100)
            assert torch.equal(hi, want_ids.cpu()) and torch.equal(hs, want_sc.cpu()), transport
            a_ids, a_sc = s.retrieve_batch(q_d, bm25)
            assert torch.equal(a_ids, hy_ids) and torch.equal(a_sc, hy_sc), transport
            b_ids, b_sc = s.retrieve_batch(q_d[:1], bm25[:1])
            assert torch.equal(b_ids, hy1_ids) and torch.equal(b_sc, hy1_sc), transport
            if s.transport != "torch":
                pend = [s.search_keys_async(q_d[i:i + 1], 100) for i in range(5)]
                mid = s.search_keys(q_d[2:3], 100)                                  # issued while exchanges are in flight
                assert torch.equal(mid, r.search_keys(q_d[2:3], 100)), transport
                assert all(torch.equal(p.result(), r.search_keys(q_d[i:i + 1], 100)) for i, p in enumerate(pend)), transport
            s.close()
    finally:
        dist.destroy_process_group()


def test_two_host_threads_on_two_streams_share_one_retriever(cuda_dev):
    """SURVEY §8(b) 'thread-safe per (device, stream)': two host threads drive ONE retriever on their own CUDA streams at
    the same time (search, rerank, hybrid pipeline); each must get exactly what a single-threaded run returns.  Scratch is
    per stream (_lib.Workspace), the TMA-descriptor cache is locked, the one-launch rerank's counter lives in the
    caller's workspace."""
    import threading

    import hybrid_rag_colbertv2_b200 as hrc
    q, tok, off = _case(77, 60_000, 8, 200, 6, 32)
    cfg = hrc.RAGConfig(device=str(cuda_dev), colbert_top_k=100, rerank_candidates=50, final_top_k=10)
    r = hrc.JinaColBERTRetriever(cfg)
    r.store = hrc.PackedStore(tok.to(cuda_dev), off.to(cuda_dev))
    idx = hrc.DualIndexer(cfg)
    idx.colbert_retriever = r
    h = hrc.HybridRetriever(cfg, idx, None, verbose=False)
    q_d = q.to(cuda_dev)
    g = torch.Generator().manual_seed(1)
    cand = torch.randint(0, 60_000, (6, 50), generator=g, dtype=torch.int32).to(cuda_dev)
    bm25 = torch.randint(0, 60_000, (6, 100), generator=g, dtype=torch.int32).to(cuda_dev)

    def work(i):
        qq = q_d[i:i + 1] if i % 2 == 0 else q_d[i:i + 3]          # fused doc-major search / batched kernel + streaming top-k
        n = qq.shape[0]
        return [r.search_keys(qq, 100), *r.rerank_ids(qq, cand[i:i + n], 10), *h.retrieve_batch(qq, bm25[i:i + n])]

    expect = [[t.clone() for t in work(i)] for i in range(4)]
    torch.cuda.synchronize()
    errors = []

    def run(tid):
        try:
            torch.cuda.set_device(cuda_dev)
            with torch.cuda.stream(torch.cuda.Stream(cuda_dev)):
                for it in range(60):
                    i = (it + 2 * tid) % 4
                    got = work(i)
                    torch.cuda.current_stream().synchronize()
                    for a, b in zip(got, expect[i]):
                        if not torch.equal(a, b):
                            errors.append(f"thread {tid}, iteration {it}, case {i}: result differs")
                            return
        except Exception as e:  # noqa: BLE001
            errors.append(f"thread {tid}: {e!r}")

    threads = [threading.Thread(target=run, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(r._workspace.bufs) >= 3, "expected one scratch buffer per stream"


# ======================================================================================================
# SURVEY §8(f) rows 2 and 3 on the device: streamed native store, encoder hook on CUDA
# ======================================================================================================
def test_native_store_streams_to_disk_and_back_by_shard(cuda_dev, tmp_path):
    """f2: PackedStore.save -> load(rank, world) for world in {1, 3}: the token file streams through two pinned staging
    buffers in both directions (several chunks, the last one partial; never a whole-shard host copy), each rank reads
    only its byte range, and a search over the loaded shards equals the search over the in-memory store bit for bit."""
    import hybrid_rag_colbertv2_b200 as hrc
    from hybrid_rag_colbertv2_b200.synth import plant, synth_queries, synth_store
    L = _lib()
    store = synth_store(60_000, 8, 120, seed=9, device=cuda_dev)            # ~3.8M tokens, ~1 GB
    q = synth_queries(3, 32, device=cuda_dev)
    plant(store, q, n_planted=30)
    path = str(tmp_path / "native")
    store.save(path, chunk_bytes=96 << 20)
    assert os.path.getsize(os.path.join(path, "tokens.bf16.bin")) == store.total_tokens * 256
    write_gbs = store.total_tokens * 256 / store.last_io_seconds / 1e9
    back = hrc.PackedStore.load(path, device=cuda_dev, chunk_bytes=80 << 20)
    read_gbs = back.total_tokens * 256 / back.last_io_seconds / 1e9
    print(f"native store: {store.total_tokens * 256 / 1e9:.2f} GB, save {write_gbs:.2f} GB/s, load {read_gbs:.2f} GB/s")
    assert torch.equal(back.tokens, store.tokens) and torch.equal(back.offsets, store.offsets) and back.doc_id_base == 0
    one = hrc.JinaColBERTRetriever(hrc.RAGConfig())
    one.store = store
    ref = one.search_keys(q, 100)
    for world in (1, 3):
        gathered = []
        for r in range(world):
            ld = hrc.PackedStore.load(path, device=cuda_dev, rank=r, world_size=world, chunk_bytes=50 << 20)
            sh = store.shard(r, world)
            assert ld.doc_id_base == sh.doc_id_base and torch.equal(ld.tokens, sh.tokens) and torch.equal(ld.offsets, sh.offsets)
            rr = hrc.JinaColBERTRetriever(hrc.RAGConfig())
            rr.store = ld
            gathered.append(rr.search_keys(q, 100))
        assert torch.equal(L.topk_merge(torch.cat(gathered, 1).contiguous(), 100), ref), f"world={world}"
    with pytest.raises(Exception):                                            # a truncated token file is refused, not mis-read
        with open(os.path.join(path, "tokens.bf16.bin"), "r+b") as f:
            f.truncate(store.total_tokens * 256 - 4096)
        hrc.PackedStore.load(path, device=cuda_dev)


def test_store_validate_catches_broken_offsets_and_non_finite_values(cuda_dev, tmp_path):
    """hrc_store_validate (run by PackedStore.load / JinaColBERTRetriever.load on every file they read): the CSR contract
    and the finiteness of the token values are checked on the device against a host restatement."""
    import hybrid_rag_colbertv2_b200 as hrc
    L = _lib()
    q, tok, off = _case(91, 5000, 0, 300, 1, 32)
    tok_d, off_d = tok.to(cuda_dev), off.to(cuda_dev)
    rep = L.store_validate(tok_d, off_d)
    assert rep["longest_doc_tokens"] == int((off[1:] - off[:-1]).max())
    assert L.store_validate(torch.zeros((0, 128), dtype=torch.bfloat16, device=cuda_dev),
                            torch.zeros(1, dtype=torch.int64, device=cuda_dev))["longest_doc_tokens"] == 0
    def first_bad(o, total):                                          # host restatement of the CSR contract
        n = o.numel() - 1
        for i in range(n + 1):
            v = int(o[i])
            ok = 0 <= v <= total and (i != 0 or v == 0) and (i != n or v == total) and (i == n or int(o[i + 1]) >= v)
            if not ok:
                return i
        return -1

    assert first_bad(off, int(off[-1])) == -1
    for entry, value in ((0, 1), (5000, int(off[-1]) - 1), (1234, int(off[1235]) + 1), (77, -3), (4000, int(off[-1]) + 5)):
        bad = off.clone()
        bad[entry] = value
        first = first_bad(bad, int(off[-1]))
        assert first >= 0
        with pytest.raises(ValueError, match=f"first at entry {first}\\b"):
            L.store_validate(tok_d, bad.to(cuda_dev), check_values=False)
    for row, col, v in ((0, 0, float("nan")), (int(off[-1]) - 1, 127, float("inf")), (31_337, 64, float("-inf"))):
        poisoned = tok_d.clone()
        poisoned[row, col] = v
        with pytest.raises(ValueError, match="1 token values are NaN or infinite"):
            L.store_validate(poisoned, off_d)
        L.store_validate(poisoned, off_d, check_values=False)          # offsets alone are sound
    # a native store whose token file was damaged after it was written is refused at load time
    st = hrc.PackedStore(tok_d, off_d)
    st.save(str(tmp_path / "store"))
    with open(tmp_path / "store" / "tokens.bf16.bin", "r+b") as f:
        f.seek(256 * 4321 + 10)
        f.write(b"\x80\x7f")                                          # bf16 +inf
    with pytest.raises(ValueError, match="NaN or infinite"):
        hrc.PackedStore.load(str(tmp_path / "store"), device=cuda_dev, allow_empty=True)


class _WordTokenizer:
    """Minimal tokenizer with the transformers call signature (no vocabulary files exist offline)."""
    pad_token_id, cls_token_id, sep_token_id, mask_token_id = 0, 1, 2, 3

    def _ids(self, text, add_special_tokens=True):
        words = [4 + (int.from_bytes(w.encode(), "little") % 900)
                 for w in text.lower().replace(",", " , ").replace(".", " . ").split()]
        return ([self.cls_token_id] + words + [self.sep_token_id]) if add_special_tokens else words

    def __call__(self, text, add_special_tokens=True, truncation=False, max_length=None, padding=False, return_tensors=None):
        single = isinstance(text, str)
        rows = [self._ids(t, add_special_tokens) for t in ([text] if single else text)]
        if truncation and max_length:
            rows = [r[:max_length] for r in rows]
        if return_tensors is None:
            return {"input_ids": rows[0] if single else rows}
        width = max(len(r) for r in rows)
        ids = torch.tensor([r + [self.pad_token_id] * (width - len(r)) for r in rows])
        mask = torch.tensor([[1] * len(r) + [0] * (width - len(r)) for r in rows])
        return {"input_ids": ids, "attention_mask": mask}


def test_colbert_encoder_on_cuda_through_index_search_rerank(cuda_dev, tmp_path):
    """f3: the ColBERT-style encoder hook (transformer backbone + 128-d projection + L2 norm; real weights do not exist
    offline, so a small randomly initialised backbone) runs ON THE GPU, its ragged output goes device-to-device into the
    packed store (no host round trip), and index() -> search() -> rerank() -> retrieve() agree with the oracle computed
    from the same embeddings."""
    import hybrid_rag_colbertv2_b200 as hrc
    from transformers import BertConfig, BertModel
    torch.manual_seed(0)
    backbone = BertModel(BertConfig(vocab_size=1000, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                                    intermediate_size=128, max_position_embeddings=600), add_pooling_layer=False)
    enc = hrc.ColBERTEncoder(backbone, _WordTokenizer(), projection=torch.nn.Linear(64, 128, bias=False), device=cuda_dev,
                             batch_size=16, doc_maxlen=48)
    words = "late interaction retrieval scores every query token against every document token and keeps the best".split()
    rng = np.random.default_rng(3)
    corpus = [" ".join(rng.choice(words, size=int(rng.integers(3, 40)))) + "." for _ in range(150)]
    cfg = hrc.RAGConfig(colbert_index_path=str(tmp_path / "ix"), colbert_top_k=40, bm25_top_k=40, rerank_candidates=20, final_top_k=5)
    r = hrc.JinaColBERTRetriever(cfg, encoder=enc)
    docs = enc.encode(corpus, convert_to_tensor=True)
    assert all(d.is_cuda and d.shape[1] == 128 for d in docs) and len({d.shape[0] for d in docs}) > 5      # ragged, on the GPU
    r.index(corpus)
    assert r.store.tokens.is_cuda and r.store.n_docs == 150 and r.store.total_tokens == sum(d.shape[0] for d in docs)
    assert torch.equal(r.store.tokens, torch.cat(docs).to(torch.bfloat16))
    query = "which query token keeps the best score"
    qe = enc.encode(query, convert_to_tensor=True)
    assert qe.is_cuda and qe.shape == (32, 128)
    exp = o.maxsim_scores(o.round_bf16(qe.float().cpu()), r.store.tokens.float().cpu(), r.store.offsets.cpu())[0]
    res = r.search(query=query, k=10)
    assert o.check_ranking([x["document_id"] for x in res], [x["score"] for x in res], exp, 10, RTOL) is None
    assert all(x["text"] == corpus[x["document_id"]] for x in res)
    picks = [7, 140, 33, 3, 99, 58]
    rr = r.rerank(query=query, documents=[corpus[i] for i in picks], k=4)          # re-encodes, like the reference (:783)
    assert o.check_ranking([x["result_index"] for x in rr], [x["score"] for x in rr], exp[picks], 4, RTOL) is None
    back = hrc.JinaColBERTRetriever(cfg, encoder=enc)
    back.load()
    assert torch.equal(back.store.tokens, r.store.tokens) and back.corpus == corpus
    idx = hrc.DualIndexer(cfg, encoder=enc)
    idx.colbert_retriever = r
    bm = rng.permutation(150)[:40].tolist()
    h = hrc.HybridRetriever(cfg, idx, None, verbose=False,
                            bm25_search=lambda qq, k: [{"chunk_id": int(i), "score": 1.0, "source": "bm25"} for i in bm[:k]])
    out = h.retrieve(query)
    col = h._colbert_search(query, 40)
    fused = o.rrf_reference([{"chunk_id": int(i)} for i in bm], col)
    cand = [f["chunk_id"] for f in fused[:20]]
    assert o.check_ranking([cand.index(x["chunk_id"]) for x in out], [x["score"] for x in out], exp[cand], 5, RTOL) is None
