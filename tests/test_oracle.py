"""CPU tests: the oracle against the committed golden fixtures (made from the real reference by
tests/golden/make_golden.py) and against an independent pure-loop implementation."""
import json
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import maxsim_oracle as o


def test_literal_reference_bit_equal_to_reference_outputs(golden_dir):
    z = np.load(os.path.join(golden_dir, "literal_maxsim.npz"))
    q, qb, D = (torch.from_numpy(z[k]) for k in ("q", "qb", "D"))
    for out, got in (("out_q_D", o.literal_reference(q, D)), ("out_qb_D", o.literal_reference(qb, D)),
                     ("out_q_D2d", o.literal_reference(q, D[0])), ("out_q_D1", o.literal_reference(q, D[:1]))):
        exp = torch.from_numpy(z[out])
        assert got.shape == exp.shape, out
        assert torch.equal(got, exp), out
    assert o.literal_reference(q, D[0]).dim() == 0          # 2-D docs are ONE document


def test_literal_reference_on_bf16_representable_fixture(golden_dir):
    """literal_bf16.npz: inputs exactly representable in bf16 (the packed store loses nothing) and the
    reference's own search()/rerank() results on them — the fixture that pins the GPU reference_literal path."""
    z = np.load(os.path.join(golden_dir, "literal_bf16.npz"))
    q, qb, D = (torch.from_numpy(z[k]) for k in ("q", "qb", "D"))
    for t in (q, qb, D):
        assert torch.equal(o.round_bf16(t), t)
    s = o.literal_reference(q, D)
    assert torch.equal(s, torch.from_numpy(z["out_q_D"]))
    assert torch.equal(o.literal_reference(qb, D), torch.from_numpy(z["out_qb_D"]))
    res = o.search_reference(s, None, 10)
    assert [r["document_id"] for r in res] == z["search_ids"].tolist()
    assert [r["score"] for r in res] == z["search_scores"].tolist()
    cand = z["rerank_cand"].tolist()
    rr = o.rerank_reference(o.literal_reference(q, D[cand]), [f"d{i}" for i in cand], 5)
    assert [r["result_index"] for r in rr] == z["rerank_index"].tolist()
    assert [r["score"] for r in rr] == z["rerank_scores"].tolist()


def test_reference_is_not_maxsim(golden_dir):
    """what the reference computes differs from MaxSim, hence the restated oracle."""
    z = np.load(os.path.join(golden_dir, "literal_maxsim.npz"))
    q, D = torch.from_numpy(z["q"]), torch.from_numpy(z["D"])
    lit = torch.from_numpy(z["out_q_D"])
    true = o.maxsim_dense(q, D)
    assert true.shape == lit.shape
    assert float(true.min()) > 2.0 and float(lit.abs().max()) < 1.0


def _pin_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "maxsim_pin.npz"))
    qrows, drows = torch.from_numpy(z["q_rows"]), torch.from_numpy(z["d_rows"])
    for lq in (1, 2, 4):
        for ld in (1, 2, 4):
            q = qrows[:, None, :].expand(-1, lq, -1).contiguous()
            tok = drows[:, None, :].expand(-1, ld, -1).reshape(-1, 128).contiguous()
            off = torch.arange(0, drows.shape[0] * ld + 1, ld, dtype=torch.int64)
            yield lq, ld, q, tok, off, torch.from_numpy(z[f"out_lq{lq}_ld{ld}"])


def test_maxsim_oracle_pinned_by_the_reference_where_meanpool_equals_maxsim(golden_dir):
    """maxsim_pin.npz holds outputs of the UNMODIFIED reference `_maxsim_score` (:821-829) on shapes where its
    mean-pool cosine IS MaxSim (identical query tokens, identical document tokens, exactly unit-norm dyadic rows):
    reference == (1/Lq) * sum_i max_t <q_i, d_t>, bit for bit.  This pins the oracle's dot product, max over document
    tokens and sum over query tokens to the real reference — packed, dense and pure-loop forms."""
    n_cases = 0
    for lq, ld, q, tok, off, ref in _pin_cases(golden_dir):
        assert torch.equal(o.round_bf16(q), q) and torch.equal(o.round_bf16(tok), tok)      # bf16 store loses nothing
        assert torch.equal(o.maxsim_scores(q, tok, off) / lq, ref), (lq, ld)
        assert torch.equal(o.maxsim_dense(q, tok.view(-1, ld, 128)) / lq, ref), (lq, ld)
        assert torch.equal(o.literal_reference(q, tok.view(-1, ld, 128)), ref), (lq, ld)
        if lq * ld <= 2:
            naive = o.maxsim_naive(q.numpy(), tok.numpy(), off.numpy())
            assert (naive / lq == ref.numpy()).all(), (lq, ld)
        n_cases += 1
    assert n_cases == 9 and float(ref.max()) == 1.0 and float(ref.min()) == -1.0


def test_maxsim_oracle_matches_float64_known_answers(golden_dir):
    """maxsim_kat_f64.npz: numpy float64 loop results (tests/golden/make_kat.py), independent of torch — ragged case A
    and case B with lengths around the 32-column chunk and the 128-token tile."""
    z = np.load(os.path.join(golden_dir, "maxsim_kat_f64.npz"))
    for name in ("a", "b"):
        q, tok, off = (torch.from_numpy(z[f"{name}_{k}"]) for k in ("q", "tok", "off"))
        exp = z[f"{name}_scores_f64"]
        got = o.maxsim_scores(q, tok, off).double().numpy()
        assert np.abs(got - exp).max() <= 2e-6 * np.abs(exp).max(), name
        naive = o.maxsim_naive(q.numpy(), tok.numpy(), off.numpy()).astype(np.float64)
        assert np.abs(naive - exp).max() <= 2e-6 * np.abs(exp).max(), name


def test_maxsim_oracle_matches_vllm_outputs(golden_dir):
    """maxsim_vllm_pin.npz: scores produced by vLLM 0.22.0's own ColBERT scoring functions (compute_maxsim_score and
    compute_maxsim_score_batched; tests/golden/make_vllm_pin.py) — an implementation of MaxSim this repo did not write.
    The oracle must reproduce them on ragged documents around every chunk / tile boundary, for 32-, 7- and 1-token queries."""
    from golden.make_vllm_pin import load_pin
    z = load_pin(os.path.join(golden_dir, "maxsim_vllm_pin.npz"))
    assert z["vllm_version"] == "0.22.0"
    tok, off = torch.from_numpy(z["tok"]), torch.from_numpy(z["off"])
    for name in ("q32", "q7", "q1"):
        q, pair, batched = z[name]
        got = o.maxsim_scores(torch.from_numpy(q), tok, off).numpy()
        for exp in (pair, batched):
            assert np.abs(got - exp).max() <= 2e-6 * np.abs(exp).max(), name
        if name != "q32":                                   # the triple loop is slow: the 7- and 1-token queries only
            naive = o.maxsim_naive(q, z["tok"], z["off"])
            assert np.abs(naive - pair).max() <= 2e-6 * np.abs(pair).max(), name


def test_rrf_reference_matches_reference_outputs(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "rrf.json")))
    assert len(cases) >= 8
    for c in cases:
        ids, scores = o.rrf_ids(c["a"], c["b"], c["k"])
        assert ids == c["ids"]
        assert [repr(s) for s in scores] == c["scores"]      # fp64 bit-exact (repr round-trips)
    known = cases[0]
    assert known["ids"] == [5, 3, 7, 9]
    assert known["scores"] == ['0.032266458495966696', '0.03225806451612903', '0.01639344262295082',
                               '0.015873015873015872']


def _fake_rows(text, n):
    seed = 0
    for ch in text:
        seed = (seed * 131 + ord(ch)) % (2**31 - 1)
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn((n, 128), generator=g), dim=-1)


def test_search_and_rerank_shapes_match_reference(golden_dir):
    api = json.load(open(os.path.join(golden_dir, "api_shapes.json")))
    corpus = [f"doc {i}" for i in range(30)]
    emb = torch.stack([_fake_rows(t, 16) for t in corpus])
    q = _fake_rows("hello", 32)
    s = o.literal_reference(q, emb)
    got = o.search_reference(s, corpus, 5)
    assert [g["document_id"] for g in got] == [e["document_id"] for e in api["search_k5"]]
    assert [g["score"] for g in got] == [e["score"] for e in api["search_k5"]]
    assert [sorted(g) for g in got] == [sorted(e) for e in api["search_k5"]]
    assert len(o.search_reference(s, corpus, 1000)) == api["search_k1000_len"] == 30
    s12 = o.literal_reference(q, emb[:12])
    rr = o.rerank_reference(s12, corpus[:12], 4)
    assert rr == api["rerank_k4"]
    assert len(o.rerank_reference(o.literal_reference(q, emb[:3]), corpus[:3], 10)) == api["rerank_k_gt_n_len"] == 3
    assert api["index_pt_keys"] == ["corpus", "embeddings"]
    assert api["search_n1"] == "TypeError" and api["maxsim_1d"] == "IndexError"


def _rand_case(seed, n_docs, min_len, max_len, bq, lq):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min_len, max_len + 1, (n_docs,), generator=g)
    off = torch.zeros(n_docs + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(lens, 0)
    tok = o.round_bf16(torch.nn.functional.normalize(torch.randn((int(off[-1]), 128), generator=g), dim=-1))
    q = o.round_bf16(torch.nn.functional.normalize(torch.randn((bq, lq, 128), generator=g), dim=-1))
    return q, tok, off, lens


def test_packed_oracle_matches_naive_loops():
    q, tok, off, _ = _rand_case(1, 7, 1, 9, 2, 5)
    a = o.maxsim_scores(q, tok, off).numpy()
    b = o.maxsim_naive(q.numpy(), tok.numpy(), off.numpy())
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)


def test_packed_oracle_matches_dense_einsum_with_mask():
    q, tok, off, lens = _rand_case(2, 40, 1, 33, 3, 32)
    D = torch.zeros((40, 33, 128))
    for i in range(40):
        D[i, :lens[i]] = tok[off[i]:off[i + 1]]
    a = o.maxsim_scores(q, tok, off)
    b = o.maxsim_dense(q, D, lens)
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
    # chunking must not change anything
    c = o.maxsim_scores(q, tok, off, doc_chunk_tokens=37)
    assert torch.equal(a, c)


def test_empty_document_scores_minus_inf():
    q, tok, _, _ = _rand_case(3, 3, 2, 2, 1, 4)
    off = torch.tensor([0, 2, 2, 4, 6])
    s = o.maxsim_scores(q, tok, off)
    assert s.shape == (1, 4) and s[0, 1] == float("-inf") and torch.isfinite(s[0, [0, 2, 3]]).all()


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), n_docs=st.integers(1, 12), max_len=st.integers(1, 20), lq=st.integers(1, 32))
def test_oracle_property_vs_naive(seed, n_docs, max_len, lq):
    q, tok, off, _ = _rand_case(seed, n_docs, 1, max_len, 1, lq)
    a = o.maxsim_scores(q, tok, off).numpy()
    qn, tn = q.numpy(), tok.numpy()
    for i in range(n_docs):                      # vectorised per-document restatement
        sim = qn[0] @ tn[off[i]:off[i + 1]].T
        assert abs(a[0, i] - sim.max(1).sum()) < 1e-4


def test_keys_roundtrip_and_order():
    s = np.array([1.5, -2.0, 0.0, 7.25, 7.25, -np.inf, np.inf], dtype=np.float32)
    ids = np.arange(7)
    keys = o.make_keys(s, ids)
    gi, gs = o.unpack_keys(keys)
    assert (gi == ids).all() and (gs == s).all()
    order = np.argsort(keys)[::-1]
    assert list(order) == [6, 3, 4, 0, 2, 1, 5]             # score desc, ties -> lower id first
    m = o.merge_keys(keys[None, :], 3)
    assert list(o.unpack_keys(m[0])[0]) == [6, 3, 4]
    assert o.unpack_keys(np.zeros(2, dtype=np.uint64))[0].tolist() == [-1, -1]


def test_check_ranking_accepts_ties_and_rejects_wrong():
    s = torch.tensor([1.0, 5.0, 5.0, 3.0, 2.0])
    assert o.check_ranking([1, 2, 3], [5.0, 5.0, 3.0], s, 3) is None
    assert o.check_ranking([2, 1, 3], [5.0, 5.0, 3.0], s, 3) is None      # tie may swap
    assert o.check_ranking([1, 3, 2], [5.0, 3.0, 5.0], s, 3) is not None  # not descending
    assert o.check_ranking([1, 2, 4], [5.0, 5.0, 2.0], s, 3) is not None  # wrong member
    assert o.check_ranking([1, 2, 3], [5.0, 5.0, 3.2], s, 3) is not None  # wrong score
