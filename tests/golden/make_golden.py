"""Generate the committed golden fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py            # needs /root/reference (not present on the GPU box)

The reference script cannot be imported as shipped (5 third-party imports are absent offline,
SURVEY.md F4), so those imports are stubbed with MagicMock exactly as in SURVEY.md §A.1; nothing in
/root/reference is edited or copied.  Outputs (small, committed):
  literal_maxsim.npz  inputs + outputs of the reference's `_maxsim_score` (what it literally computes)
  literal_bf16.npz    the same on bf16-REPRESENTABLE inputs (so the packed bf16 store loses nothing) plus the
                      reference's own search() / rerank() results on them: pins the GPU "reference_literal" path
  maxsim_pin.npz      the degenerate shapes on which the reference's mean-pool cosine (:821-829) IS MaxSim — every
                      query token the same vector, every document token the same vector, all vectors exactly
                      unit-norm with dyadic coordinates — so `_maxsim_score` of the REAL reference pins the MaxSim
                      kernels' dot-product core, max over tokens and sum over query tokens bit for bit
  api_shapes.json     search()/rerank()/index()/load() observable behaviour with a fake encoder
  rrf.json            `_reciprocal_rank_fusion` ids + fp64 scores (repr round-trips) incl. tie order
  api_signatures.json the reference's method signatures (names, parameter names, defaults) and RAGConfig fields
                      for the classes on the path: what "drops into local_rag_complete.py unchanged" must match
"""
import importlib.util
import json
import os
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/local_rag_complete.py"


def load_reference():
    for name in ["pymupdf4llm", "fitz", "bm25s", "sentence_transformers", "sqlalchemy", "sqlalchemy.ext",
                 "sqlalchemy.ext.declarative", "sqlalchemy.orm"]:
        sys.modules[name] = MagicMock()
    spec = importlib.util.spec_from_file_location("local_rag_complete", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class FakeEncoder:
    """encode(str) -> [32,128]; encode(list) -> [N,16,128]; deterministic in the text."""

    def __init__(self):
        self.calls = []

    @staticmethod
    def _rows(text, n):
        seed = 0
        for ch in text:
            seed = (seed * 131 + ord(ch)) % (2**31 - 1)
        g = torch.Generator().manual_seed(seed)
        return torch.nn.functional.normalize(torch.randn((n, 128), generator=g), dim=-1)

    def encode(self, x, **kw):
        self.calls.append((type(x).__name__, sorted(kw.items())))
        if isinstance(x, str):
            return self._rows(x, 32)
        return torch.stack([self._rows(t, 16) for t in x])


def main():
    m = load_reference()
    R = m.JinaColBERTRetriever.__new__(m.JinaColBERTRetriever)   # bypass __init__ (needs the model)

    # ---- 1. literal _maxsim_score --------------------------------------------------------------
    g = torch.Generator().manual_seed(20260101)
    q = torch.nn.functional.normalize(torch.randn((32, 128), generator=g), dim=-1)
    qb = torch.nn.functional.normalize(torch.randn((4, 32, 128), generator=g), dim=-1)
    D = torch.nn.functional.normalize(torch.randn((8, 16, 128), generator=g), dim=-1)
    np.savez_compressed(
        os.path.join(HERE, "literal_maxsim.npz"),
        q=q.numpy(), qb=qb.numpy(), D=D.numpy(),
        out_q_D=R._maxsim_score(q, D).numpy(),
        out_qb_D=R._maxsim_score(qb, D).numpy(),
        out_q_D2d=R._maxsim_score(q, D[0]).numpy(),          # 2-D docs -> ONE document -> 0-d
        out_q_D1=R._maxsim_score(q, D[:1]).numpy(),          # N == 1 -> 0-d
    )

    # ---- 1b. the same on bf16-representable inputs, plus search()/rerank() through the reference ---------
    def bf16r(x):
        return x.to(torch.bfloat16).to(torch.float32)

    g = torch.Generator().manual_seed(20260105)
    q2 = bf16r(torch.nn.functional.normalize(torch.randn((32, 128), generator=g), dim=-1))
    qb2 = bf16r(torch.nn.functional.normalize(torch.randn((3, 32, 128), generator=g), dim=-1))
    D2 = bf16r(torch.nn.functional.normalize(torch.randn((64, 24, 128), generator=g) +
                                             0.6 * torch.randn((64, 1, 128), generator=g), dim=-1))

    class FixedEncoder:                 # encode(str) -> q2; encode(list of "d<i>") -> those rows of D2
        def encode(self, x, **kw):
            if isinstance(x, str):
                return q2
            return torch.stack([D2[int(t[1:])] for t in x])

    R.config = m.RAGConfig()
    R.model = FixedEncoder()
    R.corpus_embeddings = D2
    R.corpus = [f"d{i}" for i in range(64)]
    s10 = R.search("query", k=10)
    cand = [40, 3, 17, 63, 0, 22, 9, 51, 33, 12, 5, 28]
    r5 = R.rerank("query", [f"d{i}" for i in cand], k=5)
    np.savez_compressed(
        os.path.join(HERE, "literal_bf16.npz"),
        q=q2.numpy(), qb=qb2.numpy(), D=D2.numpy(),
        out_q_D=R._maxsim_score(q2, D2).numpy(),
        out_qb_D=R._maxsim_score(qb2, D2).numpy(),
        search_ids=np.array([r['document_id'] for r in s10], dtype=np.int64),
        search_scores=np.array([r['score'] for r in s10], dtype=np.float64),
        rerank_cand=np.array(cand, dtype=np.int64),
        rerank_index=np.array([r['result_index'] for r in r5], dtype=np.int64),
        rerank_scores=np.array([r['score'] for r in r5], dtype=np.float64),
    )

    # ---- 1c. where mean-pool cosine == MaxSim: the reference itself pins the MaxSim arithmetic ------------
    # Rows with coordinates in {0, +-1/2, +-1/4, +-1/8} and squared norm exactly 1 (16a + 4b + c = 64 non-zeros of
    # each size): every dot product is a multiple of 1/64, exact in fp32 under ANY summation order, the norms are
    # exactly 1, and the mean of 1, 2 or 4 identical rows is that row.  Then cos(mean q, mean d) == <q, d> ==
    # (1 / Lq) * sum_i max_t <q_i, d_t>: the reference's function equals MaxSim (mean-reduced) bit for bit.
    def dyadic_unit_rows(n, gen):
        combos = [(2, 4, 16), (0, 8, 32), (1, 6, 24), (3, 2, 8), (0, 0, 64), (0, 16, 0), (1, 8, 16), (2, 0, 32)]
        rows = torch.zeros((n, 128))
        for i in range(n):
            a, b, c = combos[int(torch.randint(0, len(combos), (1,), generator=gen))]
            pos = torch.randperm(128, generator=gen)[: a + b + c]
            mag = torch.cat([torch.full((a,), 0.5), torch.full((b,), 0.25), torch.full((c,), 0.125)])
            sign = torch.randint(0, 2, (a + b + c,), generator=gen).float() * 2 - 1
            rows[i, pos] = mag * sign
        assert torch.equal(rows.pow(2).sum(-1), torch.ones(n))
        return rows

    g = torch.Generator().manual_seed(20260107)
    pin = {}
    n_pin, bq_pin = 96, 5
    qrows = dyadic_unit_rows(bq_pin, g)
    drows = dyadic_unit_rows(n_pin, g)
    drows[:bq_pin] = qrows                       # a document identical to each query: score exactly 1
    drows[bq_pin] = -qrows[0]                    # and one exactly opposite: score exactly -1
    pin["q_rows"], pin["d_rows"] = qrows.numpy(), drows.numpy()
    exact = (qrows.double() @ drows.double().T).float()            # multiples of 1/64
    for lq in (1, 2, 4):
        for ld in (1, 2, 4):
            qq = qrows[:, None, :].expand(bq_pin, lq, 128).contiguous()
            dd = drows[:, None, :].expand(n_pin, ld, 128).contiguous()
            out = R._maxsim_score(qq, dd)                          # the unmodified reference, [Bq, N]
            assert torch.equal(out, exact), (lq, ld)               # the premise of this fixture
            pin[f"out_lq{lq}_ld{ld}"] = out.numpy()
    np.savez_compressed(os.path.join(HERE, "maxsim_pin.npz"), **pin)

    # ---- 2. API behaviour with a fake encoder ---------------------------------------------------
    api = {}
    with tempfile.TemporaryDirectory() as tmp:
        cfg = m.RAGConfig(colbert_index_path=os.path.join(tmp, "colbert"))
        R.config = cfg
        R.model = FakeEncoder()
        R.corpus_embeddings = None
        R.corpus = None
        corpus = [f"doc {i}" for i in range(30)]
        R.index(corpus)
        saved = torch.load(os.path.join(cfg.colbert_index_path, "index.pt"))
        api["index_pt_keys"] = sorted(saved.keys())
        api["index_pt_embeddings_shape"] = list(saved["embeddings"].shape)
        api["index_pt_embeddings_dtype"] = str(saved["embeddings"].dtype)
        api["config_device"] = cfg.device
        api["config_defaults"] = {"bm25_top_k": cfg.bm25_top_k, "colbert_top_k": cfg.colbert_top_k,
                                  "final_top_k": cfg.final_top_k, "colbert_index_path": m.RAGConfig().colbert_index_path,
                                  "embedding_model": cfg.embedding_model}
        s5 = R.search("hello", k=5)
        api["search_k5"] = s5
        api["search_k1000_len"] = len(R.search("hello", k=1000))
        api["search_default_len"] = len(R.search("hello"))
        rr = R.rerank("hello", corpus[:12], k=4)
        api["rerank_k4"] = rr
        api["rerank_default_len"] = len(R.rerank("hello", corpus[:12]))
        api["rerank_k_gt_n_len"] = len(R.rerank("hello", corpus[:3], k=10))
        api["encode_calls"] = R.model.calls
        R2 = m.JinaColBERTRetriever.__new__(m.JinaColBERTRetriever)
        R2.config = cfg
        R2.load()
        api["load_roundtrip_equal"] = bool(torch.equal(R2.corpus_embeddings, R.corpus_embeddings)) and R2.corpus == corpus
        # N == 1 crashes search
        R.index(corpus[:1])
        try:
            R.search("hello", k=1)
            api["search_n1"] = "ok"
        except Exception as e:  # noqa: BLE001
            api["search_n1"] = type(e).__name__
        # 1-D inputs raise
        try:
            R._maxsim_score(torch.randn(128), torch.randn(5, 128))
            api["maxsim_1d"] = "ok"
        except Exception as e:  # noqa: BLE001
            api["maxsim_1d"] = type(e).__name__
    with open(os.path.join(HERE, "api_shapes.json"), "w") as f:
        json.dump(api, f, indent=1, sort_keys=True)

    # ---- 2b. signatures of the classes on the path ----------------------------------------------
    import dataclasses
    import inspect

    def sig(fn):
        out = []
        for name, prm in inspect.signature(fn).parameters.items():
            out.append({"name": name, "kind": prm.kind.name,
                        "default": None if prm.default is inspect.Parameter.empty else repr(prm.default),
                        "has_default": prm.default is not inspect.Parameter.empty})
        return out

    sigs = {"RAGConfig": [{"name": f.name, "default": repr(f.default)} for f in dataclasses.fields(m.RAGConfig)]}
    for cls, methods in {
        "JinaColBERTRetriever": ["__init__", "index", "load", "search", "rerank", "_maxsim_score"],
        "DualIndexer": ["__init__", "build_bm25_index", "build_colbert_index", "load_indexes"],
        "HybridRetriever": ["__init__", "retrieve", "_bm25_search", "_colbert_search", "_reciprocal_rank_fusion",
                            "_fetch_chunks_from_db", "_colbert_rerank"],
    }.items():
        sigs[cls] = {name: sig(getattr(getattr(m, cls), name)) for name in methods}
    with open(os.path.join(HERE, "api_signatures.json"), "w") as f:
        json.dump(sigs, f, indent=1, sort_keys=True)

    # ---- 3. RRF ----------------------------------------------------------------------------------
    H = m.HybridRetriever.__new__(m.HybridRetriever)
    cases = []

    def run(a, b, k=None):
        ra = [{'chunk_id': int(i), 'score': 0.0, 'source': 'bm25'} for i in a]
        rb = [{'chunk_id': int(i), 'score': 0.0, 'source': 'colbert'} for i in b]
        out = H._reciprocal_rank_fusion(ra, rb) if k is None else H._reciprocal_rank_fusion(ra, rb, k=k)
        cases.append({"a": [int(i) for i in a], "b": [int(i) for i in b], "k": 60 if k is None else k,
                      "ids": [r['chunk_id'] for r in out], "scores": [repr(r['rrf_score']) for r in out]})

    run([5, 3, 9], [7, 3, 5])                      # SURVEY.md §A.3 known answer
    run([], [1, 2, 3])
    run([4, 4, 4], [4])                            # an id repeated inside one list accumulates left to right
    run([10, 11, 12, 13], [13, 12, 11, 10])        # mathematically equal sums: tie order = insertion order
    rng = np.random.default_rng(20260104)
    for overlap in (0, 30, 100):
        a = rng.permutation(100000)[:100]
        b = np.concatenate([rng.permutation(a)[:overlap], 200000 + rng.permutation(100000)[:100 - overlap]])
        b = rng.permutation(b)
        run(a, b)
    a = rng.permutation(5000)[:100]
    run(a, rng.permutation(a), k=1)               # different constant
    with open(os.path.join(HERE, "rrf.json"), "w") as f:
        json.dump(cases, f)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
