#!/usr/bin/env python
"""Where does the tensor-core path stop paying?  (HRC_PATH_AUTO's rule must come from data.)
TC vs SIMT (and the doc-major TC kernel) for query lengths 1..32 on (a) the 50-candidate rerank (launch-latency-bound)
and (b) a full-corpus single-query scan of 200k ragged passages (bandwidth-bound).  One JSON line per point."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3     # us


dev = torch.device("cuda:0")
store = synth_store(200_000, 32, 512, seed=12, device=dev)
g = torch.Generator().manual_seed(0)
cand = torch.randint(0, store.n_docs, (1, 50), generator=g, dtype=torch.int32).to(dev)
out = torch.empty((1, store.n_docs), dtype=torch.float32, device=dev)
paths = (("tc", L.PATH_TC), ("tc_dm", L.PATH_TC_DM), ("simt", L.PATH_SIMT))
for lq in (1, 2, 4, 8, 16, 32):
    q = synth_queries(1, lq, device=dev)
    row = {"lq": lq}
    for name, path in paths:
        row[f"rerank50_{name}_us"] = round(timed(lambda: L.maxsim_scores_ids(store.tokens, store.offsets, cand, q, path=path), 200), 2)
        row[f"scan200k_{name}_us"] = round(timed(lambda: L.maxsim_scores(store.tokens, store.offsets, q, path=path, out=out), 5), 1)
    print(json.dumps(row), flush=True)
for nq in (1, 2, 3, 4):                       # few-query scans: explicit query-major vs what AUTO picks
    q = synth_queries(nq, 32, device=dev)
    o2 = torch.empty((nq, store.n_docs), dtype=torch.float32, device=dev)
    print(json.dumps({"nq": nq, "scan200k_tc_us": round(timed(lambda: L.maxsim_scores(store.tokens, store.offsets, q, path=L.PATH_TC, out=o2), 10), 1),
                      "scan200k_auto_us": round(timed(lambda: L.maxsim_scores(store.tokens, store.offsets, q, path=L.PATH_AUTO, out=o2), 10), 1)}), flush=True)
