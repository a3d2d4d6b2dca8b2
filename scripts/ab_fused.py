#!/usr/bin/env python
"""In-process A/B on ONE box: hrc_search's fused route (top-k in the MaxSim epilogue) against the staged route
(score matrix -> radix top-k), alternating, for C2, the ragged corpus with one query, and C3.  Kernel times from
hrc_trace, step times from CUDA events.  One JSON line per (config, route)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

dev = torch.device("cuda:0")
K = 100


def run(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    L.trace_enable(4 * steps + 4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    kern = L.trace_collect()
    L.trace_enable(0)
    return e0.elapsed_time(e1) / steps, sum(kern) / steps


def ab(name, store, q, steps, rounds=3):
    ws = L.Workspace()
    scores = torch.empty((q.shape[0], store.n_docs), dtype=torch.float32, device=dev)
    tws = torch.empty(max(L.topk_workspace_bytes(store.n_docs, q.shape[0], K), 1), dtype=torch.uint8, device=dev)

    def fused():
        L.search(store.tokens, store.offsets, q, K, workspace=ws, unpack=True)

    def staged():
        L.maxsim_scores(store.tokens, store.offsets, q, out=scores)
        L.keys_unpack(L.topk(scores, K, workspace=tws))

    def fused_dm():
        L.search(store.tokens, store.offsets, q, K, workspace=ws, unpack=True, path=L.PATH_TC_DM)

    routes = [("fused", fused), ("staged", staged)] + ([("fused_doc_major", fused_dm)] if q.shape[0] == 1 else [])
    res = {name: [] for name, _ in routes}
    for _ in range(rounds):
        for route, fn in routes:
            res[route].append(run(fn, steps))
    for route, xs in res.items():
        print(json.dumps({"config": name, "route": route, "step_ms": [round(a, 3) for a, _ in xs],
                          "kernel_ms": [round(b, 3) for _, b in xs]}), flush=True)


which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2", "ragged", "c3"]
if "c2" in which:
    st = synth_store(1_000_000, 128, 128, seed=20260102, device=dev)
    ab("C2 1 query", st, synth_queries(1, 32, device=dev), 60)
    ab("C2 2 queries", st, synth_queries(2, 32, device=dev), 40)
    ab("C2 4 queries", st, synth_queries(4, 32, device=dev), 40)
    del st
    torch.cuda.empty_cache()
if "ragged" in which or "c3" in which:
    st = synth_store(1_000_000, 32, 512, seed=20260103, device=dev)
    if "ragged" in which:
        ab("ragged 1 query", st, synth_queries(1, 32, device=dev), 30)
    if "c3" in which:
        ab("C3 8 queries", st, synth_queries(8, 32, device=dev), 10)
        ab("C3 64 queries", st, synth_queries(64, 32, device=dev), 4)
        ab("C3 256 queries", st, synth_queries(256, 32, device=dev), 3, rounds=2)
