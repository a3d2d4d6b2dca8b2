#!/usr/bin/env python
"""Where does a C2 search step spend its time?  CUDA events between the stages of search_keys."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hybrid_rag_colbertv2_b200 as hrc
from hybrid_rag_colbertv2_b200 import _lib
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store

dev = torch.device("cuda:0")
store = synth_store(1_000_000, 128, 128, seed=20260102, device=dev)
q = synth_queries(8, 32, device=dev)
r = hrc.JinaColBERTRetriever(hrc.RAGConfig(device="cuda:0")); r.store = store
scores = torch.empty((1, store.n_docs), dtype=torch.float32, device=dev)
ws = torch.empty(_lib.topk_workspace_bytes(store.n_docs, 1, 100), dtype=torch.uint8, device=dev)
N = 20
def ev(): return torch.cuda.Event(enable_timing=True)
for mode in ("staged", "search_keys", "maxsim_only", "topk_only"):
    for _ in range(3): r.search_keys(q[0:1], 100)
    torch.cuda.synchronize()
    e = [ev() for _ in range(3 * N + 1)]
    e[0].record()
    for i in range(N):
        if mode == "staged":
            _lib.maxsim_scores(store.tokens, store.offsets, q[i % 8:i % 8 + 1], out=scores); e[3 * i + 1].record()
            _lib.topk(scores, 100, workspace=ws); e[3 * i + 2].record(); e[3 * i + 3].record()
        elif mode == "search_keys":
            r.search_keys(q[i % 8:i % 8 + 1], 100); e[3 * i + 1].record(); e[3 * i + 2].record(); e[3 * i + 3].record()
        elif mode == "maxsim_only":
            _lib.maxsim_scores(store.tokens, store.offsets, q[i % 8:i % 8 + 1], out=scores); e[3 * i + 1].record(); e[3 * i + 2].record(); e[3 * i + 3].record()
        else:
            _lib.topk(scores, 100, workspace=ws); e[3 * i + 1].record(); e[3 * i + 2].record(); e[3 * i + 3].record()
    torch.cuda.synchronize()
    a = sum(e[3 * i].elapsed_time(e[3 * i + 1]) for i in range(N)) / N
    b = sum(e[3 * i + 1].elapsed_time(e[3 * i + 2]) for i in range(N)) / N
    tot = e[0].elapsed_time(e[3 * N]) / N
    print(json.dumps({"mode": mode, "first_stage_ms": a, "second_stage_ms": b, "per_step_ms": tot}))
