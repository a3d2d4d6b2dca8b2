#!/usr/bin/env python
"""libhrc_exp.so: what does the fused top-k cost the doc-major kernel under sustained load?  Complete search steps,
3 s back to back each, alternating: staged (scores + streaming top-k), fused, fused that never offers a key (debug 8),
fused without offers and without the final list merge (debug 8|32)."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HRC_LIB_PATH", os.path.join(ROOT, "hybrid-rag-colbertv2_b200", "libhrc_exp.so"))

import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

lib = L.load()
lib.hrc_exp_set_debug.argtypes = [ctypes.c_int]
dev = torch.device("cuda:0")
K = 100
store = synth_store(1_000_000, 128, 128, seed=20260102, device=dev)
q = synth_queries(1, 32, device=dev)
ws = L.Workspace()
scores = torch.empty((1, store.n_docs), dtype=torch.float32, device=dev)
tws = torch.empty(max(L.topk_workspace_bytes(store.n_docs, 1, K), 1), dtype=torch.uint8, device=dev)


def sustained(fn, seconds=3.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    L.trace_enable(4000)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    el = time.perf_counter() - t0
    k = L.trace_collect()
    L.trace_enable(0)
    half = k[len(k) // 2:]
    return round(el / n * 1e3, 3), round(sum(half) / max(len(half), 1), 3)


def staged(path):
    L.maxsim_scores(store.tokens, store.offsets, q, out=scores, path=path)
    L.topk(scores, K, workspace=tws)


routes = [("dm_staged", 0, lambda: staged(L.PATH_TC_DM)), ("dm_fused", 0, lambda: L.search(store.tokens, store.offsets, q, K, workspace=ws, path=L.PATH_TC_DM)),
          ("dm_fused_never_offer", 8, lambda: L.search(store.tokens, store.offsets, q, K, workspace=ws, path=L.PATH_TC_DM)),
          ("dm_fused_never_offer_no_merge", 40, lambda: L.search(store.tokens, store.offsets, q, K, workspace=ws, path=L.PATH_TC_DM)),
          ("qm_staged", 0, lambda: staged(L.PATH_TC)), ("qm_fused", 0, lambda: L.search(store.tokens, store.offsets, q, K, workspace=ws, path=L.PATH_TC))]
for rnd in range(2):
    for name, bits, fn in routes:
        lib.hrc_exp_set_debug(bits)
        step, kern = sustained(fn)
        lib.hrc_exp_set_debug(0)
        print(json.dumps({"route": name, "round": rnd, "step_ms": step, "kernel_ms": kern}), flush=True)
