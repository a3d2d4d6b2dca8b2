#!/usr/bin/env python
"""Small, fixed workloads for ncu captures (scripts/gpu_profile_r02.sh):
  qm    C2 corpus, one query, the QUERY-major single-query kernel (HRC_PATH_TC) with fused top-k, 6 launches
  c3    64 queries x 150k ragged documents on the batched CTA-pair kernel, 6 launches
  topk  streaming top-k of a 256 x 1M score matrix, 4 launches"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1]
if what == "qm":
    st = synth_store(1_000_000, 128, 128, seed=20260102, device=dev)
    q = synth_queries(1, 32, device=dev)
    ws = L.Workspace()
    for _ in range(6):
        L.search(st.tokens, st.offsets, q, 100, workspace=ws, path=L.PATH_TC)
elif what == "c3":
    st = synth_store(150_000, 32, 512, seed=20260103, device=dev)
    q = synth_queries(64, 32, device=dev)
    out = torch.empty((64, st.n_docs), dtype=torch.float32, device=dev)
    for _ in range(6):
        L.maxsim_scores(st.tokens, st.offsets, q, out=out)
elif what == "topk":
    s = torch.randn((256, 1_000_000), device=dev)
    ws = torch.empty(max(L.topk_workspace_bytes(1_000_000, 256, 100), 1), dtype=torch.uint8, device=dev)
    for _ in range(4):
        L.topk(s, 100, workspace=ws)
torch.cuda.synchronize()
print("done", what)
