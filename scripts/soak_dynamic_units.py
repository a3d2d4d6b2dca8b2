#!/usr/bin/env python
"""Soak of the doc-major kernel's run-time work distribution: thousands of searches over corpora of several sizes and
length distributions; which CTA scores which documents differs from launch to launch, the keys must not."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

dev = torch.device("cuda:0")
q = synth_queries(8, 32, device=dev)
t0 = time.time()
total = 0
for n_docs, lo, hi, reps in ((8_000, 64, 128, 4000), (30_000, 1, 300, 3000), (200_000, 32, 512, 400), (1_000_000, 128, 128, 150),
                             (5_000, 100, 4000, 2000)):
    store = synth_store(n_docs, lo, hi, seed=n_docs, device=dev)
    ws = L.Workspace()
    for qi in range(2):
        ref_keys = L.search(store.tokens, store.offsets, q[qi:qi + 1], 100, workspace=ws, unpack=False)[0].clone()
        ref_sc = L.maxsim_scores(store.tokens, store.offsets, q[qi:qi + 1], workspace=ws).clone()
        assert torch.equal(L.topk(ref_sc, 100), ref_keys)
        bad = torch.zeros((), dtype=torch.int64, device=dev)
        for i in range(reps):
            keys = L.search(store.tokens, store.offsets, q[qi:qi + 1], 100, workspace=ws, unpack=False)[0]
            bad += (keys != ref_keys).any()
            if i % 8 == 0:
                bad += (L.maxsim_scores(store.tokens, store.offsets, q[qi:qi + 1], workspace=ws) != ref_sc).any()
        torch.cuda.synchronize()
        assert int(bad) == 0, f"{n_docs} docs: {int(bad)} launches differed"
        total += reps
    print(f"ok   {n_docs} docs x U({lo}..{hi}) tokens ({store.total_tokens} tokens): {2 * reps} searches identical", flush=True)
    del store
    torch.cuda.empty_cache()
print(f"soak ok: {total} searches in {time.time() - t0:.0f} s")
