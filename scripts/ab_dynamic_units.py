#!/usr/bin/env python
"""In-process A/B of the doc-major kernel's work distribution: equal token ranges per CTA (no workspace -> no claim
counter -> static) against own + shared units claimed at run time (256-byte workspace).  Scores-only route
(hrc_maxsim_scores, HRC_PATH_TC_DM), alternating, at burst (20 launches after an idle second) and sustained (3 s back
to back, second half); kernel time from hrc_trace.

    python scripts/ab_dynamic_units.py [c2|ragged|long]
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

dev = torch.device("cuda:0")
lib = L.load()
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
store = (synth_store(1_000_000, 128, 128, seed=20260102, device=dev) if which == "c2" else
         synth_store(32_000, 3000, 5000, seed=20260104, device=dev) if which == "long" else      # 4,000-token documents
         synth_store(1_000_000, 32, 512, seed=20260103, device=dev))
q = synth_queries(1, 32, device=dev)
out = torch.empty((1, store.n_docs), dtype=torch.float32, device=dev)
ws = torch.zeros(256, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream


def launch(dynamic):
    rc = lib.hrc_maxsim_scores(store.tokens.data_ptr(), store.offsets.data_ptr(), store.n_docs, store.total_tokens,
                               q.data_ptr(), 1, 32, out.data_ptr(), L.PATH_TC_DM, ws.data_ptr() if dynamic else None,
                               256 if dynamic else 0, stream)
    assert rc == 0, lib.hrc_last_error()


launch(False)
ref = out.clone()
launch(True)
assert torch.equal(out, ref), "static and dynamic distributions must give identical scores"


def run(dynamic, seconds):
    for _ in range(3):
        launch(dynamic)
    torch.cuda.synchronize()
    L.trace_enable(4000)
    t0 = time.perf_counter()
    while True:
        for _ in range(20):
            launch(dynamic)
        torch.cuda.synchronize()
        if time.perf_counter() - t0 >= seconds:
            break
    k = L.trace_collect()
    L.trace_enable(0)
    half = k[len(k) // 2:] if seconds > 0.5 else k
    return {"kernel_ms_mean": round(sum(half) / len(half), 3), "kernel_ms_median": round(statistics.median(half), 3)}


for rnd in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    for dynamic in (False, True):
        time.sleep(1.0)
        burst = run(dynamic, 0.0)
        sus = run(dynamic, 3.0)
        print(json.dumps({"corpus": which, "distribution": "dynamic" if dynamic else "static", "round": rnd, "burst": burst,
                          "sustained": sus}), flush=True)
