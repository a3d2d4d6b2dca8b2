#!/usr/bin/env python
"""libhrc_exp.so experiment: how much of the corpus should the doc-major kernel hand out dynamically, in how many units?
(share, per_cta) = 1/share of the tokens in per_cta shared units per CTA; (0, 0) = the static distribution.
Scores-only doc-major route, configurations interleaved over 3 rounds; burst = 20 launches after an idle half second,
sustained = 2 s back to back (second half); kernel time from hrc_trace.   python scripts/exp_dyn_sweep.py [c2|ragged]"""
import ctypes
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HRC_LIB_PATH", os.path.join(ROOT, "hybrid-rag-colbertv2_b200", "libhrc_exp.so"))
import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

dev = torch.device("cuda:0")
lib = L.load()
lib.hrc_exp_set_dyn.argtypes = [ctypes.c_int, ctypes.c_int]
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
store = synth_store(1_000_000, 128, 128, seed=20260102, device=dev) if which == "c2" else synth_store(1_000_000, 32, 512, seed=20260103, device=dev)
q = synth_queries(1, 32, device=dev)
out = torch.empty((1, store.n_docs), dtype=torch.float32, device=dev)
ws = torch.zeros(256, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream


def launch(dynamic):
    rc = lib.hrc_maxsim_scores(store.tokens.data_ptr(), store.offsets.data_ptr(), store.n_docs, store.total_tokens,
                               q.data_ptr(), 1, 32, out.data_ptr(), L.PATH_TC_DM, ws.data_ptr() if dynamic else None,
                               256 if dynamic else 0, stream)
    assert rc == 0, lib.hrc_last_error()


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
    return round(statistics.median(half), 3), round(sum(half) / len(half), 3)


configs = [(0, 0), (4, 16), (4, 32), (8, 16), (8, 32), (2, 32), (16, 16), (16, 8)]
res = {c: {"burst": [], "sustained": []} for c in configs}
for rnd in range(3):
    for c in configs:
        lib.hrc_exp_set_dyn(*c) if c != (0, 0) else lib.hrc_exp_set_dyn(0, 0)
        time.sleep(0.5)
        res[c]["burst"].append(run(c != (0, 0), 0.0)[0])
        res[c]["sustained"].append(run(c != (0, 0), 2.0)[1])
lib.hrc_exp_set_dyn(0, 0)
for c in configs:
    print(json.dumps({"corpus": which, "share": c[0], "units_per_cta": c[1], "burst_kernel_ms_median": res[c]["burst"],
                      "sustained_kernel_ms_mean": res[c]["sustained"]}), flush=True)
