#!/usr/bin/env python
"""Sustained (power-capped) A/B of the single-query search routes on ONE box: each route runs back to back for ~3 s,
alternating twice; kernel time from hrc_trace, SM clock / power from NVML.  Routes: query-major fused (default),
doc-major fused (HRC_PATH_TC_DM), each also writing scores only, and the pure read probe.  (The M=64 variant measured
here in round 2 — between the two — has been removed: profiles/r02_logs/r02_sustained_ab_*.log keep its numbers.)"""
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

dev = torch.device("cuda:0")
K = 100
import pynvml  # noqa: E402
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


def sustained(fn, seconds=3.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    samples, stop = [], threading.Event()

    def poll():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.01)
    th = threading.Thread(target=poll, daemon=True)
    th.start()
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
    stop.set()
    th.join()
    half = k[len(k) // 2:]
    return {"step_ms": round(el / n * 1e3, 3), "kernel_ms_second_half": round(sum(half) / max(len(half), 1), 3) if half else None,
            "sm_mhz": statistics.median([s[0] for s in samples[len(samples) // 2:]]),
            "power_w": round(statistics.median([s[1] for s in samples[len(samples) // 2:]]), 1)}


which = sys.argv[1] if len(sys.argv) > 1 else "c2"
store = synth_store(1_000_000, 128, 128, seed=20260102, device=dev) if which == "c2" else synth_store(1_000_000, 32, 512, seed=20260103, device=dev)
q = synth_queries(1, 32, device=dev)
ws = L.Workspace()
scores = torch.empty((1, store.n_docs), dtype=torch.float32, device=dev)
routes = {"query_major_fused": lambda: L.search(store.tokens, store.offsets, q, K, workspace=ws),
          "doc_major_fused": lambda: L.search(store.tokens, store.offsets, q, K, workspace=ws, path=L.PATH_TC_DM),
          "query_major_scores_only": lambda: L.maxsim_scores(store.tokens, store.offsets, q, out=scores, path=L.PATH_TC),
          "doc_major_scores_only": lambda: L.maxsim_scores(store.tokens, store.offsets, q, out=scores, path=L.PATH_TC_DM)}
probe_out = torch.zeros(1, dtype=torch.int32, device=dev)
routes["read_probe_16B_loads"] = lambda: L.read_probe(store.tokens, probe_out)      # the memory system alone, same bytes
for rnd in range(2):
    for name, fn in routes.items():
        print(json.dumps({"corpus": which, "route": name, "round": rnd, **sustained(fn)}), flush=True)
