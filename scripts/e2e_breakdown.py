#!/usr/bin/env python
"""Where does the end-to-end step (host query in -> host ids out) spend its time beyond the kernel?
Per step: CPU time to enqueue hrc_search_host, CPU wait in the stream synchronisation, device time of the scoring
kernel (hrc_trace), device time between the first and the last operation of the step (CUDA events)."""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import hybrid_rag_colbertv2_b200 as hrc  # noqa: E402
from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

dev = torch.device("cuda:0")
store = synth_store(1_000_000, 128, 128, seed=20260102, device=dev)
r = hrc.JinaColBERTRetriever(hrc.RAGConfig())
r.store = store
qh = [synth_queries(8, 32)[i:i + 1].float().pin_memory() for i in range(8)]
qd = synth_queries(8, 32, device=dev)
for i in range(5):
    r.search_host(qh[i], 100)
steps = 40
L.trace_enable(steps + 2)
wall, dev_ms = [], []
for i in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    r.search_host(qh[i % 8], 100)
    e1.record()
    e1.synchronize()
    wall.append((time.perf_counter() - t0) * 1e3)
    dev_ms.append(e0.elapsed_time(e1))
kern = L.trace_collect()
print(json.dumps({"mode": "search_host per step", "wall_ms": statistics.median(wall), "event_ms": statistics.median(dev_ms),
                  "kernel_ms": statistics.median(kern), "wall_p90": sorted(wall)[int(.9 * steps)]}))
# device-resident steps, each followed by a synchronisation (same idle gaps, no copies)
L.trace_enable(steps + 2)
wall = []
for i in range(steps):
    t0 = time.perf_counter()
    r.search_keys(qd[i % 8:i % 8 + 1], 100)
    torch.cuda.synchronize()
    wall.append((time.perf_counter() - t0) * 1e3)
kern2 = L.trace_collect()
print(json.dumps({"mode": "search_keys + synchronize per step", "wall_ms": statistics.median(wall), "kernel_ms": statistics.median(kern2)}))
# back to back, no per-step synchronisation
L.trace_enable(steps + 2)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(steps):
    r.search_keys(qd[i % 8:i % 8 + 1], 100)
torch.cuda.synchronize()
print(json.dumps({"mode": "search_keys back to back", "wall_ms": (time.perf_counter() - t0) * 1e3 / steps,
                  "kernel_ms": statistics.median(L.trace_collect())}))
