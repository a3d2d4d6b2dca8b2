#!/usr/bin/env python
"""libhrc_exp.so experiment: does the single-query search need all 148 SMs?  The kernel is HBM-bound and, run back to
back, sits at the GPU's power cap, where SM clocks (and with them the on-chip fabric) drop; fewer resident CTAs draw less
SM power.  Sweeps the number of CTAs (= corpus segments) of the doc-major fused search at burst (20 steps after an idle
second) and sustained (3 s back to back, second half), alternating twice; kernel time from hrc_trace, clocks / power from NVML.

    python scripts/exp_cta_sweep.py [c2|ragged]
"""
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HRC_LIB_PATH", os.path.join(ROOT, "hybrid-rag-colbertv2_b200", "libhrc_exp.so"))

import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

import pynvml  # noqa: E402
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda:0")
lib = L.load()
lib.hrc_exp_set_ctas.argtypes = [ctypes.c_int]
K = 100
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
store = synth_store(1_000_000, 128, 128, seed=20260102, device=dev) if which == "c2" else synth_store(1_000_000, 32, 512, seed=20260103, device=dev)
q = synth_queries(4, 32, device=dev)


def run(seconds):
    ws = L.Workspace()                                   # fresh: the workspace layout depends on the segment count
    fn = lambda: L.search(store.tokens, store.offsets, q[:1], K, workspace=ws, unpack=False)  # noqa: E731
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
    while True:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
        if time.perf_counter() - t0 >= seconds:
            break
    el = time.perf_counter() - t0
    k = L.trace_collect()
    L.trace_enable(0)
    stop.set()
    th.join()
    half = k[len(k) // 2:] if seconds > 0.5 else k
    tail = samples[len(samples) // 2:] or [(0, 0.0)]
    return {"step_ms": round(el / n * 1e3, 3), "kernel_ms": round(sum(half) / max(len(half), 1), 3),
            "kernel_ms_median": round(statistics.median(half), 3), "sm_mhz": statistics.median([s[0] for s in tail]),
            "power_w": round(statistics.median([s[1] for s in tail]), 1)}


CTAS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [148, 140, 132, 120, 108, 96, 74]
ROUNDS = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ref = None
for rnd in range(ROUNDS):
    for ctas in CTAS:
        lib.hrc_exp_set_ctas(ctas)
        keys = L.search(store.tokens, store.offsets, q[:1], K, workspace=L.Workspace(), unpack=False)[0].clone()
        if ref is None:
            ref = keys
        assert torch.equal(keys, ref), f"{ctas} CTAs: result differs"
        time.sleep(1.0)
        burst = run(0.0)
        sus = run(3.0)
        print(json.dumps({"corpus": which, "ctas": ctas, "round": rnd, "burst": burst, "sustained": sus}), flush=True)
lib.hrc_exp_set_ctas(0)
