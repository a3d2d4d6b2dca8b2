#!/usr/bin/env python
"""Experiments on libhrc_exp.so (built with -DHRC_EXPERIMENTS; the product library has none of these hooks):
 * shared-memory ring depth sweep of the single-query kernel (C2 and ragged), fused and staged routes
 * what the fused top-k costs the batched kernel: never-append (debug 8), staged kernel with the fused kernel's smem (16)
 * the TMA-ring read peak: the single-query kernel with MMA and epilogue math switched off (debug 4 | 1)
Run:  HRC_LIB_PATH=hybrid-rag-colbertv2_b200/libhrc_exp.so python scripts/exp_sweep.py
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HRC_LIB_PATH", os.path.join(ROOT, "hybrid-rag-colbertv2_b200", "libhrc_exp.so"))

import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

lib = L.load()
lib.hrc_exp_set_debug.argtypes = [ctypes.c_int]
lib.hrc_exp_set_stages.argtypes = [ctypes.c_int]
dev = torch.device("cuda:0")
K = 100


def kernel_ms(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    L.trace_enable(4 * steps + 4)
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    k = L.trace_collect()
    L.trace_enable(0)
    return sum(k) / steps


which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["stages", "tkcost", "ring"]
c2 = synth_store(1_000_000, 128, 128, seed=20260102, device=dev)
q1 = synth_queries(1, 32, device=dev)
ws = L.Workspace()
out1 = torch.empty((1, c2.n_docs), dtype=torch.float32, device=dev)
if "stages" in which:
    for rounds in range(2):
        for st in (6, 5, 4, 3, 2):
            lib.hrc_exp_set_stages(st)
            row = {"config": "C2 1 query", "stages_cap": st,
                   "staged_kernel_ms": round(kernel_ms(lambda: L.maxsim_scores(c2.tokens, c2.offsets, q1, out=out1), 40), 3),
                   "fused_kernel_ms": round(kernel_ms(lambda: L.search(c2.tokens, c2.offsets, q1, K, workspace=ws), 40), 3)}
            print(json.dumps(row), flush=True)
    lib.hrc_exp_set_stages(0)
if "ring" in which:
    base = kernel_ms(lambda: L.maxsim_scores(c2.tokens, c2.offsets, q1, out=out1), 30)
    res = {"config": "C2 1 query kernel skeletons", "full_ms": round(base, 3)}
    for bits, name in ((5, "tma_ring_only_ms"), (1, "no_epilogue_math_ms"), (4, "no_mma_ms")):
        lib.hrc_exp_set_debug(bits)
        res[name] = round(kernel_ms(lambda: L.maxsim_scores(c2.tokens, c2.offsets, q1, out=out1), 30), 3)
    lib.hrc_exp_set_debug(0)
    res["tma_ring_read_GBps"] = round(c2.total_tokens * 256 / (res["tma_ring_only_ms"] * 1e-3) / 1e9, 1)
    print(json.dumps(res), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "r02_tma_ring_read.json"), "w") as f:
        json.dump({"what": "maxsim_tc_kernel<1,1,1> with MMA and epilogue math off (libhrc_exp.so, debug 5): the TMA ring streaming "
                           "the 32.8 GB corpus alone", "ms": res["tma_ring_only_ms"], "GBps": res["tma_ring_read_GBps"]}, f)
del c2
torch.cuda.empty_cache()
if "tkcost" in which:
    rag = synth_store(1_000_000, 32, 512, seed=20260103, device=dev)
    q64 = synth_queries(64, 32, device=dev)
    out64 = torch.empty((64, rag.n_docs), dtype=torch.float32, device=dev)
    for rounds in range(2):
        row = {"config": "C3 64 queries"}
        for bits, name in ((0, "staged_ms"), (16, "staged_with_tk_smem_ms")):
            lib.hrc_exp_set_debug(bits)
            row[name] = round(kernel_ms(lambda: L.maxsim_scores(rag.tokens, rag.offsets, q64, out=out64), 4), 2)
        for bits, name in ((0, "fused_ms"), (8, "fused_never_append_ms")):
            lib.hrc_exp_set_debug(bits)
            row[name] = round(kernel_ms(lambda: L.search(rag.tokens, rag.offsets, q64, K, workspace=ws), 4), 2)
        lib.hrc_exp_set_debug(0)
        print(json.dumps(row), flush=True)
    if "stages" in which:
        for st in (6, 5, 4, 3):
            lib.hrc_exp_set_stages(st)
            print(json.dumps({"config": "ragged 1 query", "stages_cap": st,
                              "staged_kernel_ms": round(kernel_ms(lambda: L.maxsim_scores(rag.tokens, rag.offsets, q1), 20), 3),
                              "fused_kernel_ms": round(kernel_ms(lambda: L.search(rag.tokens, rag.offsets, q1, K, workspace=ws), 20), 3)}), flush=True)
        lib.hrc_exp_set_stages(0)
