#!/usr/bin/env python
"""Sustained behaviour of the C2 kernel: per-launch time over several seconds next to nvidia-smi samples."""
import os, subprocess, sys, threading, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hybrid_rag_colbertv2_b200 import _lib
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store

dev = torch.device("cuda:0")
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 800
gap_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
store = synth_store(1_000_000, 128, 128, seed=20260102, device=dev)
q = synth_queries(8, 32, device=dev)
scores = torch.empty((1, store.n_docs), dtype=torch.float32, device=dev)
Q = "clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.active"
samples = []
p = subprocess.Popen(["nvidia-smi", "--id=0", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                     stdout=subprocess.PIPE, text=True)
def rd():
    for l in p.stdout: samples.append((time.perf_counter(), l.strip()))
threading.Thread(target=rd, daemon=True).start()
for _ in range(3): _lib.maxsim_scores(store.tokens, store.offsets, q[0:1], out=scores)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_iter + 1)]
t0 = time.perf_counter()
ev[0].record()
for i in range(n_iter):
    _lib.maxsim_scores(store.tokens, store.offsets, q[i % 8:i % 8 + 1], out=scores)
    ev[i + 1].record()
    if gap_ms: torch.cuda.synchronize(); time.sleep(gap_ms / 1e3)
torch.cuda.synchronize()
t1 = time.perf_counter()
p.terminate()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n_iter)]
step = max(1, n_iter // 40)
print("wall", round(t1 - t0, 2), "s; kernel ms every", step, ":", [round(sum(ms[i:i + step]) / len(ms[i:i + step]), 2) for i in range(0, n_iter, step)])
tail = [s for t, s in samples if t - t0 > (t1 - t0) * 0.5]
sm = sorted(float(x.split(",")[0]) for x in tail)
pw = sorted(float(x.split(",")[2]) for x in tail)
print("second half: median sm MHz", sm[len(sm) // 2] if sm else None, "median W", pw[len(pw) // 2] if pw else None, "reasons", sorted(set(x.split(",")[-1].strip() for x in tail)))
