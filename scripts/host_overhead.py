#!/usr/bin/env python
"""C1 anatomy: where the ~22 us of one rerank call go.  Times (a) back-to-back `_lib.rerank` / `rerank_ids` calls
(what bench.py's `secondary.c1` reports), (b) the same launch replayed from a CUDA graph (device-only floor: no Python,
no ctypes, no launch set-up), (c) the host-side pieces of the wrapper one by one.  One GPU."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hybrid_rag_colbertv2_b200 as hrc  # noqa: E402
from hybrid_rag_colbertv2_b200 import _lib  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
store = synth_store(20_000, 32, 512, seed=1, device=dev)
q = synth_queries(4, 32, device=dev)
cand = torch.randint(0, 20_000, (1, 50), dtype=torch.int32, device=dev)
cfg = hrc.RAGConfig(device=str(dev))
r = hrc.JinaColBERTRetriever(cfg)
r.store = store
ws = _lib.Workspace()
N = 2000


def dev_us(fn, n=N):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def host_us(fn, n=20000):
    for _ in range(100):
        fn()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t) / n * 1e6


print(f"_lib.rerank back to back        {dev_us(lambda: _lib.rerank(store.tokens, store.offsets, cand, q[:1], 10, workspace=ws)):7.2f} us/call")
print(f"rerank_ids back to back         {dev_us(lambda: r.rerank_ids(q[:1], cand, 10)):7.2f} us/call")

# host time of one call when the GPU is never the bottleneck: issue, then wait, per call
def one():
    _lib.rerank(store.tokens, store.offsets, cand, q[:1], 10, workspace=ws)
t = time.perf_counter()
for _ in range(N):
    one()
host_issue = (time.perf_counter() - t) / N * 1e6
torch.cuda.synchronize()
print(f"_lib.rerank host issue rate     {host_issue:7.2f} us/call (wall clock, queue never drained)")

# CUDA graph of the same call: device-only floor
side = torch.cuda.Stream(dev)
with torch.cuda.stream(side):
    for _ in range(3):
        out = _lib.rerank(store.tokens, store.offsets, cand, q[:1], 10, workspace=ws)
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        out_g = _lib.rerank(store.tokens, store.offsets, cand, q[:1], 10, workspace=ws)
torch.cuda.synchronize()
ref = _lib.rerank(store.tokens, store.offsets, cand, q[:1], 10, workspace=ws)
g.replay()
torch.cuda.synchronize()
assert all(torch.equal(a, b) for a, b in zip(ref[:3], out_g[:3])), "graph replay differs"
print(f"CUDA-graph replay of hrc_rerank {dev_us(g.replay):7.2f} us/replay")

lib = _lib.load()
print("host pieces (us):")
print(f"  ctypes hrc_version()                    {host_us(lib.hrc_version):6.2f}")
print(f"  torch.cuda.current_stream().cuda_stream {host_us(lambda: torch.cuda.current_stream(dev).cuda_stream):6.2f}")
def ctx():
    with torch.cuda.device(dev):
        pass
print(f"  with torch.cuda.device(dev): pass       {host_us(ctx):6.2f}")
print(f"  torch.empty((3,1,10), int32, cuda)      {host_us(lambda: torch.empty((3, 1, 10), dtype=torch.int32, device=dev)):6.2f}")
print(f"  _require_cuda(4 tensors)                {host_us(lambda: _lib._require_cuda(store.tokens, store.offsets, cand, q)):6.2f}")
print(f"  4 x data_ptr()                          {host_us(lambda: (store.tokens.data_ptr(), store.offsets.data_ptr(), cand.data_ptr(), q.data_ptr())):6.2f}")
print(f"  r._prep_queries(q)                      {host_us(lambda: r._prep_queries(q[:1])):6.2f}")
print(f"  q[:1] slice                             {host_us(lambda: q[:1]):6.2f}")
