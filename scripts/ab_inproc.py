#!/usr/bin/env python
"""In-process A/B of several libhrc builds on the SAME corpus: the builds take turns (round-robin, many rounds) so
box-to-box and thermal drift cancel.  Times hrc_maxsim_scores alone with CUDA events.

    python scripts/ab_inproc.py [--workload c2|ragged|c3] [--rounds 12] [--launches 10] libA.so libB.so ...
"""
import argparse
import ctypes
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402


def bind(path):
    lib = ctypes.CDLL(path)
    fn = lib.hrc_maxsim_scores
    fn.restype, fn.argtypes = _lib.SYMBOLS["hrc_maxsim_scores"]
    return fn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--rounds", type=int, default=12)
    ap.add_argument("--launches", type=int, default=10)
    ap.add_argument("--docs", type=int, default=0)
    ap.add_argument("--sleep", type=float, default=0.0, help="idle seconds before every measurement (burst clocks instead of the power cap)")
    ap.add_argument("--path", type=int, default=0, help="HRC_PATH_* selector passed to every library (0 auto, 2 query-major, 3 doc-major)")
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    if a.workload == "c2":
        store, nq = synth_store(a.docs or 1_000_000, 128, 128, seed=20260102, device=dev), 1
    elif a.workload == "ragged":
        store, nq = synth_store(a.docs or 1_000_000, 32, 512, seed=12, device=dev), 1
    else:
        store, nq = synth_store(a.docs or 300_000, 32, 512, seed=12, device=dev), 64
    q = synth_queries(nq, 32, device=dev)
    out = torch.empty((nq, store.n_docs), dtype=torch.float32, device=dev)
    fns = [bind(os.path.abspath(p)) for p in a.libs]
    envs = [None] * len(fns)                 # (the libraries read no environment variables any more: builds are the A/B axis)
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws = torch.empty(max(int(_lib.load().hrc_maxsim_workspace_bytes(store.n_docs, nq, 32)), 256), dtype=torch.uint8, device=dev)

    def launch(fn, env=None):
        rc = fn(store.tokens.data_ptr(), store.offsets.data_ptr(), store.n_docs, store.total_tokens, q.data_ptr(), nq, 32,
                out.data_ptr(), a.path, ws.data_ptr(), ws.numel(), stream)
        assert rc == 0, rc

    first = None
    for fn, env in zip(fns, envs):
        for _ in range(3):
            launch(fn, env)
        torch.cuda.synchronize()
        if first is None:
            first = out.clone()
        assert torch.equal(out, first), "the libraries disagree on the scores"
    times = [[] for _ in fns]
    for r in range(a.rounds):
        order = list(range(len(fns)))
        if r % 2:
            order.reverse()
        for i in order:
            if a.sleep > 0:
                import time
                time.sleep(a.sleep)
                launch(fns[i], envs[i])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.launches):
                launch(fns[i], envs[i])
            e1.record()
            torch.cuda.synchronize()
            times[i].append(e0.elapsed_time(e1) / a.launches)
    for p, t in zip(a.libs, times):
        med = statistics.median(t)
        extra = (f"{store.total_tokens * 256 / med / 1e6:8.0f} GB/s" if nq == 1 else
                 f"{2.0 * 32 * 128 * nq * store.total_tokens / med / 1e9:8.0f} TFLOP/s")
        print(f"{os.path.basename(p):34s} median {med:8.4f} ms  min {min(t):8.4f}  max {max(t):8.4f}  {extra}")


if __name__ == "__main__":
    main()
