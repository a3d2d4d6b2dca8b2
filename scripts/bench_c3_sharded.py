#!/usr/bin/env python
"""C3 at N GPUs (run under torchrun): every rank holds `--docs-per-gpu` ragged documents (32..512 tokens) of one global
corpus, 256 queries x 32 tokens are scored against all of them (CTA-pair tensor-core kernel), per-query local top-100,
all-gather of 256 x 100 keys per rank, on-device merge.  Prints one JSON line from rank 0 (CUDA events, max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
        scripts/bench_c3_sharded.py [--docs-per-gpu 1000000] [--queries 256] [--steps 3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import hybrid_rag_colbertv2_b200 as hrc  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    store = synth_store(a.docs_per_gpu * world, 32, 512, seed=12, device=dev, rank=rank, world_size=world)
    r = hrc.JinaColBERTRetriever(hrc.RAGConfig(device=str(dev)), encoder=hrc.SyntheticEncoder())
    r.store = store
    s = hrc.ShardedSearcher(r)
    q = synth_queries(a.queries, 32, device=dev)
    for _ in range(2):
        keys = s.search_keys(q, 100)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        keys = s.search_keys(q, 100)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps, float(store.total_tokens)], device=dev, dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms, tokens = float(tmax[0]), float(tsum[1])
        flops = 2.0 * 32 * 128 * a.queries * tokens
        ids, sc = hrc._lib.keys_unpack(keys)
        print(json.dumps({"config": f"C3 sharded: {a.queries} queries x 32 over {a.docs_per_gpu * world} docs x U(32..512) "
                                    f"on {world} GPUs", "n_gpus": world, "ms_per_step": ms, "tokens": tokens,
                          "useful_TFLOPs_total": flops / (ms * 1e-3) / 1e12,
                          "useful_TFLOPs_per_gpu": flops / (ms * 1e-3) / 1e12 / world,
                          "pairs_per_s": a.queries * a.docs_per_gpu * world / (ms * 1e-3),
                          "top1_query0": [int(ids[0, 0]), float(sc[0, 0])]}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
