#!/usr/bin/env python
"""Multi-GPU parity over real NCCL (run under torchrun, one rank per GPU):
the document-sharded search (local top-k -> all-gather of k keys -> on-device merge) must return
bit-identical keys to a single-GPU search over the whole corpus, for single and batched queries.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/check_sharded_nccl.py [--docs 400000]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import hybrid_rag_colbertv2_b200 as hrc  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import plant, synth_queries, synth_store  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=400_000)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    q = synth_queries(19, 32, device=dev)                      # 19: CTA pairs + an odd query group
    # this rank's shard of the global ragged corpus, planted deterministically by GLOBAL doc id
    shard = synth_store(args.docs, 32, 300, seed=77, device=dev, rank=rank, world_size=world)
    plant(shard, q[:3], n_planted=50, n_docs_global=args.docs)
    r = hrc.JinaColBERTRetriever(hrc.RAGConfig(device=str(dev)))
    r.store = shard
    s = hrc.ShardedSearcher(r)
    keys_1 = s.search_keys(q[:1], 100)
    keys_b = s.search_keys(q, 100)
    ok = True
    if rank == 0:
        full = synth_store(args.docs, 32, 300, seed=77, device=dev)
        plant(full, q[:3], n_planted=50)
        one = hrc.JinaColBERTRetriever(hrc.RAGConfig(device=str(dev)))
        one.store = full
        ref_1, ref_b = one.search_keys(q[:1], 100), one.search_keys(q, 100)
        ok = bool(torch.equal(ref_1, keys_1)) and bool(torch.equal(ref_b, keys_b))
        print(f"world={world} docs={args.docs}: sharded == single-GPU keys: {ok} "
              f"(top id {int(hrc._lib.keys_unpack(keys_1)[0][0, 0])})", flush=True)
    # every rank holds the same merged list
    gathered = [torch.empty_like(keys_b) for _ in range(world)]
    dist.all_gather(gathered, keys_b)
    same = all(torch.equal(g, keys_b) for g in gathered)
    if rank == 0:
        print(f"world={world}: all ranks hold the same merged list: {same}", flush=True)
    dist.destroy_process_group()
    if not (ok and same):
        sys.exit(1)


if __name__ == "__main__":
    main()
