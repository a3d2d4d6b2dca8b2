#!/usr/bin/env python
"""A corpus sharded over the GPUs of one box, one process per GPU: every rank encodes and stores ITS contiguous range of
documents, a search is one call per rank (local MaxSim + top-k, exchange of k keys over NVLink, merge) and every rank
gets the same answer as one GPU holding everything.  No model weights needed (SyntheticEncoder).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 \
        examples/sharded_demo.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HRC_ENCODER", "synthetic")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import hybrid_rag_colbertv2_b200 as hrc  # noqa: E402

TOPICS = ["late interaction retrieval scores every query token against every document token",
          "reciprocal rank fusion merges lexical and semantic rankings",
          "tensor memory holds the accumulators of the fifth generation tensor cores",
          "the tensor memory accelerator streams tiles into shared memory",
          "bm25 is a lexical ranking function based on term frequency",
          "high bandwidth memory feeds the streaming multiprocessors"]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    corpus = [f"{TOPICS[i % len(TOPICS)]} note {i} about item {i * 7 % 13}" for i in range(6000)]
    lo, hi = rank * len(corpus) // world, (rank + 1) * len(corpus) // world

    cfg = hrc.RAGConfig(device=str(dev), colbert_top_k=50, rerank_candidates=20, final_top_k=5)
    shard = hrc.JinaColBERTRetriever(cfg)
    emb = shard.model.encode(corpus[lo:hi])                       # this rank's documents only
    shard.index_embeddings(emb, corpus=corpus[lo:hi])
    shard.store.doc_id_base = lo                                    # ids in the results are GLOBAL corpus indices
    searcher = hrc.ShardedSearcher(shard)                           # transport="auto": peer memory, else NCCL

    query = "how does reciprocal rank fusion merge rankings"
    q = shard.model.encode(query)
    ids, scores = searcher.search_embeddings(q, k=5)
    if rank == 0:
        print(f"{world} GPUs, {len(corpus)} documents, transport {searcher.transport}")
        for i, s in zip(ids[0].tolist(), scores[0].tolist()):
            print(f"  doc {i:5d}  score {s:.4f}  {corpus[i][:70]}")

    # the same answer as one GPU holding the whole corpus
    whole = hrc.JinaColBERTRetriever(cfg)
    whole.index_embeddings(whole.model.encode(corpus), corpus=corpus)
    ref_ids, ref_scores = whole.search_embeddings(q, k=5)
    same = torch.equal(ids, ref_ids) and torch.equal(scores, ref_scores)
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("identical to the single-GPU search on every rank:", bool(flag[0]))
    searcher.close()
    dist.destroy_process_group()
    sys.exit(0 if bool(flag[0]) else 1)


if __name__ == "__main__":
    main()
