#!/usr/bin/env python
"""Multi-GPU parity over real NVLink (run under torchrun, one rank per GPU): for every transport of the exchange step —
"torch" (torch.distributed all_gather), "nccl" (ncclAllGather inside libhrc.so), "p2p" (peer stores + flags inside
libhrc.so) — the document-sharded search, the host-buffer search and the sharded hybrid pipeline must return results
bit-identical to a single GPU holding the whole corpus, on every rank, for single and batched queries, including
shards smaller than k, a rank with an EMPTY shard, and many back-to-back steps (the P2P double buffer).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/check_sharded_nccl.py [--docs 400000]
Also prints the device time of the exchange + merge per transport (CUDA events, max over ranks).
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
from hybrid_rag_colbertv2_b200 import _lib  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import plant, synth_queries, synth_store  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=400_000)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True

    def check(cond, what):
        nonlocal ok
        flag = torch.tensor([1 if cond else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        good = bool(flag[0])
        ok = ok and good
        if rank == 0:
            print(f"{'ok  ' if good else 'FAIL'} {what}", flush=True)

    q = synth_queries(19, 32, device=dev)                      # 19: CTA pairs + an odd query group
    cfg = hrc.RAGConfig(device=str(dev), colbert_top_k=100, rerank_candidates=50, final_top_k=10)
    # this rank's shard of the global ragged corpus, planted deterministically by GLOBAL doc id
    shard = synth_store(args.docs, 32, 300, seed=77, device=dev, rank=rank, world_size=world)
    plant(shard, q[:3], n_planted=50, n_docs_global=args.docs)
    r = hrc.JinaColBERTRetriever(cfg)
    r.store = shard
    # every rank also builds the whole corpus as the single-GPU reference (400k docs x ~166 tokens = 17 GB)
    full = synth_store(args.docs, 32, 300, seed=77, device=dev)
    plant(full, q[:3], n_planted=50)
    one = hrc.JinaColBERTRetriever(cfg)
    one.store = full
    ref_1, ref_b = one.search_keys(q[:1], 100), one.search_keys(q, 100)
    g = torch.Generator().manual_seed(3)
    bm25 = torch.randint(0, args.docs, (19, 100), generator=g, dtype=torch.int32).to(dev)
    col_ids = _lib.keys_unpack(ref_b)[0]
    bm25[:, :30] = col_ids[:, torch.randperm(100, generator=g)[:30].to(dev)]
    bm25[2, 90:] = -1
    bm25[3, 5] = args.docs + 7                                 # out of range: -inf on one GPU, -inf here
    idx = hrc.DualIndexer(cfg)
    idx.colbert_retriever = one
    h = hrc.HybridRetriever(cfg, idx, None, verbose=False)
    hy_ids, hy_sc = h.retrieve_batch(q, bm25)
    hy1_ids, hy1_sc = h.retrieve_batch(q[:1], bm25[:1])

    timings = {}
    for transport in ("torch", "nccl", "p2p"):
        s = hrc.ShardedSearcher(r, transport=transport)
        check(torch.equal(s.search_keys(q[:1], 100), ref_1), f"[{transport}] 1 query: sharded keys == single-GPU keys")
        check(torch.equal(s.search_keys(q, 100), ref_b), f"[{transport}] 19 queries: sharded keys == single-GPU keys")
        ids, sc = s.search_embeddings(q, 100)
        ri, rs = _lib.keys_unpack(ref_b)
        check(torch.equal(ids, ri) and torch.equal(sc, rs), f"[{transport}] search_embeddings (unpacked)")
        hi, hs = s.search_host(q.float().cpu(), 100)
        check(torch.equal(hi, ri.cpu()) and torch.equal(hs, rs.cpu()), f"[{transport}] search_host")
        same = True
        for step in range(40):                                  # back-to-back steps, alternating shapes
            qq = q[step % 19: step % 19 + 1]
            same = same and torch.equal(s.search_keys(qq, 100), one.search_keys(qq, 100))
        check(same, f"[{transport}] 40 back-to-back steps")
        if transport != "torch":                                # pipelined form: exchange on a side stream
            pend = [s.search_keys_async(q[i % 19: i % 19 + 1], 100) for i in range(30)]
            same = all(torch.equal(p.result(), one.search_keys(q[i % 19: i % 19 + 1], 100)) for i, p in enumerate(pend))
            check(same, f"[{transport}] 30 pipelined steps (search_keys_async)")
            # synchronous searches issued while pipelined exchanges are still in flight (no .result() in between)
            same = True
            for rnd in range(6):
                pend = [s.search_keys_async(q[(rnd + i) % 19: (rnd + i) % 19 + 1], 100) for i in range(3)]
                sync_keys = s.search_keys(q[rnd:rnd + 1], 100)
                same = same and torch.equal(sync_keys, one.search_keys(q[rnd:rnd + 1], 100))
                same = same and all(torch.equal(p.result(), one.search_keys(q[(rnd + i) % 19: (rnd + i) % 19 + 1], 100))
                                    for i, p in enumerate(pend))
            check(same, f"[{transport}] synchronous searches interleaved with pipelined ones")
        a_ids, a_sc = s.retrieve_batch(q, bm25)
        check(torch.equal(a_ids, hy_ids) and torch.equal(a_sc, hy_sc), f"[{transport}] sharded hybrid retrieve, 19 queries")
        b_ids, b_sc = s.retrieve_batch(q[:1], bm25[:1])
        check(torch.equal(b_ids, hy1_ids) and torch.equal(b_sc, hy1_sc), f"[{transport}] sharded hybrid retrieve, 1 query")
        # exchange + merge alone
        local_keys = r.search_keys(q[:1], 100).contiguous()
        if transport == "torch":
            from hybrid_rag_colbertv2_b200.sharded import all_gather_keys
            fn = lambda: _lib.topk_merge(all_gather_keys(local_keys, 100), 100)  # noqa: E731
        else:
            tr = _lib.TRANSPORT_NCCL if transport == "nccl" else _lib.TRANSPORT_P2P
            ws = _lib.Workspace()
            fn = lambda: _lib.allgather_merge_topk(s.comm, local_keys, 100, transport=tr, workspace=ws)  # noqa: E731
        for _ in range(5):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50 * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        timings[transport] = round(float(t[0]), 2)
        s.close()

    # shards smaller than k and an EMPTY shard: 30 documents over `world` ranks, k = 25
    tiny_full = synth_store(30, 5, 20, seed=5, device=dev)
    tiny = hrc.JinaColBERTRetriever(cfg)
    tiny.store = tiny_full
    if world > 2:       # leave the last rank without documents
        lo, hi = (rank * 30) // (world - 1), ((rank + 1) * 30) // (world - 1)
        lo, hi = (lo, hi) if rank < world - 1 else (30, 30)
    else:
        lo, hi = (rank * 30) // world, ((rank + 1) * 30) // world
    off = tiny_full.offsets
    t0, t1 = int(off[lo]), int(off[hi])
    part = hrc.JinaColBERTRetriever(cfg)
    part.store = hrc.PackedStore(tiny_full.tokens[t0:t1].contiguous() if t1 > t0 else torch.zeros((0, 128), dtype=torch.bfloat16, device=dev),
                                 (off[lo:hi + 1] - off[lo]).contiguous(), doc_id_base=lo)
    for transport in ("nccl", "p2p"):
        s = hrc.ShardedSearcher(part, transport=transport)
        check(torch.equal(s.search_keys(q[:2], 25), tiny.search_keys(q[:2], 25)), f"[{transport}] shards smaller than k / empty shard")
        s.close()
    # uneven shards, ONE query: rank 0 holds 100 documents (k = 25: its search's final kernel does the exchange itself),
    # every other rank 10 (fewer than k: separate push + merge kernels) — the two routes share one slot / flag protocol
    n_small = 100 + 10 * (world - 1)
    mixed_full = synth_store(n_small, 5, 20, seed=6, device=dev)
    mixed = hrc.JinaColBERTRetriever(cfg)
    mixed.store = mixed_full
    lo = 0 if rank == 0 else 100 + 10 * (rank - 1)
    hi = 100 if rank == 0 else lo + 10
    off = mixed_full.offsets
    t0, t1 = int(off[lo]), int(off[hi])
    part = hrc.JinaColBERTRetriever(cfg)
    part.store = hrc.PackedStore(mixed_full.tokens[t0:t1].contiguous(), (off[lo:hi + 1] - off[lo]).contiguous(), doc_id_base=lo)
    launches = {}
    for transport in ("nccl", "p2p"):
        s = hrc.ShardedSearcher(part, transport=transport)
        same = all(torch.equal(s.search_keys(q[i:i + 1], 25), mixed.search_keys(q[i:i + 1], 25)) for i in range(19))
        check(same, f"[{transport}] uneven shards, single queries (fused-exchange and separate routes mixed)")
        s.close()
        s = hrc.ShardedSearcher(r, transport=transport)              # launches of one sharded single-query search
        s.search_keys(q[:1], 100)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        s.search_keys(q[:1], 100)
        launches[transport] = _lib.launch_count() - l0
        s.close()
    # what the exchange costs on top of a search, seen where the search itself is short: 4,000 documents per rank
    small = hrc.JinaColBERTRetriever(cfg)
    small.store = synth_store(4000 * world, 32, 64, seed=9, device=dev, rank=rank, world_size=world)
    step_us = {}

    def time_us(fn, reps=200):
        for _ in range(10):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t[0]), 2)

    step_us["local_search"] = time_us(lambda: small.search_keys(q[:1], 100))
    for transport in ("torch", "nccl", "p2p"):
        s = hrc.ShardedSearcher(small, transport=transport)
        step_us[f"sharded_search_{transport}"] = time_us(lambda: s.search_keys(q[:1], 100))
        s.close()
    if rank == 0:
        print(json.dumps({"world": world, "small_shard_step_us": step_us}), flush=True)
        print(json.dumps({"world": world, "launches_per_single_query_search": launches}), flush=True)
        print(json.dumps({"world": world, "exchange_plus_merge_us": timings}), flush=True)
        print(f"world={world}: {'ALL OK' if ok else 'FAILED'}", flush=True)
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
