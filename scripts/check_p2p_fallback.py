#!/usr/bin/env python
"""The ranks must AGREE when the peer-memory transport cannot be set up (run under torchrun on >= 2 GPUs, against
libhrc_exp.so): `hrc_exp_fail_p2p(1)` makes hrc_comm_enable_p2p fail locally on rank 1 only; every rank must then get the
same error from the call (nobody hangs in a later collective), ShardedSearcher(transport="auto") must settle on NCCL
everywhere, and searches must keep returning the single-GPU result.

    HRC_LIB_PATH=hybrid-rag-colbertv2_b200/libhrc_exp.so python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29536 scripts/check_p2p_fallback.py
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HRC_LIB_PATH", os.path.join(ROOT, "hybrid-rag-colbertv2_b200", "libhrc_exp.so"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import hybrid_rag_colbertv2_b200 as hrc  # noqa: E402
from hybrid_rag_colbertv2_b200 import _lib  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    lib.hrc_exp_fail_p2p.argtypes = [ctypes.c_int]
    lib.hrc_exp_fail_p2p.restype = None
    cfg = hrc.RAGConfig(device=str(dev))
    q = synth_queries(3, 32, device=dev)
    shard = hrc.JinaColBERTRetriever(cfg)
    shard.store = synth_store(50_000, 32, 200, seed=4, device=dev, rank=rank, world_size=world)
    one = hrc.JinaColBERTRetriever(cfg)
    one.store = synth_store(50_000, 32, 200, seed=4, device=dev)
    ok = True

    def check(cond, what):
        nonlocal ok
        flag = torch.tensor([1 if cond else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag[0])
        if rank == 0:
            print(f"{'ok  ' if bool(flag[0]) else 'FAIL'} {what}", flush=True)

    lib.hrc_exp_fail_p2p(1)                                   # every process sets it; only rank 1 is affected
    raised = False
    try:
        hrc.ShardedSearcher(shard, transport="p2p")
    except _lib.HrcError as exc:
        raised = "unavailable" in str(exc)
    check(raised, "forced failure on rank 1: hrc_comm_enable_p2p reports 'unavailable' on EVERY rank")
    s = hrc.ShardedSearcher(shard, transport="auto")
    check(s.transport == "nccl", "transport='auto' settles on nccl on every rank")
    check(all(torch.equal(s.search_keys(q[i:i + 1], 100), one.search_keys(q[i:i + 1], 100)) for i in range(3)),
          "searches over the fallback transport == single-GPU keys")
    s.close()
    lib.hrc_exp_fail_p2p(-1)
    s = hrc.ShardedSearcher(shard, transport="auto")
    check(s.transport == "p2p", "without the forced failure transport='auto' picks p2p")
    check(all(torch.equal(s.search_keys(q[i:i + 1], 100), one.search_keys(q[i:i + 1], 100)) for i in range(3)),
          "searches over p2p (exchange fused into the search's final kernel) == single-GPU keys")
    s.close()
    if rank == 0:
        print(f"world={world}: {'ALL OK' if ok else 'FAILED'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
