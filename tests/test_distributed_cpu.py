"""world_size-2 gloo test (CPU) of the sharded-search host logic: token-balanced document ranges,
global-id keys, all-gather of k keys per rank and the merge invariant (merged == single-shard answer).
The per-shard scores and the merge come from the oracle here; the CUDA merge is covered by -m gpu."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import maxsim_oracle as o


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hybrid_rag_colbertv2_b200.sharded import all_gather_keys
        from hybrid_rag_colbertv2_b200.store import PackedStore, lengths_to_offsets
        g = torch.Generator().manual_seed(5)
        lens = torch.randint(1, 30, (97,), generator=g)
        off = lengths_to_offsets(lens)
        tok = o.round_bf16(torch.nn.functional.normalize(torch.randn((int(off[-1]), 128), generator=g), dim=-1))
        q = o.round_bf16(torch.nn.functional.normalize(torch.randn((3, 32, 128), generator=g), dim=-1))
        full = PackedStore.from_packed(tok, off, device="cpu")
        shard = full.shard(rank, world)
        local_scores = o.maxsim_scores(q, shard.tokens.float(), shard.offsets)
        ids = np.arange(shard.n_docs) + shard.doc_id_base                      # GLOBAL ids
        keys = o.make_keys(local_scores.numpy(), np.broadcast_to(ids, local_scores.shape))
        local_top = o.merge_keys(keys, min(k, shard.n_docs))
        gathered = all_gather_keys(torch.from_numpy(local_top.view(np.int64).copy()), k)
        assert gathered.shape == (3, world * k)
        merged = o.merge_keys(gathered.numpy().view(np.uint64), k)
        np.save(os.path.join(out_dir, f"merged_{rank}.npy"), merged)
        if rank == 0:
            ref_scores = o.maxsim_scores(q, tok, off).numpy()
            ref = o.merge_keys(o.make_keys(ref_scores, np.broadcast_to(np.arange(97), ref_scores.shape)), k)
            np.save(os.path.join(out_dir, "ref.npy"), ref)
    finally:
        dist.destroy_process_group()


def test_sharded_merge_equals_single_shard_gloo(tmp_path):
    world, k = 2, 60          # k > documents of one shard: padding with empty keys is exercised
    mp.spawn(_worker, args=(world, _free_port(), k, str(tmp_path)), nprocs=world, join=True)
    ref = np.load(tmp_path / "ref.npy")
    for r in range(world):
        assert (np.load(tmp_path / f"merged_{r}.npy") == ref).all()
    ids, scores = o.unpack_keys(ref[0])
    assert len(set(ids.tolist())) == k and (np.diff(scores) <= 0).all()
