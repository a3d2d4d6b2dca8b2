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


# ------------------------------------------------------------------------------------------------------------------
# The sharded HYBRID pipeline's host logic (ShardedSearcher.retrieve_batch, transport "torch") on two gloo ranks.
# The five libhrc device calls it makes are replaced, IN THIS TEST ONLY, by CPU stand-ins built on the oracle, so what
# is exercised is the product's sharding logic: global ColBERT list -> RRF on global ids on every rank -> every rank
# scores only the candidates it OWNS (rank 0 also answers for ids no shard holds) -> exchange of (score, position)
# keys -> merge.  Expected: the single-process pipeline over the unsharded corpus, bit for bit.
# ------------------------------------------------------------------------------------------------------------------
def _cpu_lib_standins(monkeypatch_target):
    L = monkeypatch_target

    def keys_unpack(keys):
        ids, sc = o.unpack_keys(keys.numpy().view(np.uint64))
        return torch.from_numpy(ids.copy()), torch.from_numpy(sc.copy())

    def rrf_fuse(a, b, rrf_k, top_n):
        ids = torch.full((a.shape[0], top_n), -1, dtype=torch.int32)
        sc = torch.zeros((a.shape[0], top_n), dtype=torch.float64)
        cnt = torch.zeros((a.shape[0],), dtype=torch.int32)
        for r in range(a.shape[0]):
            ri, rs = o.rrf_ids(a[r].tolist(), b[r].tolist(), rrf_k)
            n = min(len(ri), top_n)
            ids[r, :n] = torch.tensor(ri[:n], dtype=torch.int32)
            sc[r, :n] = torch.tensor(rs[:n], dtype=torch.float64)
            cnt[r] = len(ri)
        return ids, sc, cnt

    def maxsim_scores_ids(tokens, offsets, cand, q, path=0, workspace=None):
        full = o.maxsim_scores(q.float(), tokens.float(), offsets)
        n = offsets.numel() - 1
        ok = (cand >= 0) & (cand < n)
        out = torch.gather(full, 1, cand.clamp(0, max(n - 1, 0)).long()) if n else torch.zeros(cand.shape)
        return torch.where(ok, out, torch.full_like(out, float("-inf")))

    def topk(scores, k, ids=None, id_base=0, workspace=None):
        idn = ids.numpy() if ids is not None else np.broadcast_to(np.arange(scores.shape[1]) + id_base, scores.shape)
        return torch.from_numpy(o.merge_keys(o.make_keys(scores.numpy(), idn), k).view(np.int64).copy())

    def topk_merge(keys, k):
        return torch.from_numpy(o.merge_keys(keys.numpy().view(np.uint64), k).view(np.int64).copy())

    for name, fn in (("keys_unpack", keys_unpack), ("rrf_fuse", rrf_fuse), ("maxsim_scores_ids", maxsim_scores_ids),
                     ("topk", topk), ("topk_merge", topk_merge)):
        setattr(L, name, fn)


def _hybrid_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import hybrid_rag_colbertv2_b200 as hrc
        from hybrid_rag_colbertv2_b200 import _lib
        from hybrid_rag_colbertv2_b200.store import PackedStore, lengths_to_offsets
        _cpu_lib_standins(_lib)

        class CpuRetriever(hrc.JinaColBERTRetriever):          # the product has no CPU path: this test fakes the device calls
            device = property(lambda self: torch.device("cpu"))

            def search_keys(self, query_embeddings, k):
                s = self.store
                sc = o.maxsim_scores(query_embeddings.float(), s.tokens.float(), s.offsets)
                ids = np.broadcast_to(np.arange(s.n_docs) + s.doc_id_base, sc.shape)
                return torch.from_numpy(o.merge_keys(o.make_keys(sc.numpy(), ids), min(k, s.n_docs)).view(np.int64).copy())

        g = torch.Generator().manual_seed(11)
        n_docs, nq = 83, 4
        lens = torch.randint(1, 25, (n_docs,), generator=g)
        off = lengths_to_offsets(lens)
        tok = o.round_bf16(torch.nn.functional.normalize(torch.randn((int(off[-1]), 128), generator=g), dim=-1))
        q = o.round_bf16(torch.nn.functional.normalize(torch.randn((nq, 32, 128), generator=g), dim=-1))
        bm25 = torch.randint(0, n_docs, (nq, 30), generator=g, dtype=torch.int32)
        bm25[1, 25:] = -1                                       # absent entries
        bm25[2, 3] = n_docs + 5                                 # an id no shard holds: -inf, answered by rank 0
        cfg = hrc.RAGConfig(colbert_top_k=30, rerank_candidates=20, final_top_k=6)
        full = PackedStore.from_packed(tok, off, device="cpu")
        r = CpuRetriever(cfg, encoder=hrc.SyntheticEncoder())
        r.store = full.shard(rank, world)
        s = hrc.ShardedSearcher(r, transport="torch")
        ids, scores = s.retrieve_batch(q, bm25)
        assert s.n_docs_global() == n_docs
        np.save(os.path.join(out_dir, f"hy_ids_{rank}.npy"), ids.numpy())
        np.save(os.path.join(out_dir, f"hy_sc_{rank}.npy"), scores.numpy())
        if rank == 0:                                           # the unsharded pipeline, stage by stage, with the oracle
            sc = o.maxsim_scores(q, tok, off)
            exp_ids, exp_sc = [], []
            for i in range(nq):
                order = o.topk_deterministic(sc[i], 30)[0].tolist()
                fused, _ = o.rrf_ids(bm25[i].tolist(), order, 60)
                cand = (fused + [-1] * 20)[:20]
                cs = torch.tensor([float(sc[i, c]) if 0 <= c < n_docs else float("-inf") for c in cand])
                top = o.topk_deterministic(cs, 6)[0].tolist()
                exp_ids.append([cand[p] for p in top])
                exp_sc.append([float(cs[p]) for p in top])
            np.save(os.path.join(out_dir, "hy_exp_ids.npy"), np.array(exp_ids))
            np.save(os.path.join(out_dir, "hy_exp_sc.npy"), np.array(exp_sc, dtype=np.float32))
    finally:
        dist.destroy_process_group()


def test_sharded_hybrid_pipeline_host_logic_gloo(tmp_path):
    world = 2
    mp.spawn(_hybrid_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    exp_ids, exp_sc = np.load(tmp_path / "hy_exp_ids.npy"), np.load(tmp_path / "hy_exp_sc.npy")
    for r in range(world):
        assert (np.load(tmp_path / f"hy_ids_{r}.npy") == exp_ids).all(), f"rank {r}"
        assert (np.load(tmp_path / f"hy_sc_{r}.npy") == exp_sc).all(), f"rank {r}"
