"""Guard-band and poisoned-output checks of every C-ABI entry point (raw ctypes calls, caller-owned buffers).

compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer_closed_on_this_pool.txt), so its memcheck /
initcheck roles are played here: every input, output and workspace of a call is carved out of ONE device arena with
4 KB guard bands on both sides, the arena is filled with a pattern, outputs are poisoned, workspaces are given at
EXACTLY the size the *_workspace_bytes function returns — and after the call every guard band must be untouched
(no out-of-bounds write, workspace sizes honest), every output element overwritten (no uninitialised result) and
the results equal to the oracle's.  (Races: tests/test_gpu_parity.py::test_results_are_bitwise_repeatable.)"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as o

pytestmark = pytest.mark.gpu

GUARD = 4096
PATTERN = 0xA5
POISON = 0xFF        # fp32 0xFFFFFFFF = NaN, int32 = -1, uint64 = all ones: never a legal result here


class Arena:
    def __init__(self, dev, nbytes=512 << 20):
        self.buf = torch.full((nbytes,), PATTERN, dtype=torch.uint8, device=dev)
        self.off = GUARD
        self.regions = []          # (offset, nbytes)

    def alloc(self, nbytes, poison=False, src=None):
        """A 256-byte aligned region followed by a guard band; returns (device pointer, uint8 view)."""
        nbytes = int(nbytes)
        off = (self.off + 255) & ~255
        assert off + nbytes + GUARD <= self.buf.numel(), "arena too small"
        view = self.buf[off:off + nbytes]
        if src is not None:
            view.copy_(src.contiguous().view(torch.uint8).reshape(-1).to(self.buf.device))
        elif poison:
            view.fill_(POISON)
        self.regions.append((off, nbytes))
        self.off = off + nbytes + GUARD
        return self.buf.data_ptr() + off, view

    def check_guards(self, what):
        torch.cuda.synchronize()
        mask = torch.ones(self.off, dtype=torch.bool, device=self.buf.device)
        for off, n in self.regions:
            mask[off:off + n] = False
        bad = (self.buf[:self.off][mask] != PATTERN).nonzero()
        assert bad.numel() == 0, f"{what}: {bad.numel()} guard bytes overwritten (first at arena offset {int(bad[0])})"


def _typed(view, dtype, shape):
    return view.view(dtype).reshape(shape)


def _case(seed, lens, bq, lq):
    g = torch.Generator().manual_seed(seed)
    lens = torch.as_tensor(lens, dtype=torch.int64)
    off = torch.zeros(lens.numel() + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(lens, 0)
    tok = torch.nn.functional.normalize(torch.randn((int(off[-1]), 128), generator=g), dim=-1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn((bq, lq, 128), generator=g), dim=-1).to(torch.bfloat16)
    return q, tok, off


def _close(got, exp, what):
    fin = torch.isfinite(exp)
    assert not torch.isnan(got).any(), f"{what}: an output element was never written (poison left)"
    assert torch.equal(torch.isfinite(got), fin), what
    err = float((got[fin] - exp[fin]).abs().max() / exp[fin].abs().max().clamp_min(1e-6))
    assert err <= 4e-6, f"{what}: {err}"


@pytest.mark.parametrize("nq,lq,path", [(1, 32, "tc"), (2, 32, "tc"), (4, 17, "tc"), (1, 32, "dm"), (1, 20, "dm"), (7, 32, "tc"),
                                        (16, 32, "tc"), (24, 32, "tc"), (2, 70, "tc"), (9, 40, "tc"), (2, 32, "simt")])
def test_scoring_entry_points_stay_inside_their_buffers(cuda_dev, nq, lq, path):
    from hybrid_rag_colbertv2_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(7)
    lens = np.concatenate([rng.integers(1, 200, 400), [0, 1, 31, 32, 33, 127, 128, 129, 400]]).tolist()
    q, tok, off = _case(5, lens, nq, lq)
    n_docs, T = len(lens), int(off[-1])
    exp = o.maxsim_scores(q.float(), tok.float(), off)
    P = {"tc": L.PATH_TC, "dm": L.PATH_TC_DM, "simt": L.PATH_SIMT}[path]
    st = torch.cuda.current_stream(cuda_dev).cuda_stream
    A = Arena(cuda_dev)
    p_tok, _ = A.alloc(T * 256, src=tok)
    p_off, _ = A.alloc((n_docs + 1) * 8, src=off)
    p_q, _ = A.alloc(nq * lq * 256, src=q)
    ws_bytes = lib.hrc_maxsim_workspace_bytes(n_docs, nq, lq)
    p_ws, _ = A.alloc(max(ws_bytes, 16), poison=True)
    p_sc, v_sc = A.alloc(nq * n_docs * 4, poison=True)
    rc = lib.hrc_maxsim_scores(p_tok, p_off, n_docs, T, p_q, nq, lq, p_sc, P, p_ws if ws_bytes else None, ws_bytes, st)
    assert rc == 0, lib.hrc_last_error()
    A.check_guards(f"maxsim_scores {nq}x{lq} {path}")
    _close(_typed(v_sc, torch.float32, (nq, n_docs)).cpu(), exp, f"maxsim_scores {nq}x{lq} {path}")
    # candidates
    n_cand = 37
    cand = torch.from_numpy(rng.integers(0, n_docs, (nq, n_cand)).astype(np.int32))
    cand[0, 3] = -1
    p_cand, _ = A.alloc(nq * n_cand * 4, src=cand)
    cws = lib.hrc_maxsim_workspace_bytes(n_cand, nq, lq)
    p_cws, _ = A.alloc(max(cws, 16), poison=True)
    p_cs, v_cs = A.alloc(nq * n_cand * 4, poison=True)
    rc = lib.hrc_maxsim_scores_ids(p_tok, p_off, n_docs, T, p_cand, n_cand, p_q, nq, lq, p_cs, P, p_cws if cws else None, cws, st)
    assert rc == 0, lib.hrc_last_error()
    A.check_guards(f"maxsim_scores_ids {nq}x{lq} {path}")
    expc = torch.gather(exp, 1, cand.clamp(0, n_docs - 1).long())
    expc[0, 3] = float("-inf")
    _close(_typed(v_cs, torch.float32, (nq, n_cand)).cpu(), expc, f"maxsim_scores_ids {nq}x{lq} {path}")
    # a workspace one byte too small must be refused, not overrun
    if ws_bytes and lq > 32:      # (for one short query the 256 bytes hold an OPTIONAL claim counter: nothing to refuse)
        rc = lib.hrc_maxsim_scores(p_tok, p_off, n_docs, T, p_q, nq, lq, p_sc, P, p_ws, ws_bytes - 1, st)
        assert rc != 0 and b"workspace" in lib.hrc_last_error()


@pytest.mark.parametrize("nq,lq,k", [(1, 32, 100), (3, 32, 128), (8, 20, 10), (24, 32, 100), (2, 32, 200), (2, 40, 50)])
def test_search_rerank_hybrid_stay_inside_their_buffers(cuda_dev, nq, lq, k):
    """hrc_search (fused for k <= 128 and lq <= 32, staged otherwise), hrc_search_host, hrc_rerank (one launch / staged)
    and hrc_hybrid_retrieve with workspaces of exactly the advertised size."""
    from hybrid_rag_colbertv2_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(9)
    lens = rng.integers(1, 40, 6000).tolist()
    q, tok, off = _case(6, lens, nq, lq)
    n_docs, T = len(lens), int(off[-1])
    exp = o.maxsim_scores(q.float(), tok.float(), off)
    st = torch.cuda.current_stream(cuda_dev).cuda_stream
    A = Arena(cuda_dev)
    p_tok, _ = A.alloc(T * 256, src=tok)
    p_off, _ = A.alloc((n_docs + 1) * 8, src=off)
    p_q, _ = A.alloc(nq * lq * 256, src=q)
    base = 1000
    # --- search
    wsb = lib.hrc_search_workspace_bytes(n_docs, T, nq, lq, k, L.PATH_AUTO)
    p_ws, _ = A.alloc(wsb, poison=True)
    p_keys, v_keys = A.alloc(nq * k * 8, poison=True)
    p_ids, v_ids = A.alloc(nq * k * 4, poison=True)
    p_sc, v_sc = A.alloc(nq * k * 4, poison=True)
    rc = lib.hrc_search(p_tok, p_off, n_docs, T, p_q, nq, lq, k, base, p_ws, wsb, p_keys, p_ids, p_sc, L.PATH_AUTO, st)
    assert rc == 0, lib.hrc_last_error()
    A.check_guards(f"search {nq}x{lq} k={k}")
    ids = _typed(v_ids, torch.int32, (nq, k)).cpu()
    sc = _typed(v_sc, torch.float32, (nq, k)).cpu()
    keys = _typed(v_keys, torch.int64, (nq, k)).cpu().numpy().view(np.uint64)
    ui, us = o.unpack_keys(keys)
    assert (ui == ids.numpy()).all() and (us == sc.numpy()).all()
    for i in range(nq):
        assert o.check_ranking((ids[i] - base).tolist(), sc[i].tolist(), exp[i], k, 1e-3) is None
    rc = lib.hrc_search(p_tok, p_off, n_docs, T, p_q, nq, lq, k, base, p_ws, wsb - 1, p_keys, p_ids, p_sc, L.PATH_AUTO, st)
    assert rc != 0 and b"workspace" in lib.hrc_last_error()
    # --- search_host (pinned host buffers)
    hq = q.float().pin_memory()
    h_ids = torch.full((nq, k), -7, dtype=torch.int32).pin_memory()
    h_sc = torch.full((nq, k), float("nan")).pin_memory()
    hwb = lib.hrc_search_host_workspace_bytes(n_docs, T, nq, lq, k, L.PATH_AUTO)
    p_hws, _ = A.alloc(hwb, poison=True)
    rc = lib.hrc_search_host(p_tok, p_off, n_docs, T, hq.data_ptr(), nq, lq, k, base, p_hws, hwb, h_ids.data_ptr(),
                             h_sc.data_ptr(), L.PATH_AUTO, st)
    assert rc == 0, lib.hrc_last_error()
    A.check_guards(f"search_host {nq}x{lq} k={k}")
    assert torch.equal(h_ids, ids) and torch.equal(h_sc, sc)
    # --- rerank
    n_cand, rk = 50, 10
    cand = torch.from_numpy(rng.integers(0, n_docs, (nq, n_cand)).astype(np.int32))
    cand[nq - 1, 4] = n_docs
    p_cand, _ = A.alloc(nq * n_cand * 4, src=cand)
    rwb = lib.hrc_rerank_workspace_bytes(n_cand, nq, lq, rk)
    p_rws, _ = A.alloc(rwb, poison=True)
    p_pos, v_pos = A.alloc(nq * rk * 4, poison=True)
    p_rid, v_rid = A.alloc(nq * rk * 4, poison=True)
    p_rsc, v_rsc = A.alloc(nq * rk * 4, poison=True)
    rc = lib.hrc_rerank(p_tok, p_off, n_docs, T, p_cand, n_cand, p_q, nq, lq, rk, p_rws, rwb, p_pos, p_rid, p_rsc, None,
                        L.PATH_AUTO, st)
    assert rc == 0, lib.hrc_last_error()
    A.check_guards(f"rerank {nq}x{lq}")
    pos = _typed(v_pos, torch.int32, (nq, rk)).cpu()
    rsc = _typed(v_rsc, torch.float32, (nq, rk)).cpu()
    rid = _typed(v_rid, torch.int32, (nq, rk)).cpu()
    expc = torch.gather(exp, 1, cand.clamp(0, n_docs - 1).long())
    expc[nq - 1, 4] = float("-inf")
    for i in range(nq):
        assert o.check_ranking(pos[i].tolist(), rsc[i].tolist(), expc[i], rk, 1e-3) is None
    assert torch.equal(rid, torch.gather(cand, 1, pos.long()))
    # --- hybrid retrieve
    if k <= n_docs:
        n_bm, ck, nc, fk = 30, min(k, 100), 20, 5
        bm = torch.from_numpy((rng.integers(0, n_docs, (nq, n_bm)) + base).astype(np.int32))
        p_bm, _ = A.alloc(nq * n_bm * 4, src=bm)
        hyb = lib.hrc_hybrid_retrieve_workspace_bytes(n_docs, T, nq, lq, ck, nc, fk, L.PATH_AUTO)
        p_hy, _ = A.alloc(hyb, poison=True)
        p_oi, v_oi = A.alloc(nq * fk * 4, poison=True)
        p_os, v_os = A.alloc(nq * fk * 4, poison=True)
        rc = lib.hrc_hybrid_retrieve(p_tok, p_off, n_docs, T, p_q, nq, lq, p_bm, n_bm, ck, 60, nc, fk, base, p_hy, hyb, p_oi,
                                     p_os, L.PATH_AUTO, st)
        assert rc == 0, lib.hrc_last_error()
        A.check_guards(f"hybrid {nq}x{lq}")
        oi = _typed(v_oi, torch.int32, (nq, fk)).cpu()
        os_ = _typed(v_os, torch.float32, (nq, fk)).cpu()
        for i in range(nq):
            col = (ids[i, :ck]).tolist()
            fused, _ = o.rrf_ids(bm[i].tolist(), col, 60)
            cands = [c - base for c in fused[:nc]]
            got_pos = [cands.index(int(x) - base) for x in oi[i].tolist()]
            assert o.check_ranking(got_pos, os_[i].tolist(), exp[i][cands], fk, 1e-3) is None


def test_selection_and_fusion_entry_points_stay_inside_their_buffers(cuda_dev):
    from hybrid_rag_colbertv2_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(11)
    st = torch.cuda.current_stream(cuda_dev).cuda_stream
    A = Arena(cuda_dev)
    for n, rows, k in ((50, 2, 10), (8192, 1, 100), (8193, 2, 100), (100_000, 3, 128), (30_000, 1, 2048)):
        s = torch.from_numpy(rng.standard_normal((rows, n)).astype(np.float32))
        p_s, _ = A.alloc(rows * n * 4, src=s)
        wsb = lib.hrc_topk_workspace_bytes(n, rows, k)
        p_ws, _ = A.alloc(max(wsb, 16), poison=True)
        p_k, v_k = A.alloc(rows * k * 8, poison=True)
        rc = lib.hrc_topk(p_s, None, n, rows, k, 5, p_k, p_ws if wsb else None, wsb, st)
        assert rc == 0, lib.hrc_last_error()
        A.check_guards(f"topk n={n}")
        ref = o.merge_keys(o.make_keys(s.numpy(), np.broadcast_to(np.arange(n) + 5, s.shape)), k)
        got = _typed(v_k, torch.int64, (rows, k)).cpu().numpy().view(np.uint64)
        assert (got == ref).all()
        p_i, v_i = A.alloc(rows * k * 4, poison=True)
        p_f, v_f = A.alloc(rows * k * 4, poison=True)
        assert lib.hrc_keys_unpack(p_k, rows * k, p_i, p_f, st) == 0
        A.check_guards("keys_unpack")
        ui, us = o.unpack_keys(ref)
        assert (_typed(v_i, torch.int32, (rows, k)).cpu().numpy() == ui).all()
        assert (_typed(v_f, torch.float32, (rows, k)).cpu().numpy() == us).all()
    allk = np.concatenate([o.make_keys(rng.standard_normal((3, 100)).astype(np.float32), np.arange(100)[None] + 1000 * r)
                           for r in range(8)], 1)
    p_in, _ = A.alloc(allk.nbytes, src=torch.from_numpy(allk.view(np.int64).copy()))
    p_out, v_out = A.alloc(3 * 100 * 8, poison=True)
    assert lib.hrc_topk_merge(p_in, 800, 3, 100, p_out, st) == 0
    A.check_guards("topk_merge")
    assert (_typed(v_out, torch.int64, (3, 100)).cpu().numpy().view(np.uint64) == o.merge_keys(allk, 100)).all()
    a = rng.integers(0, 300, (5, 100)).astype(np.int32)
    b = rng.integers(0, 300, (5, 100)).astype(np.int32)
    p_a, _ = A.alloc(a.nbytes, src=torch.from_numpy(a))
    p_b, _ = A.alloc(b.nbytes, src=torch.from_numpy(b))
    p_fi, v_fi = A.alloc(5 * 50 * 4, poison=True)
    p_fs, v_fs = A.alloc(5 * 50 * 8, poison=True)
    p_fc, v_fc = A.alloc(5 * 4, poison=True)
    assert lib.hrc_rrf_fuse(p_a, 100, p_b, 100, 5, 60, 50, p_fi, p_fs, p_fc, st) == 0
    A.check_guards("rrf_fuse")
    fi = _typed(v_fi, torch.int32, (5, 50)).cpu()
    fs = _typed(v_fs, torch.float64, (5, 50)).cpu()
    for r in range(5):
        ri, rs = o.rrf_ids(a[r].tolist(), b[r].tolist(), 60)
        assert fi[r].tolist() == ri[:50] and fs[r].tolist() == rs[:50]
    # mean-pool cosine (the reference's function as coded)
    q, tok, off = _case(8, rng.integers(1, 60, 500).tolist(), 2, 32)
    p_tok, _ = A.alloc(tok.numel() * 2, src=tok)
    p_off, _ = A.alloc(off.numel() * 8, src=off)
    p_q, _ = A.alloc(q.numel() * 2, src=q)
    p_m, v_m = A.alloc(2 * 500 * 4, poison=True)
    assert lib.hrc_meanpool_cosine_scores(p_tok, p_off, 500, int(off[-1]), p_q, 2, 32, p_m, st) == 0
    A.check_guards("meanpool_cosine")
    got = _typed(v_m, torch.float32, (2, 500)).cpu()
    assert not torch.isnan(got).any()
    for d in (0, 17, 499):
        e = o.literal_reference(q.float(), tok[int(off[d]):int(off[d + 1])].float())
        assert float((got[:, d] - e).abs().max()) < 1e-5
