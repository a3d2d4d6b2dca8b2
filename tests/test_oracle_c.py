"""The plain-C oracle (oracle/maxsim_oracle.c) against the Python oracle and the reference's golden vectors."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def clib():
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    lib = ctypes.CDLL(path)
    lib.oracle_rrf.restype = ctypes.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_c_maxsim_matches_python_oracle(clib):
    g = torch.Generator().manual_seed(3)
    lens = torch.randint(0, 12, (15,), generator=g)
    off = np.concatenate([[0], np.cumsum(lens.numpy())]).astype(np.int64)
    tok = o.round_bf16(torch.nn.functional.normalize(torch.randn((int(off[-1]), 128), generator=g), dim=-1)).numpy()
    q = o.round_bf16(torch.nn.functional.normalize(torch.randn((2, 9, 128), generator=g), dim=-1)).numpy()
    out = np.empty((2, 15), dtype=np.float32)
    clib.oracle_maxsim_scores(_p(q), 2, 9, _p(tok), _p(off), ctypes.c_int64(15), _p(out))
    exp = o.maxsim_scores(torch.from_numpy(q), torch.from_numpy(tok), torch.from_numpy(off)).numpy()
    fin = np.isfinite(exp)
    assert (np.isfinite(out) == fin).all() and (out[~fin] == exp[~fin]).all()
    np.testing.assert_allclose(out[fin], exp[fin], rtol=1e-5, atol=1e-6)


def test_c_maxsim_pinned_by_reference_and_float64_vectors(clib, golden_dir):
    """The C restatement of MaxSim reproduces (a) the unmodified reference bit for bit on the shapes where its
    mean-pool cosine is MaxSim (maxsim_pin.npz) and (b) the float64 known answers (maxsim_kat_f64.npz)."""
    z = np.load(os.path.join(golden_dir, "maxsim_pin.npz"))
    qrows, drows = z["q_rows"], z["d_rows"]
    for lq in (1, 2, 4):
        for ld in (1, 2, 4):
            q = np.ascontiguousarray(np.repeat(qrows[:, None, :], lq, 1))
            tok = np.ascontiguousarray(np.repeat(drows[:, None, :], ld, 1).reshape(-1, 128))
            off = np.arange(0, drows.shape[0] * ld + 1, ld, dtype=np.int64)
            out = np.empty((qrows.shape[0], drows.shape[0]), dtype=np.float32)
            clib.oracle_maxsim_scores(_p(q), q.shape[0], lq, _p(tok), _p(off), ctypes.c_int64(drows.shape[0]), _p(out))
            assert (out / lq == z[f"out_lq{lq}_ld{ld}"]).all(), (lq, ld)
    k = np.load(os.path.join(golden_dir, "maxsim_kat_f64.npz"))
    for name in ("a", "b"):
        q, tok, off = (np.ascontiguousarray(k[f"{name}_{x}"]) for x in ("q", "tok", "off"))
        out = np.empty(k[f"{name}_scores_f64"].shape, dtype=np.float32)
        clib.oracle_maxsim_scores(_p(q), q.shape[0], q.shape[1], _p(tok), _p(off), ctypes.c_int64(len(off) - 1), _p(out))
        assert np.abs(out - k[f"{name}_scores_f64"]).max() <= 2e-6 * np.abs(k[f"{name}_scores_f64"]).max()


def test_c_maxsim_matches_vllm_outputs(clib, golden_dir):
    """The C restatement against the vectors vLLM's MaxSim functions produced (maxsim_vllm_pin.npz)."""
    from golden.make_vllm_pin import load_pin
    z = load_pin(os.path.join(golden_dir, "maxsim_vllm_pin.npz"))
    tok, off = np.ascontiguousarray(z["tok"]), np.ascontiguousarray(z["off"])
    for name in ("q32", "q7", "q1"):
        q, pair, _ = z[name]
        q = np.ascontiguousarray(q)
        out = np.empty(pair.shape, dtype=np.float32)
        clib.oracle_maxsim_scores(_p(q), q.shape[0], q.shape[1], _p(tok), _p(off), ctypes.c_int64(len(off) - 1), _p(out))
        assert np.abs(out - pair).max() <= 2e-6 * np.abs(pair).max(), name


def test_c_literal_matches_reference_outputs(clib, golden_dir):
    """The C restatement of what the reference's `_maxsim_score` literally computes, against the vectors the
    unmodified reference produced (both fixtures)."""
    for name in ("literal_maxsim.npz", "literal_bf16.npz"):
        z = np.load(os.path.join(golden_dir, name))
        q, D = np.ascontiguousarray(z["q"]), np.ascontiguousarray(z["D"])
        out = np.empty(D.shape[0], dtype=np.float32)
        clib.oracle_literal_scores(_p(q), q.shape[0], _p(D), ctypes.c_int64(D.shape[0]), D.shape[1], _p(out))
        np.testing.assert_allclose(out, z["out_q_D"], rtol=0, atol=2e-6)


def test_c_rrf_bit_equal_to_reference_fixtures(clib, golden_dir):
    for c in json.load(open(os.path.join(golden_dir, "rrf.json"))):
        a = np.asarray(c["a"], dtype=np.int32)
        b = np.asarray(c["b"], dtype=np.int32)
        ids = np.empty(len(a) + len(b) + 1, dtype=np.int32)
        sc = np.empty(len(a) + len(b) + 1, dtype=np.float64)
        n = clib.oracle_rrf(_p(a), len(a), _p(b), len(b), c["k"], _p(ids), _p(sc))
        assert ids[:n].tolist() == c["ids"]
        assert [repr(float(x)) for x in sc[:n]] == c["scores"]


def test_c_topk_keys_match_python_keys(clib):
    g = torch.Generator().manual_seed(5)
    s = torch.randn(500, generator=g)
    s[:100] = torch.round(s[:100])
    s[7] = float("nan")
    sn = s.numpy()
    out = np.empty(40, dtype=np.uint64)
    clib.oracle_topk_keys(_p(sn), None, ctypes.c_int64(500), 40, 11, _p(out))
    ref = o.merge_keys(o.make_keys(sn, np.arange(500) + 11)[None], 40)[0]
    assert (out == ref).all()
