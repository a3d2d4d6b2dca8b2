"""Known-answer vectors for TRUE MaxSim, computed in numpy float64 with explicit loops — independent of torch, of the
oracle and of the kernels (the reference itself never computes MaxSim, SURVEY.md F2, so these are the committed
fixed points the oracle's `maxsim_scores` and the CUDA kernels are both checked against).

    python tests/golden/make_kat.py        # writes tests/golden/maxsim_kat_f64.npz (no reference needed)

Case A: 3 ragged documents (5, 2, 4 tokens), 2 queries x 3 tokens.   Case B: 7 documents whose lengths straddle the
kernels' 32-column chunk and 128-token tile (1, 31, 32, 33, 127, 128, 129 tokens), 1 query x 32 tokens.
Inputs are unit-normalised normals rounded to bf16 (stored as float32, exactly representable).
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def bf16_round(x):
    """float32 -> nearest-even bfloat16 -> float32, in pure numpy."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def rows(rng, n):
    v = rng.standard_normal((n, 128))
    return bf16_round((v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32))


def maxsim_f64(q, tok, off):
    q64, t64 = q.astype(np.float64), tok.astype(np.float64)
    out = np.empty((q.shape[0], len(off) - 1), dtype=np.float64)
    for b in range(q.shape[0]):
        for d in range(len(off) - 1):
            total = 0.0
            for i in range(q.shape[1]):
                best = -np.inf
                for t in range(off[d], off[d + 1]):
                    dot = 0.0
                    for c in range(128):
                        dot += q64[b, i, c] * t64[t, c]
                    best = max(best, dot)
                total += best
            out[b, d] = total
    return out


def main():
    rng = np.random.default_rng(20260108)
    out = {}
    for name, lens, bq, lq in (("a", [5, 2, 4], 2, 3), ("b", [1, 31, 32, 33, 127, 128, 129], 1, 32)):
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        tok = rows(rng, int(off[-1]))
        q = rows(rng, bq * lq).reshape(bq, lq, 128)
        out[f"{name}_q"], out[f"{name}_tok"], out[f"{name}_off"] = q, tok, off
        out[f"{name}_scores_f64"] = maxsim_f64(q, tok, off)
    np.savez_compressed(os.path.join(HERE, "maxsim_kat_f64.npz"), **out)
    print("wrote maxsim_kat_f64.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
