"""An INDEPENDENT third-party pin for true MaxSim: golden vectors produced by vLLM's own ColBERT scoring function.

The reference never computes MaxSim, so no reference output exists for the function every kernel here is
graded on.  vLLM 0.22.0 (installed in the build image; Apache-2.0; not written by this repo) ships the same
late-interaction score in `vllm/entrypoints/pooling/scoring/utils.py:compute_maxsim_score` — "sum over query tokens of
max similarity to any doc token", fp32 matmul — and a batched form in `vllm/v1/pool/late_interaction.py:
compute_maxsim_score_batched`.  This script calls BOTH on seeded inputs and commits inputs + outputs; the oracle
(`tests/test_oracle.py`, `tests/test_oracle_c.py`) and the CUDA kernels (`tests/test_gpu_parity.py`) are then checked
against numbers neither this repo's oracle nor its kernels produced.

    python tests/golden/make_vllm_pin.py      # needs `import vllm` (CPU is enough); writes tests/golden/maxsim_vllm_pin.npz

Inputs: unit-normalised normals rounded to bf16 (stored as their 16 bits; `load_pin` below widens them back to float32): 33 ragged documents whose
lengths include 1, 31..33, 127..129, 255..257 and 300 tokens; 5 queries of 32 tokens plus one of 7 and one of 1 token.
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def load_pin(path):
    """-> dict: 'tok' fp32 [T,128], 'off' int64, and per case name ('q32','q7','q1'): (q fp32 [nq,lq,128], scores [nq,33])."""
    z = np.load(path)
    wide = lambda b: (b.astype(np.uint32) << 16).view(np.float32)      # noqa: E731
    out = {"tok": wide(z["tok_bf16_bits"]), "off": z["off"], "vllm_version": str(z["vllm_version"])}
    for name in ("q32", "q7", "q1"):
        out[name] = (wide(z[f"{name}_q_bf16_bits"]), z[f"{name}_scores"], z[f"{name}_scores_batched"])
    return out


def rows(g, n):
    v = torch.nn.functional.normalize(torch.randn((n, 128), generator=g), dim=-1)
    return v.to(torch.bfloat16).float()


def main():
    import vllm
    from vllm.entrypoints.pooling.scoring.utils import compute_maxsim_score
    from vllm.v1.pool.late_interaction import compute_maxsim_score_batched

    g = torch.Generator().manual_seed(20260109)
    lens = [1, 2, 31, 32, 33, 64, 127, 128, 129, 255, 256, 257, 300] + torch.randint(3, 100, (20,), generator=g).tolist()
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tok = rows(g, int(off[-1]))
    bits = lambda x: (x.numpy().view(np.uint32) >> 16).astype(np.uint16)      # exact: the values ARE bf16  # noqa: E731
    out = {"tok_bf16_bits": bits(tok), "off": off, "vllm_version": np.array(vllm.__version__)}
    for name, nq, lq in (("q32", 5, 32), ("q7", 1, 7), ("q1", 1, 1)):
        q = rows(g, nq * lq).reshape(nq, lq, 128)
        pair = torch.empty((nq, len(lens)), dtype=torch.float32)
        for b in range(nq):
            for d in range(len(lens)):
                pair[b, d] = compute_maxsim_score(q[b], tok[off[d]:off[d + 1]])
        q_list = [q[b] for b in range(nq) for _ in range(len(lens))]
        d_list = [tok[off[d]:off[d + 1]] for _ in range(nq) for d in range(len(lens))]
        batched = torch.stack([s.reshape(()) for s in compute_maxsim_score_batched(q_list, d_list)]).reshape(nq, len(lens))
        # the two vLLM code paths must agree with each other before either is trusted as a pin
        assert float((batched.float() - pair).abs().max()) <= 2e-6 * float(pair.abs().max()), name
        out[f"{name}_q_bf16_bits"], out[f"{name}_scores"], out[f"{name}_scores_batched"] = bits(q), pair.numpy(), batched.float().numpy()
    np.savez_compressed(os.path.join(HERE, "maxsim_vllm_pin.npz"), **out)
    print("wrote maxsim_vllm_pin.npz (vllm", vllm.__version__ + ")", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
