#!/usr/bin/env python
"""libhrc_exp.so experiment: when do the CTAs of the doc-major search finish, and on which SM?  The corpus is cut into
EQUAL token ranges, one per CTA; if SMs pull data at different rates (GPCs of 16 / 18 / 20 SMs sharing a port, near / far
L2 die), the slowest CTA sets the kernel time and the spread is what a dynamic work distribution could recover.

    python scripts/exp_cta_times.py [c2|ragged] [ctas]
"""
import ctypes
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HRC_LIB_PATH", os.path.join(ROOT, "hybrid-rag-colbertv2_b200", "libhrc_exp.so"))

import torch  # noqa: E402

from hybrid_rag_colbertv2_b200 import _lib as L  # noqa: E402
from hybrid_rag_colbertv2_b200.synth import synth_queries, synth_store  # noqa: E402

dev = torch.device("cuda:0")
lib = L.load()
lib.hrc_exp_set_ctas.argtypes = [ctypes.c_int]
lib.hrc_exp_set_cta_times.argtypes = [ctypes.c_void_p]
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
ctas = int(sys.argv[2]) if len(sys.argv) > 2 else 148
store = synth_store(1_000_000, 128, 128, seed=20260102, device=dev) if which == "c2" else synth_store(1_000_000, 32, 512, seed=20260103, device=dev)
q = synth_queries(1, 32, device=dev)
lib.hrc_exp_set_ctas(ctas if ctas != 148 else 0)
ws = L.Workspace()
times = torch.zeros(2 * 148, dtype=torch.int64, device=dev)
for _ in range(5):
    L.search(store.tokens, store.offsets, q, 100, workspace=ws, unpack=False)
torch.cuda.synchronize()
lib.hrc_exp_set_cta_times(times.data_ptr())
rows = []
for rep in range(5):
    times.zero_()
    L.search(store.tokens, store.offsets, q, 100, workspace=ws, unpack=False)
    torch.cuda.synchronize()
    t = times.cpu().view(-1, 2)[:ctas]
    end, smid = t[:, 0].double(), t[:, 1]
    rel = (end - end.min()) / 1e3                              # us after the first CTA finished
    rows.append((rel, smid))
lib.hrc_exp_set_cta_times(None)
rel = torch.stack([r[0] for r in rows]).median(0).values
smid = rows[-1][1]
order = torch.argsort(rel)
print(json.dumps({"corpus": which, "ctas": ctas,
                  "finish_spread_us": {"p10": round(float(rel.kthvalue(max(1, ctas // 10)).values), 1),
                                       "median": round(float(rel.median()), 1),
                                       "p90": round(float(rel.kthvalue(ctas - ctas // 10).values), 1),
                                       "max": round(float(rel.max()), 1)},
                  "note": "us after the first CTA finished (median of 5 launches per CTA); the kernel takes ~4,500 us (c2)"}))
# by SM id: do the late CTAs sit together (a GPC)?
late = sorted(int(smid[i]) for i in order[-20:])
early = sorted(int(smid[i]) for i in order[:20])
print(json.dumps({"last_20_ctas_smids": late, "first_20_ctas_smids": early}))
by_sm = sorted((int(smid[i]), round(float(rel[i]), 1)) for i in range(ctas))
print(json.dumps({"finish_us_by_smid": by_sm}))
