#!/usr/bin/env python
"""Print the essentials of a bench.py JSON line (headline, roofline, e2e, secondary, breakdown)."""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
r = d["roofline"]
print(f"N={d['n_gpus']} value {d['value'] / 1e6:.1f} M docs/s  step {d['ms_per_step']:.3f} ms  kernel {r['kernel_ms']:.3f} ms  "
      f"{r['achieved']:.0f} GB/s  frac {r['frac']:.3f}  vs read peak {r.get('frac_vs_read_peak')}  launches {d['gpu_launches']}")
e = d["e2e"]
print(f"  e2e {e['value'] / 1e6:.1f} M docs/s  step {e['ms_per_step']:.3f} ms  kernel {e.get('kernel_ms', {}).get('median') if e.get('kernel_ms') else None}")
print(f"  clocks {d.get('clocks')}")
for k in ("parity_check", "breakdown"):
    if k in d:
        print(f"  {k}: {json.dumps(d[k])[:700]}")
for k, v in (d.get("secondary") or {}).items():
    v = dict(v)
    v.pop("what", None)
    print(f"  {k}: {json.dumps(v)[:1000]}")
if "cpu_baseline" in d:
    print(f"  cpu {d['cpu_baseline']['value'] / 1e6:.3f} M docs/s on {d['cpu_baseline']['cores']} cores")
