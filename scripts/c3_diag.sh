#!/bin/bash
# What bounds the batched kernel?  Same box, C3 shape (64 queries), with parts of the kernel disabled
# (HRC_TC_DEBUG bits: 1 = no epilogue math, 2 = no document TMA, 4 = no MMA).  Results in debug modes are garbage.
for dbg in 0 1 2 3 4 5 0; do
  echo -n "HRC_TC_DEBUG=$dbg  "
  HRC_TC_DEBUG=$dbg python scripts/bench_configs.py --configs c3 --docs 300000 --c3-queries 64 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        if '64 queries' in d['config']: print('TF', round(d['useful_TFLOPs']), 'ms', round(d['kernel_ms'],2), 'MHz', d['sm_mhz'], 'W', d['power_w'])
"
done
