#!/bin/bash
# same-box A/B of two library builds: alternate them, C2 bench + ragged + C3(64 queries)
P=/root/repo/hybrid-rag-colbertv2_b200
for rep in 1 2; do
for lib in libhrc_prev.so libhrc.so; do
  export HRC_LIB_PATH=$P/$lib
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench.log 2>&1; echo -n "$lib  "; python scripts/fmt_bench.py gpurun_out/ab_bench.log | cut -c1-150
  python scripts/bench_configs.py --configs ragged,c3 --c3-queries 64 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        if 'ragged' in d['config']: print('   ragged GB/s', round(d['achieved_GBps']))
        else: print('   ', d['config'][:24], 'TF', round(d['useful_TFLOPs']), 'MHz', d['sm_mhz'], 'W', d['power_w'])
"
done
done
