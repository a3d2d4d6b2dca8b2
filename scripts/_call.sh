mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
P=/root/repo/hybrid-rag-colbertv2_b200
run() { python scripts/bench_configs.py --configs $1 --docs 300000 --c3-queries 64 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        if 'ragged' in d['config']: print('   ragged GB/s', round(d['achieved_GBps']))
        else: print('   ', d['config'][:24], 'TF', round(d['useful_TFLOPs']), 'ms', round(d['kernel_ms'],2), 'MHz', d['sm_mhz'], 'W', d['power_w'])
"; }
for rep in 1 2; do
  echo "a (doc-aligned)";      HRC_LIB_PATH=$P/libhrc_a.so run c3
  echo "new (pipelined)";       HRC_LIB_PATH=$P/libhrc.so run c3
done
echo "new, no MMA (epilogue + TMA only)"; HRC_TC_DEBUG=4 HRC_LIB_PATH=$P/libhrc.so run c3
