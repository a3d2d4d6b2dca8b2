#!/bin/bash
# ncu evidence for the bench command: (1) launch list with per-launch device time, (2) one full
# capture of the dominant kernel.  Each ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_tc -s 3 -c 1 -f -o gpurun_out/prof_maxsim_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
python scripts/fmt_bench.py gpurun_out/plain.log
# batched kernel (C3 shape, reduced corpus so the replay passes stay short)
CMD3="python scripts/bench_configs.py --configs c3 --docs 150000 --c3-queries 64"
$CMD3 > gpurun_out/c3_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_tc -s 20 -c 1 -f -o gpurun_out/prof_c3 $CMD3 > gpurun_out/ncu_c3.log 2>&1
echo "c3 capture exit $?"
