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
tail -n 3 gpurun_out/plain.log | cut -c1-400
