#!/bin/bash
# Round-2 ncu evidence.  Each ncu run follows a plain run of the same command that exited 0.
#  (1) launch list of the bench command (per-launch device time; cold-cache, serialised)
#  (2) full capture of the dominant kernel of the headline: maxsim_dm_kernel<TK> (doc-major, fused top-k)
#  (3) full capture of the query-major single-query kernel on the same corpus (HRC_PATH_TC), for the A/B of the two
#  (4) full capture of the batched CTA-pair kernel on a C3-shaped run (reduced corpus: replay passes stay short)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
$CMD > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/r02_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_dm -s 3 -c 1 -f -o gpurun_out/r02_prof_maxsim_dm $CMD > gpurun_out/r02_ncu_dm.log 2>&1
echo "doc-major capture exit $?"
CMDQ="python scripts/profile_target.py qm"
$CMDQ > gpurun_out/r02_plain_qm.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_maxsim_qm $CMDQ > gpurun_out/r02_ncu_qm.log 2>&1
echo "query-major capture exit $?"
CMD3="python scripts/profile_target.py c3"
$CMD3 > gpurun_out/r02_plain_c3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_c3_pair $CMD3 > gpurun_out/r02_ncu_c3.log 2>&1
echo "c3 capture exit $?"
CMDT="python scripts/profile_target.py topk"
$CMDT > gpurun_out/r02_plain_topk.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_stream -s 2 -c 1 -f -o gpurun_out/r02_prof_topk_stream $CMDT > gpurun_out/r02_ncu_topk.log 2>&1
echo "topk capture exit $?"
ls -la gpurun_out/*.ncu-rep
