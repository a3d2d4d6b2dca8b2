#!/bin/bash
# compute-sanitizer over scripts/sanitizer_subset.py (VERDICT r1 "Next" #3).  One tool per run, each under its own
# timeout; full logs in gpurun_out/, the summaries are copied to profiles/r02_sanitizer_<tool>.txt by hand.
mkdir -p gpurun_out
python scripts/sanitizer_subset.py > gpurun_out/sanitizer_plain.txt 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitizer_plain.txt; exit 1; }
for tool in ${TOOLS:-memcheck synccheck racecheck initcheck}; do
  extra=""
  [ "$tool" = "initcheck" ] && extra="--track-unused-memory no"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 20 --launch-timeout 0 \
      python scripts/sanitizer_subset.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "$tool exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitizer_subset ok' gpurun_out/sanitizer_$tool.txt | tr '\n' ' ')"
done
