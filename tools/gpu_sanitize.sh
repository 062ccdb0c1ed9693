#!/bin/bash
# compute-sanitizer over one small optimisation step (the scatter-add backward and every kernel around it).
# Summaries land in gpurun_out/sanitize_*.log; copy the tails into profiles/.
mkdir -p gpurun_out
for mode in ${MODES:-tf32x3 bf16x3}; do
  for tool in memcheck racecheck initcheck synccheck; do
    out=gpurun_out/sanitize_${tool}_${mode}.log
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py $mode > $out 2>&1
    echo "== $tool $mode exit $? : $(grep -c 'ERROR SUMMARY' $out) summary line(s): $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $out | tail -1)"
    grep "sanitize_step" $out | tail -1
  done
done
