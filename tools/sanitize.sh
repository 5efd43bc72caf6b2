#!/bin/bash
# compute-sanitizer over the small-shape GPU tests (run on the B200 box: gpurun -- bash tools/sanitize.sh).
# Writes gpurun_out/sanitizer_<tool>.log (tail) + gpurun_out/sanitizer_summary.txt; commit the summary under profiles/.
set -u
OUT=gpurun_out
mkdir -p $OUT
SEL_E2E='tests/test_gpu_e2e.py -k tiny'
SEL_K='tests/test_gpu_kernels.py -k "conv_tc or head_tc or groupnorm_silu_fir or combine or conv_in4 or conv_out4 or gn_stats"'
: > $OUT/sanitizer_summary.txt
for tool in memcheck racecheck synccheck; do
  for sel in "$SEL_E2E" "$SEL_K"; do
    tag=$(echo "$sel" | cut -d/ -f2 | cut -d. -f1)
    log=$OUT/sanitizer_${tool}_${tag}.log
    start=$(date +%s)
    eval timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --target-processes all --error-exitcode 99 --print-limit 20 \
        python -m pytest $sel -x -q -p no:cacheprovider > $log 2>&1
    rc=$?
    end=$(date +%s)
    errs=$(grep -c "========= .*\(Invalid\|hazard\|Error\|error\|Barrier\)" $log)
    summ=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -1)
    pyt=$(grep -E "passed|failed" $log | tail -1)
    echo "$tool | $sel | rc=$rc | $((end-start)) s | flagged_lines=$errs | $summ | pytest: $pyt" | tee -a $OUT/sanitizer_summary.txt
    tail -c 6000 $log > $log.tail && mv $log.tail $log
  done
done
