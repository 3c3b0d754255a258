#!/bin/bash
# tools/sanitize.sh OUTDIR -- compute-sanitizer over the small configurations (tools/sanitize_driver.py).
# racecheck only sees shared memory; the global AGG/PREF/ticket protocols are relaxed atomics by design.
out=${1:-gpurun_out}
for tool in memcheck racecheck synccheck; do
  for fam in sort scan rng; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py $fam > $out/r02_sanitizer_${tool}_${fam}.log 2>&1
    echo "$tool $fam rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/r02_sanitizer_${tool}_${fam}.log | tail -1)"
  done
done
