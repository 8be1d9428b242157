#!/bin/bash
# compute-sanitizer over every kernel family (race detection); logs -> profiles/r02_sanitizer.md
O=gpurun_out/r02
mkdir -p $O
for tool in memcheck racecheck; do
  for w in plain vector stream slab graph persistent; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_target.py $w > $O/san_${tool}_$w.log 2>&1
    echo "$tool $w: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/san_${tool}_$w.log | tail -1) | $(grep -E '^(plain|vector|stream|slab|graph|persistent) ' $O/san_${tool}_$w.log | tail -1)"
  done
done | tee $O/sanitizer_summary.txt
