#!/bin/bash
# compute-sanitizer over the kernel families round 3 changed (staged states: plain grid, row slabs, tile loop)
# and the reference's own suite against the new defaults; logs -> profiles/r03_sanitizer.md
O=gpurun_out/r03
mkdir -p $O
for tool in memcheck racecheck; do
  for w in plain slab loop; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_target.py $w > $O/san_${tool}_$w.log 2>&1
    echo "$tool $w: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/san_${tool}_$w.log | tail -1) | $(grep -E '^(plain|slab|loop) ' $O/san_${tool}_$w.log | tail -1)"
  done
done | tee $O/sanitizer_summary.txt
timeout 900 python scripts/run_reference_suite.py 2>&1 | tail -15 | tee $O/reference_suite.txt
