#!/bin/bash
# One GPU, ~10 minutes: GPU suite after the ABI v5 / division / exp changes,
# the r3a option sweep of the C3 kernel, one bench line.
#   gpurun --timeout 1200 -- 'bash scripts/round3_sweep.sh'
O=gpurun_out/r3
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $O/gpu.txt
timeout 240 env SWEEP_SET=${SWEEP_SET:-r3a} SWEEP_STEPS=30 python scripts/sweep_c3.py 2>&1 | grep -v Warn | tee $O/sweep_${SWEEP_SET:-r3a}.log
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $O/gpu_tests.log
timeout 400 python bench.py 2>$O/bench.err | tee $O/bench.json
tail -5 $O/bench.err
