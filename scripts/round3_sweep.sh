#!/bin/bash
# One GPU: option sweep of the C3 kernel (+ optionally ncu captures, the GPU suite, bench lines).
#   gpurun --timeout 1200 -- 'SWEEP_SET=r3c PROFILE=1 bash scripts/round3_sweep.sh'
O=gpurun_out/r3
mkdir -p $O
S=${SWEEP_SET:-r3a}
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "staged" 2>&1 | tail -15 | tee $O/staged_tests.log
timeout 400 env SWEEP_SET=$S SWEEP_STEPS=30 python scripts/sweep_c3.py 2>&1 | grep -v Warn | tee $O/sweep_${S}_${SWEEP_GRID:-2048}.log
if [ -n "$PROFILE" ]; then
  MKB_PROFILE_KEYFILE=$O/key_c3.txt timeout 300 ncu --set full --clock-control none --import-source on -k regex:mkb_cell_step -s 4 -c 1 -f -o $O/prof_c3 python scripts/profile_target.py c3 6 > $O/ncu_c3.log 2>&1
  tail -1 $O/ncu_c3.log
fi
if [ -n "$FULL" ]; then
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $O/gpu_tests.log
timeout 400 python bench.py --steps 20 --warmup 3 2>$O/bench.err | tee $O/bench_20.json
timeout 400 python bench.py --no-cpu --scale-grid 0 2>>$O/bench.err | tee $O/bench_200.json
fi
