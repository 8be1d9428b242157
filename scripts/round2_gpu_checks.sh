#!/bin/bash
# First GPU call of the next round (one GPU, ~4 minutes): everything that was
# prepared after round 1's GPU budget ran out. See DESIGN.md section 8.
#   gpurun --timeout 900 -- 'bash scripts/round2_gpu_checks.sh'
O=gpurun_out/round2
mkdir -p $O
export MKB_TEST_EXPERIMENTAL=1
# 1. device paths never run before: fibre-tissue pair, lean row-slab kernel
timeout 300 python -m pytest tests/test_fiber_tissue_gpu.py tests/test_persistent_gpu.py tests/test_multigpu_gpu.py -q -m gpu -k "fiber or pair or lean or persistent or split" 2>&1 | tail -15 | tee $O/experimental_tests.log
# 2. the regular GPU suite (the V-tile race fix and the larger kernel argument struct are new)
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/gpu_tests.log
# 3. C3 kernel variants: default, div_parallel, estrin, both
SWEEP_STEPS=50 timeout 300 python scripts/sweep_c3.py 2>&1 | tee $O/sweep_c3.log
# 4. fp32 kernel with the 6-instruction expf; small cables with the persistent kernel
timeout 300 python scripts/bench_configs.py c4 c1 2>&1 | grep -v Warn | tee $O/c4_c1.log
