#!/bin/bash
# Last call of the round: the GPU suite and the driver's bench command on the final tree.
O=gpurun_out/r03
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $O/gpu_tests_final.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 2>$O/bench_final.err | tee $O/bench_final.json | cut -c1-330
