#!/bin/bash
# 2 GPUs: kernel options of the row-slab form on slabs of 1024 and 256 rows
O=gpurun_out/r3
mkdir -p $O
timeout 400 python scripts/slab_sweep.py 2048 2>&1 | grep -v Warn | tee $O/slab_sweep_1024.txt
timeout 400 python scripts/slab_sweep.py 512 2>&1 | grep -v Warn | tee $O/slab_sweep_256.txt
