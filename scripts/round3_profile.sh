#!/bin/bash
# Runs on the GPU box (gpurun): bench lines, launch list and the ncu capture for profiles/r03_*.
O=gpurun_out/r03
mkdir -p $O
python -c "
import sys, json; sys.path.insert(0, '.')
import myokit_b200
from myokit_b200 import capi
print(json.dumps(dict(capi.measure_peaks(0), device=capi.device_info(0))))" 2>/dev/null | tail -1 > $O/pipe_peaks.json
cat $O/pipe_peaks.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_ref.json
cut -c1-300 $O/bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 5 2>$O/bench_n1.err | tail -1 > $O/bench_n1.json
cut -c1-400 $O/bench_n1.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --advance 200 --scale-grid 0 > $O/bench_under_ncu.log 2>&1
MKB_PROFILE_KEYFILE=$O/key_c3.txt timeout 300 ncu --set full --clock-control none --import-source on -k regex:mkb_cell_step -s 4 -c 1 -f -o $O/prof_c3 python scripts/profile_target.py c3 6 > $O/ncu_c3.log 2>&1
tail -1 $O/ncu_c3.log
timeout 300 python scripts/bench_configs.py c1 c2 c4 2>/dev/null | grep -v Warn | tee $O/configs.txt
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $O/gpu_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -4 | tee $O/smoke.log
ls -la $O
