#!/bin/bash
# Runs on the GPU box (gpurun): bench line, launch list and ncu captures for profiles/r02_*.
O=gpurun_out/r02
mkdir -p $O
python -c "
import sys, json; sys.path.insert(0, '.')
import myokit_b200
from myokit_b200 import capi
print(json.dumps(dict(capi.measure_peaks(0), device=capi.device_info(0))))" 2>/dev/null | tail -1 > $O/pipe_peaks.json
cat $O/pipe_peaks.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tail -1 > $O/bench_ref.json
cut -c1-300 $O/bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 3 2>$O/bench_n1.err | tail -1 > $O/bench_n1.json
cut -c1-400 $O/bench_n1.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --advance 200 --scale-grid 0 > $O/bench_under_ncu.log 2>&1
for w in c3 stencil32 stencil64; do
  MKB_PROFILE_KEYFILE=$O/key_$w.txt MKB_PROFILE_OPTS="${PROFILE_OPTS:-dict()}" timeout 300 ncu --set full --clock-control none --import-source on -k regex:mkb_cell_step -s 4 -c 1 -f -o $O/prof_$w python scripts/profile_target.py $w 6 > $O/ncu_$w.log 2>&1
  tail -1 $O/ncu_$w.log
done
ls -la $O
