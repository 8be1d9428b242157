#!/bin/bash
# Runs on the GPU box (gpurun): numbers and ncu captures for profiles/.
set -x
O=gpurun_out
mkdir -p $O
python -c "
import sys, json; sys.path.insert(0, '.')
import myokit_b200
from myokit_b200 import capi
print(json.dumps(dict(capi.measure_peaks(0), device=capi.device_info(0))))" 2>/dev/null | tail -1 > $O/pipe_peaks.json
cat $O/pipe_peaks.json
python bench.py --impl reference --steps 50 --warmup 3 2>/dev/null | tail -1 > $O/bench_ref.json
python bench.py --steps 200 --warmup 5 2>/dev/null | tail -1 > $O/bench_n1.json
cut -c1-400 $O/bench_n1.json
python scripts/bench_configs.py c1 c2 c4 c5 2>/dev/null | grep -v Warn > $O/configs.txt
cat $O/configs.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
for w in c3 stencil32 stencil64 lr91_fp32; do
  ncu --set full --clock-control none --import-source on -k regex:mkb_cell_step -s 4 -c 1 -o $O/prof_$w python scripts/profile_target.py $w 6 > $O/ncu_$w.log 2>&1
done
ls -la $O
