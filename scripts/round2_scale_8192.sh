#!/bin/bash
# Strong scaling on the grid the target names (north_star: ">= 85 % at 8 x B200 on the 8192^2 O'Hara-Rudy grid"):
# 8192 x 8192 decker-2009 fp64 Rush-Larsen = 67 M cells, 25.8 GB of state. Round 1 measured scaling on the 2048^2
# bench workload only (83 % at 8 GPUs: 256-row slabs) and on the 8192^2 fp32 LR1991 grid (89 %).
#   gpurun --gpus 8 --timeout 600 -- 'bash scripts/round2_scale_8192.sh'      (about 3 minutes of box time x 8)
O=gpurun_out/scale8192
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --grid 8192 --steps 20 --warmup 3 --no-cpu 2>$O/n8.err | tail -1 > $O/bench_8192_n8.json
cut -c1-200 $O/bench_8192_n8.json
timeout 300 python bench.py --gpus 1 --grid 8192 --steps 20 --warmup 3 --no-cpu 2>$O/n1.err | tail -1 > $O/bench_8192_n1.json
cut -c1-200 $O/bench_8192_n1.json
python - <<'PY'
import json
a = json.load(open('gpurun_out/scale8192/bench_8192_n1.json'))
b = json.load(open('gpurun_out/scale8192/bench_8192_n8.json'))
print('8192^2: 1 GPU %.3e, 8 GPUs %.3e cell-steps/s -> speed-up %.2f, efficiency %.1f %%'
      % (a['value'], b['value'], b['value'] / a['value'], 100 * b['value'] / a['value'] / 8))
PY
