#!/bin/bash
# Runs on an N-GPU box (gpurun --gpus N): multi-GPU tests on distinct devices, the bench at N (and below), slab overhead.
N=${1:-2}
ONLY=${2:-"1 2 4 8"}   # which GPU counts to bench
O=gpurun_out/${OUT:-r02}
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_multigpu_gpu.py -q -m gpu 2>&1 | tail -3 | tee $O/multigpu_tests_n$N.log
port=29700
for n in $(seq 1 8); do
  if [ $n -le $N ] && [[ " $ONLY " == *" $n "* ]]; then
    port=$((port + 1))
    if [ $n = 1 ]; then
      timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu 2>$O/scale_n$n.err | tail -1 > $O/scale_n$n.json
    else
      timeout 400 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 20 --warmup 3 2>$O/scale_n$n.err | tail -1 > $O/scale_n$n.json
    fi
    python - <<PY
import json
try:
    d = json.load(open('$O/scale_n$n.json'))
    b = d.get('scaling_8192') or {}
    print('N=$n: %.4f ms/step %.3e cs/s | e2e %.3e resident %.3e | 8192^2 %.3f ms %.3e | check %s' % (
        d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['resident']['value'],
        b.get('ms_per_step', 0), b.get('value', 0), d.get('sharded_check')))
except Exception as e:
    print('N=$n failed', e)
    print(open('$O/scale_n$n.err').read()[-1500:])
PY
  fi
done
# what a slab costs beyond its cells: the same rows unsharded on one GPU
[ -n "$SKIP_SINGLE" ] || python - <<'PY' 2>&1 | grep -v Warn
import sys; sys.path.insert(0, '.')
import myokit_b200
from myokit_b200 import workloads
for ny in (1024, 512, 256):
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=2048, ny=ny)
    i = s.benchmark_steps(40, warmup=10)
    print('single GPU, 2048 x %d unsharded: %.4f ms/step' % (ny, i['device_ms'] / i['steps']))
PY
# the other BASELINE configurations, sharded (device-resident timing)
if [ -n "$WITH_CONFIGS" ]; then
  port=$((port + 1))
  timeout 300 $TR --nproc-per-node $N --master-port $port scripts/bench_configs.py c4 mesh c5 2>$O/configs_n$N.err | grep "GPUs\]" | tee $O/configs_n$N.txt
  timeout 300 python scripts/bench_configs.py c4 mesh c5 2>/dev/null | grep -v Warn | grep "cell-steps" | tee $O/configs_n1.txt
fi
