#!/bin/bash
# Runs on an 8-GPU box (gpurun --gpus 8): multi-GPU tests and the scaling numbers
# for profiles/r01_scaling.md. Most valuable first; every command has its own limit.
O=gpurun_out/scale
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 120 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu > $O/tests.log 2>&1
tail -2 $O/tests.log
port=29600
for n in 8 4 2; do
  port=$((port + 1))
  timeout 150 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 200 --warmup 5 2>$O/bench_n$n.err | tail -1 > $O/bench_n$n.json
  cut -c1-160 $O/bench_n$n.json
  if [ $n != 2 ]; then
    port=$((port + 1))
    timeout 150 $TR --nproc-per-node $n --master-port $port scripts/bench_configs.py c4 mesh c5 2>$O/configs_n$n.err | grep "GPUs\]" > $O/configs_n$n.txt
    cat $O/configs_n$n.txt
  fi
done
