#!/bin/bash
# 8-GPU measurements of the BASELINE config shapes (one gpurun --gpus 8 call):
#   bench.py (config 2, weak scaling) + configs 3, 4a, 4b, 5 through tools/bench_configs.py
set -u
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
OUT=gpurun_out/scale8.log
: > $OUT
timeout 300 $TR --master-port 29501 bench.py --gpus 8 --steps 32 --warmup 8 2>&1 | tail -n 1 >> $OUT
for c in 3 4a 4b 5; do
  timeout 600 $TR --master-port 2951${c:0:1} tools/bench_configs.py --config $c --gpus 8 2>&1 | tail -n 1 >> $OUT
done
cut -c1-700 $OUT
