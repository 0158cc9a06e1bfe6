#!/bin/bash
# runs every diagnostic stage in its own process (a faulting stage cannot poison the next one)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/debug.log 2>&1
for st in pool kmeans eig affinity_bf16 affinity_f32 e2e time; do
  echo "=== $st" >> gpurun_out/debug.log
  timeout 240 python tools/gpu_debug.py $st >> gpurun_out/debug.log 2>&1
  echo "exit $?" >> gpurun_out/debug.log
done
tail -c 6000 gpurun_out/debug.log
