#!/bin/bash
# Runs every bring-up stage in its own process under `timeout` (a deadlocked kernel must not hang the box).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/smi.txt
for st in "${@:-gemm attn sample forward}"; do
  for s in $st; do
    echo "######## $s"
    timeout 240 python tools/gpu_check.py $s 2>&1 | tee gpurun_out/check_$s.log | tail -60
    echo "exit: ${PIPESTATUS[0]}"
  done
done
