#!/bin/bash
# run every bring-up stage in its own process (a trap in one kernel must not poison the next stage)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/diag.log 2>&1
for s in ${@:-ln gemm_probe gemm attn_probe attn model xl}; do
  echo "=== stage $s" >> gpurun_out/diag.log
  timeout 300 python tools/gpu_diag.py $s >> gpurun_out/diag.log 2>&1
  echo "=== stage $s exit $?" >> gpurun_out/diag.log
done
tail -n 150 gpurun_out/diag.log
