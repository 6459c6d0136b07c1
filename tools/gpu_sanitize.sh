#!/bin/bash
# compute-sanitizer gate: memcheck, synccheck and racecheck over one small invocation of every kernel
# family (tools/gpu_sanitize.py).  Summaries land in gpurun_out/sanitize_*.txt; copy them to profiles/ after reading.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
  for part in latte gemm attn t2v vae sampler; do
    timeout 600 $CS --tool $tool --print-limit 20 --launch-timeout 120 python tools/gpu_sanitize.py $part \
      > gpurun_out/sanitize_${tool}_${part}.txt 2>&1
    echo "rc=$?" >> gpurun_out/sanitize_${tool}_${part}.txt
  done
done
grep -H -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|max err|Error:|hazard" gpurun_out/sanitize_*.txt | sort | uniq -c | sort -rn | head -80
