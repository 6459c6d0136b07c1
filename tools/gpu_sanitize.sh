#!/bin/bash
# compute-sanitizer gate: memcheck over one small invocation of every kernel family, synccheck and
# racecheck over the kernels with cross-warp protocols (tools/gpu_sanitize.py).  Summaries land in
# gpurun_out/sanitize_*.txt; copy them to profiles/ after reading.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool part
  timeout 420 $CS --tool $1 --print-limit 15 --launch-timeout 120 python tools/gpu_sanitize.py $2 > gpurun_out/sanitize_$1_$2.txt 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$1_$2.txt
}
for part in latte gemm attn t2v vae sampler; do run memcheck $part; done
for part in gemm attn latte; do run synccheck $part; done
for part in gemm attn; do run racecheck $part; done
grep -H -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|max err|finite|flags clean" gpurun_out/sanitize_*.txt | cut -c1-200
