#!/bin/bash
# engine-isolation experiment on attention v3 (timing only; results are wrong for dbg != 0, 16)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1 B200_MB_IMPLS=3
: > gpurun_out/e_dbg.txt
for d in 0 1 8 64 72 32 2 128 4 256 260 262 335 16; do
  echo "== B200_ATTN_DBG=$d" >> gpurun_out/e_dbg.txt
  B200_ATTN_DBG=$d timeout 120 python tools/gpu_microbench.py attn 2>&1 | grep cold >> gpurun_out/e_dbg.txt
done
cat gpurun_out/e_dbg.txt
