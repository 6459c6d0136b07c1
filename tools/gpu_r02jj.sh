#!/bin/bash
cd "$(dirname "$0")/.."
export LATTE_B200_NO_BUILD=1
echo "separate:"; timeout 120 python tools/gpu_train_bench.py 5 2>/dev/null | cut -c1-140
echo "fused (attention half):"; LATTE_B200_FUSED_RESIDUAL_LN=1 timeout 120 python tools/gpu_train_bench.py 5 2>/dev/null | cut -c1-140
echo "separate again:"; timeout 120 python tools/gpu_train_bench.py 5 2>/dev/null | cut -c1-140
