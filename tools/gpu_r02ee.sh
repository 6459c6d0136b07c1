#!/bin/bash
cd "$(dirname "$0")/.."
export LATTE_B200_NO_BUILD=1
for bn in 0 128 256; do echo "--- B200_WGRAD_BN=$bn"; B200_WGRAD_BN=$bn timeout 300 python tools/gpu_train_gemm_probe.py 2>&1 | grep wgrad; done
