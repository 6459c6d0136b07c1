#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests/test_gpu_train.py -q > gpurun_out/p_ops.log 2>&1; echo "rc=$?" >> gpurun_out/p_ops.log
timeout 600 python tools/gpu_train_profile.py 5 > gpurun_out/p_profile.txt 2> gpurun_out/p_profile.err
tail -n 30 gpurun_out/p_ops.log; cat gpurun_out/p_profile.txt; tail -5 gpurun_out/p_profile.err
