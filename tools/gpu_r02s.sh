#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python tools/gpu_train_micro.py > gpurun_out/s_micro.txt 2> gpurun_out/s_micro.err
cat gpurun_out/s_micro.txt; tail -3 gpurun_out/s_micro.err
