#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests/tesu_gpu_train.py -q > gpurun_out/u_ops.log 2>&1; echo "rc=$?" >> gpurun_out/u_ops.log
timeout 300 python tools/gpu_train_micro.py worker > gpurun_out/u_micro.txt 2>&1
timeout 600 python tools/gpu_train_profile.py 5 > gpurun_out/u_profile.txt 2> gpurun_out/u_profile.err
timeout 600 python tools/gpu_train_bench.py 5 > gpurun_out/u_train.json 2> gpurun_out/u_train.err
timeout 600 python -m pytest tests/tesu_gpu_ops.py tests/tesu_gpu_model.py -q -x > gpurun_out/u_regress.log 2>&1; echo "rc=$?" >> gpurun_out/u_regress.log
tail -n 30 gpurun_out/u_ops.log; cat gpurun_out/u_micro.txt; cat gpurun_out/u_profile.txt; tail -3 gpurun_out/u_profile.err; cat gpurun_out/u_train.json; tail -4 gpurun_out/u_regress.log
