#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests/test_gpu_train.py -q > gpurun_out/aa_ops.log 2>&1; echo "rc=$?" >> gpurun_out/aa_ops.log
timeout 600 python tools/gpu_train_profile.py 5 > gpurun_out/aa_profile.txt 2> gpurun_out/aa_profile.err
timeout 600 python tools/gpu_train_bench.py 5 > gpurun_out/aa_train.json 2> gpurun_out/aa_train.err
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_t5.py tests/test_gpu_vae.py -q -x > gpurun_out/aa_regress.log 2>&1; echo "rc=$?" >> gpurun_out/aa_regress.log
tail -n 30 gpurun_out/aa_ops.log; head -24 gpurun_out/aa_profile.txt; tail -3 gpurun_out/aa_profile.err; cat gpurun_out/aa_train.json; tail -4 gpurun_out/aa_regress.log
