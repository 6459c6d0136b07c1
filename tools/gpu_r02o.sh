#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests/test_gpu_train.py -q -x --deselect tests/test_gpu_train.py::test_training_step_xl_head_geometry > gpurun_out/o_ops.log 2>&1; echo "rc=$?" >> gpurun_out/o_ops.log
timeout 600 python -m pytest tests/test_gpu_train.py -q -k "xl_head or reference_gradients" > gpurun_out/o_engine.log 2>&1; echo "rc=$?" >> gpurun_out/o_engine.log
timeout 600 python tools/gpu_train_bench.py 5 > gpurun_out/o_train.json 2> gpurun_out/o_train.err
tail -n 25 gpurun_out/o_ops.log; tail -n 25 gpurun_out/o_engine.log; cat gpurun_out/o_train.json; tail -5 gpurun_out/o_train.err
