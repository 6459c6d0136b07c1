#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests/test_gpu_train.py -q > gpurun_out/dd_ops.log 2>&1; echo "rc=$?" >> gpurun_out/dd_ops.log
timeout 600 python tools/gpu_train_bench.py 5 > gpurun_out/dd_train.json 2> gpurun_out/dd_train.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/dd_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/dd_smoke.log
tail -n 12 gpurun_out/dd_ops.log; cat gpurun_out/dd_train.json; tail -3 gpurun_out/dd_smoke.log
