#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests/test_gpu_train.py -q > gpurun_out/w_ops.log 2>&1; echo "rc=$?" >> gpurun_out/w_ops.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/w_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/w_smoke.log
tail -n 25 gpurun_out/w_ops.log; tail -4 gpurun_out/w_smoke.log
