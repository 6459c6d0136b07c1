#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests/test_gpu_t5.py -m gpu -q > gpurun_out/k_pytest_t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest_t5.log
# launch list with the caches left as the step leaves them (no flush between kernels): in-step durations
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 620 -c 212 --csv --log-file gpurun_out/k_launches_warm.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-video > gpurun_out/k_ncu_launches.log 2>&1
grep -E "AssertionError|passed|failed" gpurun_out/k_pytest_t5.log | tail -8
python tools/launch_summary.py gpurun_out/k_launches_warm.csv | head -24
