#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "attention" > gpurun_out/i_pytest_attn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest_attn.log
timeout 900 python -m pytest tests/test_gpu_t2v.py -m gpu -q > gpurun_out/i_pytest_t2v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest_t2v.log
timeout 300 python tools/gpu_microbench.py attn_long > gpurun_out/i_micro_long.txt 2>&1
timeout 600 python bench.py --workload t2v --steps 5 --warmup 3 > gpurun_out/i_bench_t2v.json 2> gpurun_out/i_bench_t2v.err
tail -n 8 gpurun_out/i_pytest_attn.log; tail -n 4 gpurun_out/i_pytest_t2v.log; cat gpurun_out/i_micro_long.txt; cut -c1-1500 gpurun_out/i_bench_t2v.json; tail -3 gpurun_out/i_bench_t2v.err
