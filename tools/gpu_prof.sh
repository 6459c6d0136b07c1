#!/bin/bash
# ncu captures: launch list of one step + full sets for the LN / attention / GEMM kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
ncu --metrics gpu__time_duration.sum --clock-control none -s 203 -c 215 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ln_modulate|attn_kernel" -s 4 -c 4 -f -o gpurun_out/prof_ln_attn \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ln.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 8 -c 4 -f -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
