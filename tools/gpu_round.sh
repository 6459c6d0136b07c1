#!/bin/bash
# one GPU call: parity tests, bench line, ncu launch list + full captures of the two tensor-core kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "$1" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -s 203 -c 215 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-video > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 8 -c 4 -f -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-video > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attn_|ln_modulate" -s 4 -c 4 -f -o gpurun_out/prof_attn_ln \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-video > gpurun_out/ncu_attn.log 2>&1
fi
tail -n 5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
