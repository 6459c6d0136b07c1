#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests/test_gpu_vae.py -q > gpurun_out/y_vae.log 2>&1; echo "rc=$?" >> gpurun_out/y_vae.log
timeout 300 python tools/gpu_vae_encode_bench.py 80 > gpurun_out/y_enc.json 2> gpurun_out/y_enc.err
tail -n 25 gpurun_out/y_vae.log; cat gpurun_out/y_enc.json; tail -3 gpurun_out/y_enc.err
