#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 600 python -m pytest tests/test_gpu_t2v.py tests/test_gpu_ops.py -m gpu -q -k "t2v or ln_modulate" > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
timeout 300 python tools/gpu_hostcost.py > gpurun_out/b_hostcost.txt 2>&1
timeout 300 python tools/gpu_microbench.py ln > gpurun_out/b_micro.txt 2>&1
# full ncu capture of the v3 attention kernels (spatial 256-key and temporal) from the microbenchmark
B200_ATTN_IMPL=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_v3 -s 6 -c 1 -f -o gpurun_out/b_prof_attn_spatial \
    python tools/gpu_microbench.py attn > gpurun_out/b_ncu_spatial.log 2>&1
B200_ATTN_IMPL=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_v3 -s 120 -c 1 -f -o gpurun_out/b_prof_attn_temporal \
    python tools/gpu_microbench.py attn > gpurun_out/b_ncu_temporal.log 2>&1
tail -n 5 gpurun_out/b_pytest.log; cat gpurun_out/b_hostcost.txt gpurun_out/b_micro.txt; tail -3 gpurun_out/b_ncu_spatial.log; ls -la gpurun_out | grep b_prof
