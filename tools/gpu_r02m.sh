#!/bin/bash
# round-2 profiling evidence: ncu --set full captures of the dominant kernels (read here with tools/ncu_summary.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:gemm_kernel -s 230 -c 4 -o gpurun_out/m_prof_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-video > gpurun_out/m_ncu_gemm.log 2>&1
timeout 600 $NCU -k regex:attn_v3 -s 60 -c 2 -o gpurun_out/m_prof_attn python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-video > gpurun_out/m_ncu_attn.log 2>&1
timeout 600 $NCU -k regex:ln_modulate -s 120 -c 2 -o gpurun_out/m_prof_ln python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-video > gpurun_out/m_ncu_ln.log 2>&1
B200_MB_IMPLS=3 timeout 300 $NCU -k regex:attn_stream -s 3 -c 1 -o gpurun_out/m_prof_attn_stream python tools/gpu_microbench.py attn_long > gpurun_out/m_ncu_stream.log 2>&1
timeout 300 python tools/gpu_vae_bench.py > gpurun_out/m_vae.json 2> gpurun_out/m_vae.err
timeout 600 $NCU -k regex:"gemm_kernel|gn_" -s 150 -c 6 -o gpurun_out/m_prof_vae python tools/gpu_vae_bench.py > gpurun_out/m_ncu_vae.log 2>&1
cat gpurun_out/m_vae.json; ls -la gpurun_out | grep m_prof
