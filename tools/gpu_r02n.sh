#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
timeout 300 python tools/gpu_vae_bench.py > gpurun_out/n_vae.json 2> gpurun_out/n_vae.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
tail -n 5 gpurun_out/n_pytest.log; cat gpurun_out/n_vae.json; tail -2 gpurun_out/n_smoke.log; cut -c1-3000 gpurun_out/n_bench.json; tail -3 gpurun_out/n_bench.err
