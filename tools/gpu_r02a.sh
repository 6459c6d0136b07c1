#!/bin/bash
# round 2, GPU call A: parity of everything new (graph capture, stream-K flags, T2V goldens + masks, both attention kernels),
# microbenchmarks of attention v2 vs v3 and LN, bench lines with either attention kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
# v3 tests in their own process: a trap in the new kernel must not poison the rest of the suite
timeout 900 python -m pytest tests -m gpu -q -k "attn_v3" > gpurun_out/a_pytest_v3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest_v3.log
timeout 1500 python -m pytest tests -m gpu -q -k "not attn_v3" > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 300 python tools/gpu_microbench.py attn ln > gpurun_out/a_micro.txt 2>&1
B200_ATTN_DBG=16 timeout 200 python tools/gpu_microbench.py attn > gpurun_out/a_micro_nopoly.txt 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/a_bench_v2.json 2> gpurun_out/a_bench_v2.err
B200_ATTN_IMPL=3 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-video > gpurun_out/a_bench_v3.json 2> gpurun_out/a_bench_v3.err
tail -n 4 gpurun_out/a_pytest_v3.log; tail -n 4 gpurun_out/a_pytest.log; cat gpurun_out/a_micro.txt; cat gpurun_out/a_micro_nopoly.txt | grep v3
python - <<'PY'
import json
for f in ("a_bench_v2", "a_bench_v3"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], {k: d["roofline"][k] for k in ("gemm_ms_per_step","attn_ms_per_step","ln_ms_per_step","other_ms_per_step")}, d.get("gpu_eager_baseline"), d.get("frames_per_sec_e2e"))
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
