#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q -k "attn_v3" > gpurun_out/d_pytest_v3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest_v3.log
timeout 1500 python -m pytest tests -m gpu -q -k "not attn_v3" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
timeout 300 python tools/gpu_microbench.py attn > gpurun_out/d_micro.txt 2>&1
B200_ATTN_IMPL=3 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-video > gpurun_out/d_bench_v3.json 2> gpurun_out/d_bench_v3.err
B200_ATTN_IMPL=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_v3 -s 6 -c 1 -f -o gpurun_out/d_prof_attn_spatial \
    python tools/gpu_microbench.py attn > gpurun_out/d_ncu_spatial.log 2>&1
tail -n 3 gpurun_out/d_pytest_v3.log; tail -n 6 gpurun_out/d_pytest.log; cat gpurun_out/d_micro.txt
python - <<'PY'
import json
for f in ("d_bench_v3",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["sustained"]["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], {k: d["roofline"][k] for k in ("gemm_ms_per_step","attn_ms_per_step","ln_ms_per_step","other_ms_per_step")})
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
