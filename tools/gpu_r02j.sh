#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests/test_gpu_t5.py -m gpu -q > gpurun_out/j_pytest_t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest_t5.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_t5.py > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-video > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
tail -n 12 gpurun_out/j_pytest_t5.log; tail -n 5 gpurun_out/j_pytest.log
python - <<'PY'
import json
for f in ("j_bench",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["sustained"]["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], {k: d["roofline"][k] for k in ("gemm_ms_per_step","attn_ms_per_step","ln_ms_per_step","other_ms_per_step")})
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
