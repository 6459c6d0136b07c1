#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q -k "attn_v3 or test_gpu_model or t2v" > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
B200_MB_IMPLS=3 timeout 200 python tools/gpu_microbench.py attn > gpurun_out/f_micro.txt 2>&1
B200_MB_IMPLS=3 B200_ATTN_DBG=512 timeout 200 python tools/gpu_microbench.py attn > gpurun_out/f_micro_strided.txt 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-video > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
tail -n 3 gpurun_out/f_pytest.log; cat gpurun_out/f_micro.txt; echo strided; cat gpurun_out/f_micro_strided.txt
python - <<'PY'
import json
for f in ("f_bench",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["sustained"]["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], {k: d["roofline"][k] for k in ("gemm_ms_per_step","attn_ms_per_step","ln_ms_per_step","other_ms_per_step")})
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
