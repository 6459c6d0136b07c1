#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 1500 python -m pytest tests -m gpu -q -k "model or t2v or t5 or graph or linear or sampler" > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
for i in 1 2; do
B200_GEMM_NO_WPREFETCH=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-video > gpurun_out/l_bench_nopre_$i.json 2> gpurun_out/l_bench_nopre_$i.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-video > gpurun_out/l_bench_pre_$i.json 2> gpurun_out/l_bench_pre_$i.err
done
tail -n 4 gpurun_out/l_pytest.log
python - <<'PY'
import json
for f in ("l_bench_nopre_1", "l_bench_pre_1", "l_bench_nopre_2", "l_bench_pre_2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["sustained"]["ms_per_step"],3), round(d["e2e"]["value"],2), {k: round(d["roofline"][k],3) for k in ("gemm_ms_per_step","attn_ms_per_step","ln_ms_per_step","other_ms_per_step")})
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-800:])
PY
