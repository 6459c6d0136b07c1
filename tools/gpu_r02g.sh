#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 600 python tools/gpu_microbench.py gemm > gpurun_out/g_micro_gemm.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
timeout 600 python bench.py --workload t2v --steps 5 --warmup 3 > gpurun_out/g_bench_t2v.json 2> gpurun_out/g_bench_t2v.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 620 -c 212 --csv --log-file gpurun_out/g_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-video > gpurun_out/g_ncu_launches.log 2>&1
grep -E "cublas|bn256|bn192" gpurun_out/g_micro_gemm.txt | grep cold
python - <<'PY'
import json
for f in ("g_bench", "g_bench_t2v"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(json.dumps(d)[:3500])
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
python tools/launch_summary.py gpurun_out/g_launches.csv 2>/dev/null | head -30
