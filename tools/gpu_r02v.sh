#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
tail -n 6 gpurun_out/v_pytest.log; python - <<'PY'
import json
r = json.loads(open("gpurun_out/v_bench.json").read().strip().splitlines()[-1])
print("value", r["value"], "e2e", r["e2e"]["value"], "frames/s", r.get("frames_per_sec_e2e", {}).get("value"))
print("train", json.dumps(r.get("train_fwd_bwd"))[:1500])
PY
tail -3 gpurun_out/v_bench.err
