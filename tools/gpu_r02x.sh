#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/x_bench2.json 2> gpurun_out/x_bench2.err
python - <<'PY'
import json
r = json.loads(open("gpurun_out/x_bench2.json").read().strip().splitlines()[-1])
print("value", r["value"], "e2e", r["e2e"]["value"], "frames/s", r.get("frames_per_sec_e2e", {}).get("value"))
print("train", json.dumps(r.get("train_fwd_bwd"))[:900])
PY
tail -3 gpurun_out/x_bench2.err
