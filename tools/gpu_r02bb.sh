#!/bin/bash
# final validation of round 2: whole GPU suite, smoke, bench line (all legs), T2V workload, backward-GEMM probe + ncu --set full of it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/bb_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/bb_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/bb_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/bb_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bb_bench.json 2> gpurun_out/bb_bench.err
timeout 300 python tools/gpu_train_gemm_probe.py > gpurun_out/bb_gemm_probe.txt 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gemm_kernel -s 2 -c 4 -o gpurun_out/bb_prof_train_gemm -f python tools/gpu_train_gemm_probe.py > gpurun_out/bb_ncu.log 2>&1
ncu -i gpurun_out/bb_prof_train_gemm.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY' > gpurun_out/bb_train_gemm_ncu.txt
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) > 2:
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum"]
    idx = [hdr.index(w) for w in want if w in hdr]
    for r in rows[2:]:
        print(" | ".join(f"{hdr[i]}={r[i][:70]}" for i in idx))
PY
timeout 600 python bench.py --workload t2v --steps 5 --warmup 3 > gpurun_out/bb_t2v.json 2> gpurun_out/bb_t2v.err
tail -n 5 gpurun_out/bb_pytest.log; tail -3 gpurun_out/bb_smoke.log; cat gpurun_out/bb_gemm_probe.txt; cat gpurun_out/bb_train_gemm_ncu.txt
python - <<'PY'
import json
r = json.loads(open("gpurun_out/bb_bench.json").read().strip().splitlines()[-1])
print("value", r["value"], "e2e", r["e2e"]["value"], "sustained", r["sustained"]["value"], "frames/s", r.get("frames_per_sec_e2e", {}).get("value"), "roofline", r["roofline"]["frac"])
print("train", json.dumps(r.get("train_fwd_bwd"))[:1400])
print("eager", json.dumps(r.get("gpu_eager_baseline"))[:300])
try:
    t = json.loads(open("gpurun_out/bb_t2v.json").read().strip().splitlines()[-1]); print("t2v", t.get("ms_per_step"), t.get("value"))
except Exception as e: print("t2v parse", e)
PY
tail -2 gpurun_out/bb_bench.err
