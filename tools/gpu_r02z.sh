#!/bin/bash
# round-2 additions under compute-sanitizer + an ncu launch list of one training step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool part
  timeout 400 $CS --tool $1 --print-limit 15 --launch-timeout 120 python tools/gpu_sanitize.py $2 > gpurun_out/sanitize_$1_$2.txt 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$1_$2.txt
}
for part in train train72 vae_enc; do run memcheck $part; done
for part in train72; do run racecheck $part; run synccheck $part; done
grep -H -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|max err|finite" gpurun_out/sanitize_*_train*.txt gpurun_out/sanitize_*_vae_enc.txt | cut -c1-220
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/z_train_launches.csv python tools/gpu_train_bench.py 5 > gpurun_out/z_ncu_train.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/z_train_launches.csv", errors="ignore")) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    name = r[ki].split("(")[0][:90]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)-1} launches captured, {tot/1e6:.2f} ms of kernel time")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t/1e6:9.3f} ms {n:6d} x {k}")
PY
