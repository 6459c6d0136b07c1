"""profiles/gemm_traffic.json from an `ncu --set full` capture of the GEMM launches of one block (run here, no GPU):
python tools/ncu_traffic.py gpurun_out/prof_gemm.ncu-rep  ->  mean dram__bytes_read.sum + dram__bytes_write.sum per launch."""
import csv
import json
import os
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ir, iw, it, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum"), hdr.index("Kernel Name")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = []
for r in rows[2:]:
    if "gemm_kernel" not in r[ik]:
        continue
    b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    per.append({"kernel": r[ik].split("gemm_kernel")[1][:24], "dram_bytes": b, "time_us_under_ncu": float(r[it])})
res = {"dram_bytes_per_launch": sum(p["dram_bytes"] for p in per) / len(per), "launches": per,
       "source": f"ncu --set full, {os.path.basename(rep)} ({len(per)} consecutive gemm_kernel launches = one transformer block; cold L2 per replay)"}
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "gemm_traffic.json")
json.dump(res, open(dst, "w"), indent=1)
print(json.dumps(res, indent=1))
