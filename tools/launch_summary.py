"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel, share of the step."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
seq = []
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("b200::<unnamed>::", "").replace("<unnamed>::", "")
    t = float(r[vi].replace(",", ""))
    t = t / 1e3 if r[ui].startswith("ns") else t  # -> us
    seq.append((name, t))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(t for _, t in seq)
print(f"total {tot:.1f} us over {len(seq)} launches (cold-cache, serialised: compare SHARES)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {100 * t / tot:5.1f}%  n={n:3d} avg {t / n:7.1f} us  {k}")
