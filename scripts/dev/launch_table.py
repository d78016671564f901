"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: total us, launches, us per launch.
    python scripts/dev/launch_table.py gpurun_out/xxx_launches.csv [top_n]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr, data = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    name = re.sub(r"\(.*", "", d["Kernel Name"])[:100]
    t = float(d["Metric Value"].replace(",", ""))
    t = t / 1000 if d["Metric Unit"] == "ns" else (t * 1000 if d["Metric Unit"] == "ms" else t)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
total = sum(t for _, t in agg.values())
print(f"{len(data)} launches, {total:.1f} us")
for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{t:11.1f} us {100 * t / total:5.1f}% {c:6d} x {t / c:9.1f} us  {name}")
