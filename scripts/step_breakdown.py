"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of ONE training step
(the launches between two consecutive codes_kernel launches = one forward+loss+backward+optimizer)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[h]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
launches = [(r[ki].split('(')[0].replace('<unnamed>::', '').replace('void ', '')[:70], float(r[vi].replace(',', '')))
            for r in rows[h + 1:] if len(r) > vi]
starts = [i for i, (n, _) in enumerate(launches) if n.startswith('codes')]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
seg = launches[starts[which]:starts[which + 1]]
agg = collections.OrderedDict()
for n, v in seg:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in seg)
print(f"one step: {len(seg)} launches, {tot / 1e6:.3f} ms (cold-cache, serialised)")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v / 1e6:8.3f} ms {100 * v / tot:5.1f}%  n={n:3d}  {k}")
