"""Summarise `ncu -i <rep> --page raw --csv` dumps: per-kernel means of the metrics DESIGN.md quotes, and the per-launch DRAM
traffic file bench.py reads (profiles/ncu_traffic.json).

    python scripts/ncu_summary.py profiles/<name>_summary.json raw1.csv [raw2.csv ...]
"""
import collections
import csv
import json
import os
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread']
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3}

agg = collections.OrderedDict()
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
        d = agg.setdefault(name, collections.defaultdict(list))
        for w in WANT:
            if w in idx and r[idx[w]]:
                v = float(r[idx[w]].replace(',', ''))
                if w.startswith('dram__bytes') or w == 'gpu__time_duration.sum':
                    v *= SCALE.get(units[idx[w]], 1)
                d[w].append(v)
out = {}
for k, d in agg.items():
    out[k] = {'launches': len(d['gpu__time_duration.sum'])}
    for w in WANT:
        if d[w]:
            out[k][w + (' (us)' if w == 'gpu__time_duration.sum' else '')] = round(sum(d[w]) / len(d[w]), 3)
json.dump(out, open(sys.argv[1], 'w'), indent=1)
traffic = {k: v['dram__bytes_read.sum'] + v['dram__bytes_write.sum'] for k, v in out.items() if 'dram__bytes_read.sum' in v}
json.dump(traffic, open(os.path.join(os.path.dirname(sys.argv[1]), 'ncu_traffic.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))
