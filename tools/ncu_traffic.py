#!/usr/bin/env python
"""ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv)
-> JSON: per kernel name the launches, total duration and DRAM bytes of the captured step.

    python tools/ncu_traffic.py gpurun_out/X_traffic.csv "1024^3 step" 1073741824 > profiles/X_traffic.json
"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]
kn, mn, mu, mv, idc = (hdr.index(c) for c in ('Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value', 'ID'))
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12,
        'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6}
per = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r'^void ', '', r[kn]).split('(')[0]
    name = re.sub(r'<.*$', '', name)
    d = per.setdefault((int(r[idc]), name), {})
    d[r[mn]] = float(r[mv].replace(',', '')) * UNIT.get(r[mu], 1.0)
fam = collections.OrderedDict()
for (_, name), d in per.items():
    f = fam.setdefault(name, {"launches": 0, "duration_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
    f["launches"] += 1
    f["duration_us"] += d.get('gpu__time_duration.sum', 0.0)
    f["dram_read_bytes"] += d.get('dram__bytes_read.sum', 0.0)
    f["dram_write_bytes"] += d.get('dram__bytes_write.sum', 0.0)
out = {"what": sys.argv[2] if len(sys.argv) > 2 else "",
       "voxels": int(sys.argv[3]) if len(sys.argv) > 3 else None, "source": sys.argv[1].split('/')[-1],
       "note": "ncu replays kernels with cold caches and serialised; durations are for shares, not for throughput",
       "kernels": fam}
print(json.dumps(out, indent=1))
