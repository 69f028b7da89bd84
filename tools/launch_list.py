#!/usr/bin/env python
"""Print an ncu `--metrics gpu__time_duration.sum --csv` launch list compactly.
    python tools/launch_list.py gpurun_out/launches.csv [tail N]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
seq = []
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(',', ''))
    v = v / 1e3 if r[mu] in ('ns', 'nsecond') else v * 1e3 if r[mu] in ('ms', 'msecond') else v
    seq.append((r[kn][:70], v))
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(seq)
for s in seq[-n:]:
    print("%-70s %10.1f us" % s)
tot = defaultdict(lambda: [0, 0.0])
for k, v in seq:
    tot[k][0] += 1
    tot[k][1] += v
print("---- totals over %d launches" % len(seq))
T = sum(v for _, v in seq)
for k, (c, v) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print("%-70s x%-4d %10.1f us  avg %8.1f  share %.3f" % (k, c, v, v / c, v / T))
