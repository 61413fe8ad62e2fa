#!/usr/bin/env python
"""Print the kernels of the LAST bench step from an ncu gpu__time_duration launch list (csv)."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[1:]
names = [r[ix['Kernel Name']] for r in data]
starts = [i for i, n in enumerate(names) if 'revcomp' in n]
last = starts[-1]
tot = 0; agg = {}
small = 0
for r in data[last:]:
    v = float(r[ix['Metric Value']]) / 1000; tot += v
    nm = r[ix['Kernel Name']].split('(')[0].replace('void ', '')
    if v >= float(sys.argv[2]) if len(sys.argv) > 2 else 20:
        print(f"{v:9.1f} us  {nm[:90]}")
    else:
        small += v
    agg[nm] = agg.get(nm, 0) + v
print(f"{small:9.1f} us  (all launches below the threshold)")
print('launches', len(data) - last, 'total us', round(tot, 1))
