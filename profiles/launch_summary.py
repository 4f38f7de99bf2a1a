#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
seq = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = row['Kernel Name']
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1000 if unit == 'ns' else (v * 1000 if unit == 'ms' else v)
    grid = row.get('Grid Size', '')
    key = (name[:90], grid)
    agg.setdefault(key, []).append(v)
    seq.append((name[:60], grid, v))
tot = sum(sum(v) for v in agg.values())
print(f'{len(seq)} launches, {tot:.1f} us total')
for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f'{sum(v):10.1f} us {100 * sum(v) / tot:5.1f}%  n={len(v):4d} avg={sum(v) / len(v):8.2f} min={min(v):8.2f} max={max(v):8.2f}  grid={g:>14s} {k}')
if len(sys.argv) > 2:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for s in seq[a:b]:
        print(s)
