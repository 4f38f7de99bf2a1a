#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export: executed warp-instructions and stall samples
by opcode, plus the hottest instructions."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
by_op = collections.Counter()
samp_op = collections.Counter()
tot = tots = 0
items = []
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    src = r[ix['Source']].strip()
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    op = op.split('.')[0]
    n = int(r[ix['Instructions Executed']] or 0)
    s = int(r[ix['# Samples']] or 0)
    by_op[op] += n
    samp_op[op] += s
    tot += n
    tots += s
    items.append((s, n, k, src))
print('total warp-instructions', tot, 'samples', tots)
for op, n in by_op.most_common(22):
    print(f'{op:12s} {n:10d} {100 * n / tot:5.1f}%   samples {samp_op[op]:6d} {100 * samp_op[op] / max(tots, 1):5.1f}%')
print('--- hottest by samples')
for s, n, k, src in sorted(items, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f'{s:6d} {n:8d} #{k:4d} {src[:110]}')
