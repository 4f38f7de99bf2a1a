#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export: executed warp-instructions and stall samples
by opcode, plus the hottest instructions.  One section per kernel in the export.
usage: src_hot.py file.csv [top_n] [kernel-substring]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ''
sections, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        sections.append(cur)
    elif cur is not None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and len(r) >= len(cur['hdr']):
        cur['rows'].append(r)
for sec in sections:
    if want not in sec['name']:
        continue
    ix = {h: i for i, h in enumerate(sec['hdr'])}
    by_op, samp_op = collections.Counter(), collections.Counter()
    tot = tots = 0
    items = []
    for k, r in enumerate(sec['rows']):
        src = r[ix['Source']].strip()
        parts = src.split()
        op = parts[0] if not src.startswith('@') else parts[1]
        op = op.split('.')[0]
        n = int(r[ix['Instructions Executed']] or 0)
        s = int(r[ix['# Samples']] or 0)
        by_op[op] += n
        samp_op[op] += s
        tot += n
        tots += s
        items.append((s, n, k, src))
    print('=====', sec['name'][:120])
    print('total warp-instructions', tot, 'samples', tots)
    for op, n in by_op.most_common(22):
        print(f'{op:12s} {n:10d} {100 * n / max(tot, 1):5.1f}%   samples {samp_op[op]:6d} {100 * samp_op[op] / max(tots, 1):5.1f}%')
    print('--- hottest by samples')
    for s, n, k, src in sorted(items, reverse=True)[:top]:
        print(f'{s:6d} {n:8d} #{k:4d} {src[:110]}')
