#!/usr/bin/env python
"""Opcode mix and stall-sample share of one kernel from `ncu -i R.ncu-rep --page source --csv --print-source sass`.
usage: sass_mix.py file.csv n_warps_times_steps"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
norm = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = next(r for r in rows if 'Source' in r and '# Samples' in r)
ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
byop, byopx = collections.Counter(), collections.Counter()
lines = []
for r in rows:
    if len(r) <= max(isamp, iex) or not r[isamp].isdigit():
        continue
    toks = r[ia].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    parts = op.split('.')
    op = '.'.join(parts[:2]) if parts[0] in ('LDS', 'STS', 'LDG', 'STG', 'BAR', 'MUFU') else parts[0]
    byop[op] += int(r[isamp])
    byopx[op] += int(r[iex])
    lines.append((int(r[isamp]), r[ia].strip(), int(r[iex])))
tot, totx = sum(byop.values()), sum(byopx.values())
print('total samples %d, warp instructions %d (%.0f per unit)' % (tot, totx, totx / norm))
for op, c in byop.most_common(28):
    print('%-14s samples %5.1f%%   inst %5.1f%% (%.1f per unit)' % (op, 100 * c / tot, 100 * byopx[op] / totx, byopx[op] / norm))
print('hottest instructions:')
for s, src, ex in sorted(lines, reverse=True)[:25]:
    print('  %5.2f%%  %s' % (100 * s / tot, src[:110]))
