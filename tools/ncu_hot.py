#!/usr/bin/env python
"""Top stall-sample SASS lines of one kernel from an .ncu-rep source page:  ncu_hot.py rep regex [n] [instance]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
inst = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == 'Kernel Name':
        cur = {'name': row[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(row)
b = blocks[inst]
hdr = b['rows'][0]
rows = b['rows'][1:]
si, so = hdr.index('# Samples'), hdr.index('Source')
ie = hdr.index('Instructions Executed')
tot = sum(int(r[si] or 0) for r in rows)
print(b['name'][:100], 'total samples', tot, 'sass lines', len(rows))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i][si] or 0))[:n]
for i in sorted(idx):
    r = rows[i]
    print('%5d %6.1f%%  exec=%-8s %s' % (i, 100.0 * int(r[si] or 0) / max(tot, 1), r[ie], r[so].strip()[:110]))
