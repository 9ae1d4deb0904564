#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of one training step
(the launches between two consecutive inc_step_kernel launches) and the setup (prepare_data) kernels."""
import collections
import csv
import sys


def load(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(lines[start:]))


def main(path):
    rows = load(path)
    names = [r['Kernel Name'].split('(')[0] for r in rows]
    idx = [i for i, n in enumerate(names) if n == 'inc_step_kernel']
    out = []
    if len(idx) >= 2:
        seg = rows[idx[-2]:idx[-1]]
        agg = collections.OrderedDict()
        for r in seg:
            a = agg.setdefault(r['Kernel Name'].split('(')[0], [0, 0.0])
            a[0] += 1
            a[1] += float(r['Metric Value'])
        tot = sum(v[1] for v in agg.values())
        out.append('one training step: %d launches, %.1f us summed kernel time (ncu: serialised, cold caches)' % (len(seg), tot / 1e3))
        out.append('%-44s %5s %12s %7s' % ('kernel', 'calls', 'us', 'share'))
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append('%-44s %5d %12.1f %6.1f%%' % (n, c, t / 1e3, 100 * t / tot))
    pre = rows[:idx[0]] if idx else rows
    agg2 = collections.OrderedDict()
    for r in pre:
        a = agg2.setdefault(r['Kernel Name'].split('(')[0], [0, 0.0])
        a[0] += 1
        a[1] += float(r['Metric Value'])
    out.append('')
    out.append('setup phase (prepare_data + table build): %d launches' % len(pre))
    for n, (c, t) in sorted(agg2.items(), key=lambda kv: -kv[1][1])[:16]:
        out.append('%-60s %5d %12.1f us' % (n[:60], c, t / 1e3))
    print('\n'.join(out))


if __name__ == '__main__':
    main(sys.argv[1])
