#!/usr/bin/env python
"""Per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch, mean over the launches of one
captured step) and duration from an `ncu --set full` report -> JSON for profiles/ (read by bench.py's roofline.traffic).

    python tools/ncu_traffic.py gpurun_out/full.ncu-rep > profiles/r02_traffic_<workload>.json

The table is stamped with the sha256 of every file under subgnn_b200/csrc: bench.py refuses it for an entry point whose source
files have changed since the capture."""
import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
TIME = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'second': 1e6, 's': 1e6}


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    iN, iR, iW, iT = (hdr.index(k) for k in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum'))
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r'^void\s+', '', r[iN]).replace('<unnamed>::', '').split('(')[0].split('<')[0].strip()
        a = agg.setdefault(name, {'launches': 0, 'dram_bytes': 0.0, 'us': 0.0})
        a['launches'] += 1
        a['dram_bytes'] += float(r[iR]) * UNIT[units[iR]] + float(r[iW]) * UNIT[units[iW]]
        a['us'] += float(r[iT]) * TIME[units[iT]]
    res = {k: {'launches_per_step': v['launches'], 'dram_bytes_per_launch': v['dram_bytes'] / v['launches'], 'us_per_launch': v['us'] / v['launches']}
           for k, v in agg.items()}
    import bench
    print(json.dumps({'source': path.split('/')[-1], 'csrc_digest': bench.csrc_digest(), 'csrc_files': bench.csrc_file_digests(), 'note': 'ncu --set full --clock-control none, one eager step, caches flushed before every kernel '
                      '(cold): an upper bound on the traffic of the same kernel inside the step graph, where producers leave their outputs in L2',
                      'kernels': res}, indent=1))


if __name__ == '__main__':
    main(sys.argv[1])
