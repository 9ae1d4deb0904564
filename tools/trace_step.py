#!/usr/bin/env python
"""In-graph timeline of one training step: replays the captured step graph under torch.profiler (Kineto / CUPTI activity
records, which carry start time, duration and stream of every kernel node of a graph replay) and prints, for the median
replay, every kernel with its start offset, duration and stream, plus the stream-overlap summary.  Unlike the ncu launch
list (serialised, cold) this shows what is on the critical path of the step as it runs in the bench (L2 flushed before
every replay, as bench.py does).

    python tools/trace_step.py --workload ppi_bp [--steps 7] [--no-flush] > profiles/rNN_timeline_<workload>.txt
"""
import argparse
import json
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='ppi_bp')
    ap.add_argument('--steps', type=int, default=7)
    ap.add_argument('--no-flush', action='store_true')
    a = ap.parse_args()
    from subgnn_b200.engine import Engine
    torch.cuda.set_device(0)
    hp, g, prepared, _ = bench.build_workload(a.workload, 'cuda:0')
    eng = Engine(hp, prepared, device='cuda:0', graph=g, seed=1234)
    eng.init_parameters(seed=7)
    n_train = len(prepared['labels']['train'])
    batches = bench.batches_for(n_train, hp['batch_size'], 8 + a.steps, 0, 1)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda:0')
    for i in range(8):
        eng.train_step(batches[i], use_graph=True)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(a.steps):
            if not a.no_flush:
                flush.zero_()
            torch.cuda.synchronize()
            eng.train_step(batches[8 + i], use_graph=True)
            torch.cuda.synchronize()
    with tempfile.NamedTemporaryFile(suffix='.json') as f:
        prof.export_chrome_trace(f.name)
        tr = json.load(open(f.name))
    ks = [e for e in tr['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and e.get('ph') == 'X']
    ks.sort(key=lambda e: e['ts'])
    # split into replays: a gap of > 50 us of GPU idleness that follows the flush kernel separates steps
    steps, cur, last_end = [], [], None
    for e in ks:
        nm = e['name']
        if last_end is not None and e['ts'] - last_end > 100 and cur:      # host-side synchronize between replays
            steps.append(cur)
            cur = []
        last_end = e['ts'] + e['dur']
        if 'FillFunctor' in nm or 'vectorized_elementwise' in nm and e['dur'] > 20:
            if cur:
                steps.append(cur)
            cur = []
            continue
        cur.append(e)
    if cur:
        steps.append(cur)
    steps = [s for s in steps if len(s) > 10]
    spans = [max(e['ts'] + e['dur'] for e in s) - min(e['ts'] for e in s) for s in steps]
    order = sorted(range(len(steps)), key=lambda i: spans[i])
    med = order[len(order) // 2]
    s = steps[med]
    t0 = min(e['ts'] for e in s)
    print('workload %s: %d replays traced, span of each (us): %s; median replay shown (%d kernels, %.1f us)' %
          (a.workload, len(steps), ' '.join('%.1f' % x for x in spans), len(s), spans[med]))
    streams = sorted({e['args'].get('stream', 0) for e in s})
    print('%8s %8s %8s  %-3s %s' % ('start', 'dur', 'end', 'str', 'kernel'))
    for e in s:
        nm = e['name'].replace('void ', '').replace('(anonymous namespace)::', '')
        nm = nm.split('(')[0][:70]
        print('%8.1f %8.1f %8.1f  %-3d %s  grid=%s' % (e['ts'] - t0, e['dur'], e['ts'] + e['dur'] - t0, streams.index(e['args'].get('stream', 0)), nm,
                                                      e['args'].get('grid', '')))
    # busy time per stream and union
    iv = sorted((e['ts'] - t0, e['ts'] + e['dur'] - t0) for e in s)
    union, end = 0.0, 0.0
    for a_, b_ in iv:
        if b_ > end:
            union += b_ - max(a_, end)
            end = b_
    print('sum of kernel durations %.1f us, union (GPU busy) %.1f us, idle inside the step %.1f us' % (sum(e['dur'] for e in s), union, spans[med] - union))


if __name__ == '__main__':
    main()
