#!/usr/bin/env python
"""Runs a few eager training steps of one bench workload with the cudaProfiler range around the LAST ones, for

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv \
        python tools/profile_step.py --workload ppi_bp --steps 2
    ncu --profile-from-start off --set full --clock-control none --import-source on -o R \
        python tools/profile_step.py --workload ppi_bp --steps 1

(the launch list of a step, and the full-section capture of every kernel of one step)."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='ppi_bp')
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--warm', type=int, default=3)
    a = ap.parse_args()
    from subgnn_b200.engine import Engine
    torch.cuda.set_device(0)
    hp, g, prepared, _ = bench.build_workload(a.workload, 'cuda:0')
    eng = Engine(hp, prepared, device='cuda:0', graph=g, seed=1234)
    eng.init_parameters(seed=7)
    n_train = len(prepared['labels']['train'])
    batches = bench.batches_for(n_train, hp['batch_size'], a.warm + a.steps, 0, 1)
    for i in range(a.warm):
        eng.train_step(batches[i], use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(a.steps):
        eng.train_step(batches[a.warm + i], use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == '__main__':
    main()
