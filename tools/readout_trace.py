#!/usr/bin/env python
"""Phase clocks of CTA 0 of the cluster readout kernel inside a real step (debug build -DWS_TRACE, SUBGNN_B200_LIB=build/libsubgnn_trace.so)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from subgnn_b200 import _abi  # noqa: E402
from subgnn_b200.engine import Engine  # noqa: E402


def main():
    torch.cuda.set_device(0)
    hp, g, prepared, _ = bench.build_workload('ppi_bp', 'cuda:0')
    eng = Engine(hp, prepared, device='cuda:0', graph=g, seed=1234)
    eng.init_parameters(seed=7)
    batches = bench.batches_for(len(prepared['labels']['train']), hp['batch_size'], 12, 0, 1)
    names = ['start', 'weights staged', 'after pdl wait', 'Z staged', 'phase 1 done', 'cluster barrier 1', 'fwd + loss done', 'bwd small done', 'cluster barrier 2', 'end']
    for use_graph in (False, True):
        for i in range(6):
            eng.train_step(batches[i], use_graph=use_graph)
        buf = (C.c_longlong * 64)()
        _abi.lib.subgnn_ro_trace_read(buf)
        print('graph' if use_graph else 'eager')
        for i, n in enumerate(names):
            print('  %-18s %8d' % (n, buf[i] - buf[0]))


if __name__ == '__main__':
    main()
