#!/usr/bin/env python
"""Sweeps the tile width / operand-buffer count / reduction splits of the tcgen05 GEMM entry points at the shapes that sit on
the critical chain of the training step (tuning aid: the launch heuristics in csrc/tcgemm.cu are set from this table)."""
import itertools
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from subgnn_b200._abi import call, ptr, stream_ptr  # noqa: E402


def timeit(fn, reps=40):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def sweep(label, fn, keys):
    rows = []
    for combo in itertools.product(*[v for _, v in keys]):
        for (k, _), v in zip(keys, combo):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        try:
            rows.append((timeit(fn), combo))
        except Exception as e:                                   # noqa: BLE001
            rows.append((float('inf'), combo + (str(e)[:40],)))
    for k, _ in keys:
        os.environ.pop(k, None)
    rows.sort(key=lambda r: r[0])
    print('%s  (%s)' % (label, ', '.join(k.replace('SUBGNN_TC_', '') for k, _ in keys)))
    for t, c in rows[:6]:
        print('    %7.1f us  %s' % (t, c))
    dflt = [r for r in rows if all(v is None for v in r[1])]
    if dflt:
        print('    default: %.1f us' % dflt[0][0])


def main():
    dev = 'cuda'
    shapes = {'ppi_bp': (10000, 64), 'hpo_metab': (7200, 128), 'em_user': (16100, 128)}
    which = sys.argv[1:] or ['ppi_bp']
    g = torch.Generator(device=dev).manual_seed(0)
    st = stream_ptr()
    for name in which:
        M, H = shapes[name]
        table = torch.randn(20000, H, device=dev, generator=g)
        ids = torch.randint(1, 20000, (M,), device=dev, dtype=torch.int32, generator=g)
        hprev = torch.randint(0, M, (M,), device=dev, dtype=torch.int32, generator=g)
        G = torch.randn(M, 8 * H, device=dev, generator=g)
        X1 = torch.randn(M + 1, 2 * H, device=dev, generator=g)
        w0 = torch.randn(8 * H, H, device=dev, generator=g)
        w1 = torch.randn(8 * H, 2 * H, device=dev, generator=g)
        bias = torch.randn(8 * H, device=dev, generator=g)
        dX1 = torch.zeros(M, 2 * H, device=dev)
        dE = torch.zeros(20000, H, device=dev)
        dw0, dw1, dwh = torch.zeros(8 * H, H, device=dev), torch.zeros(8 * H, 2 * H, device=dev), torch.zeros(4 * H, H, device=dev)
        S, NT = ('SUBGNN_TC_STAGES', [None, 1, 2]), [None, 32, 64, 128]
        print('==== %s: M = %d rows, H = %d' % (name, M, H))
        sweep('fwd proj layer 0  (M x 8H x H, gather)', lambda: call('subgnn_tc_linear_fwd', ptr(table), H, ptr(ids), ptr(w0), H, ptr(bias), ptr(G), 8 * H, M, 8 * H, H, 0, st),
              [S, ('SUBGNN_TC_NT_FWD', NT)])
        sweep('fwd proj layer 1  (M x 4H x 2H)', lambda: call('subgnn_tc_linear_fwd', ptr(X1), 2 * H, None, ptr(w1), 2 * H, ptr(bias), ptr(G), 8 * H, M, 4 * H, 2 * H, 0, st),
              [S, ('SUBGNN_TC_NT_FWD', NT)])
        sweep('bwd_input layer 1 (M x 2H, reduce 4H)', lambda: call('subgnn_tc_linear_bwd_input', ptr(G), 8 * H, ptr(w1), 2 * H, ptr(dX1), 2 * H, None, M, 4 * H, 2 * H, 0, st),
              [S, ('SUBGNN_TC_NT_BWI', NT)])
        sweep('bwd_input layer 0 (scatter M x H, reduce 8H)', lambda: call('subgnn_tc_linear_bwd_input', ptr(G), 8 * H, ptr(w0), H, ptr(dE), H, ptr(ids), M, 8 * H, H, 1, st),
              [S, ('SUBGNN_TC_NT_BWI', NT), ('SUBGNN_TC_SPLITS_BWI', [None, 1, 2, 4])])
        sweep('bwd_weight W_ih layer 0 (8H x H, reduce M, gather)', lambda: call('subgnn_tc_linear_bwd_weight', ptr(G), 8 * H, ptr(table), H, ptr(ids), ptr(dw0), H, None, M, 8 * H, H, st),
              [S, ('SUBGNN_TC_NT_BWW', NT), ('SUBGNN_TC_SPLITS_BWW', [None, 20, 40, 79, 158])])
        sweep('bwd_weight W_ih layer 1 (4H x 2H, reduce M)', lambda: call('subgnn_tc_linear_bwd_weight', ptr(G), 8 * H, ptr(X1), 2 * H, None, ptr(dw1), 2 * H, None, M, 4 * H, 2 * H, st),
              [S, ('SUBGNN_TC_NT_BWW', NT), ('SUBGNN_TC_SPLITS_BWW', [None, 20, 40, 79, 158])])
        sweep('bwd_weight W_hh (4H x H, reduce M, row gather)', lambda: call('subgnn_tc_linear_bwd_weight', ptr(G), 8 * H, ptr(X1), 2 * H, ptr(hprev), ptr(dwh), H, None, M, 4 * H, H, st),
              [S, ('SUBGNN_TC_NT_BWW', NT), ('SUBGNN_TC_SPLITS_BWW', [None, 20, 40, 79, 158])])


if __name__ == '__main__':
    main()
