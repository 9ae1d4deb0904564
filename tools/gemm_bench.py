#!/usr/bin/env python
"""Micro-benchmark of the dense entry points (FFMA tiles vs tcgen05 3xTF32) at the LSTM shapes of the PPI-BP workload."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from subgnn_b200._abi import call, ptr, stream_ptr  # noqa: E402


def timeit(fn, reps=30):
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


def main():
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    table = torch.randn(17081, 128, device=dev, generator=g)
    for (M, N, K, gather) in [(10000, 512, 64, True), (10000, 256, 128, False), (10000, 512, 128, False), (16100, 1024, 128, True)]:
        x = table[:, :K].contiguous() if gather else torch.randn(M, K, device=dev, generator=g)
        ids = torch.randint(1, 17081, (M,), device=dev, dtype=torch.int32, generator=g) if gather else None
        w = torch.randn(N, K, device=dev, generator=g)
        bias = torch.randn(N, device=dev, generator=g)
        y = torch.empty(M, N, device=dev)
        dy = torch.randn(M, N, device=dev, generator=g)
        dx = torch.zeros_like(x)
        dw = torch.zeros(N, K, device=dev)
        db = torch.zeros(N, device=dev)
        st = stream_ptr()
        fl = 2.0 * M * N * K
        for tc in (False, True):
            p = 'subgnn_tc_' if tc else 'subgnn_'
            t1 = timeit(lambda: call(p + 'linear_fwd', ptr(x), K, ptr(ids), ptr(w), K, ptr(bias), ptr(y), N, M, N, K, 0, st))
            t2 = timeit(lambda: call(p + 'linear_bwd_input', ptr(dy), N, ptr(w), K, ptr(dx), K, ptr(ids), M, N, K, 1, st))
            if tc:
                t3 = timeit(lambda: call(p + 'linear_bwd_weight', ptr(dy), N, ptr(x), K, ptr(ids), ptr(dw), K, ptr(db), M, N, K, st))
            else:
                t3 = timeit(lambda: call(p + 'linear_bwd_weight', ptr(dy), N, ptr(x), K, ptr(ids), ptr(dw), K, ptr(db), M, N, K, None, st))
            print('M=%d N=%d K=%d gather=%d %-5s fwd %7.1f us (%5.1f TF)  bwd_input %7.1f us (%5.1f TF)  bwd_weight %7.1f us (%5.1f TF)' %
                  (M, N, K, gather, 'tc' if tc else 'ffma', t1, fl / t1 / 1e6, t2, fl / t2 / 1e6, t3, fl / t3 / 1e6))


if __name__ == '__main__':
    main()
