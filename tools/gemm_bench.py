#!/usr/bin/env python
"""Micro-benchmark of the dense entry points at the LSTM shapes of the benchmark workloads: FFMA tiles, the round-1 tcgen05 kernel
(thread-staged operands), the TMA-fed warp-specialised kernel (tcgemm_ws.cu), the grouped backward of a layer, and — for context
only, never used by the product — the library sgemm the reference's nn.LSTM path runs (torch.matmul, fp32, TF32 off).
Every timing replays a CUDA graph of `reps` back-to-back launches (device time, no host launch overhead).

    python tools/gemm_bench.py > profiles/r02_gemm_bench.txt"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from subgnn_b200 import _abi  # noqa: E402
from subgnn_b200._abi import call, ptr  # noqa: E402


def timeit(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(s.cuda_stream)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn(s.cuda_stream)
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            a.record(s)
            g.replay()
            b.record(s)
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / reps * 1e3)
    return best


def main():
    dev = 'cuda'
    gen = torch.Generator(device=dev).manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    hbm = 6557.8
    for (M, N, K, T) in [(10000, 512, 64, 10), (10000, 512, 128, 10), (7200, 1024, 128, 10), (16100, 1024, 128, 23)]:
        x = torch.randn(M, K, device=dev, generator=gen)
        w = torch.randn(N, K, device=dev, generator=gen)
        bias = torch.randn(N, device=dev, generator=gen)
        y = torch.empty(M, N, device=dev)
        dy = torch.randn(M, N, device=dev, generator=gen)
        dx = torch.zeros(M, K, device=dev)
        dw = torch.zeros(N, K, device=dev)
        H = N // 8
        out_h = torch.randn(M + 1, 2 * H, device=dev, generator=gen)
        dwhh = torch.zeros(2, 4 * H, H, device=dev)
        fl = 2.0 * M * N * K
        by_f = 4.0 * (M * K + N * K + M * N)
        rows = []
        for name, env in (('ffma', None), ('tc-r1', '1'), ('tc-ws', '0')):
            if env is not None:
                os.environ['SUBGNN_TC_LEGACY_FORCE'] = env
            p = 'subgnn_' if name == 'ffma' else 'subgnn_tc_'
            extra = (None,) if name == 'ffma' else ()
            t1 = timeit(lambda st: call(p + 'linear_fwd', ptr(x), K, None, ptr(w), K, ptr(bias), ptr(y), N, M, N, K, 0, st))
            t2 = timeit(lambda st: call(p + 'linear_bwd_input', ptr(dy), N, ptr(w), K, ptr(dx), K, None, M, N, K, 0, st))
            t3 = timeit(lambda st: call(p + 'linear_bwd_weight', ptr(dy), N, ptr(x), K, None, ptr(dw), K, None, M, N, K, *extra, st))
            rows.append((name, t1, t2, t3))
        c1 = timeit(lambda st: torch.addmm(bias, x, w.t(), out=y))
        c2 = timeit(lambda st: torch.matmul(dy, w, out=dx))
        c3 = timeit(lambda st: torch.matmul(dy.t(), x, out=dw))
        rows.append(('cublas-fp32', c1, c2, c3))
        for name, t1, t2, t3 in rows:
            print('M=%5d N=%4d K=%3d %-11s fwd %6.1f us (%5.1f TF, %4.2f of HBM)  bwd_input %6.1f us (%5.1f TF)  bwd_weight %6.1f us (%5.1f TF, %4.2f of HBM)' %
                  (M, N, K, name, t1, fl / t1 / 1e6, by_f / t1 / 1e3 / hbm, t2, fl / t2 / 1e6, t3, fl / t3 / 1e6, by_f / t3 / 1e3 / hbm))
        # grouped backward of one layer: input gradient + input-weight gradient + both recurrent-weight gradients, ONE launch
        gd = _abi.gemm_desc
        descs = [gd(_abi.GEMM_BWD_INPUT, dy.data_ptr(), N, w.data_ptr(), K, dx.data_ptr(), K, M, N, K),
                 gd(_abi.GEMM_BWD_WEIGHT, dy.data_ptr(), N, x.data_ptr(), K, dw.data_ptr(), K, M, N, K)]
        for d, sh in ((0, -1), (1, 1)):
            descs.append(gd(_abi.GEMM_BWD_WEIGHT_SHIFT, dy.data_ptr() + 4 * d * 4 * H, N, out_h.data_ptr() + 4 * d * H, 2 * H, dwhh.data_ptr() + 4 * d * 4 * H * H, H,
                            M, 4 * H, H, shift=sh, period=T))
        tg = timeit(lambda st: _abi.gemm_group(descs, st))
        t_each = [timeit(lambda st, d=d: _abi.gemm_group([d], st)) for d in descs]
        by_g = 4.0 * (M * N + M * K * 2 + M * 2 * H + 2 * N * K + 2 * 4 * H * H)
        print('M=%5d N=%4d K=%3d grouped layer backward (4 products, dG read from HBM once): %6.1f us = %4.2f of HBM; the same four as separate launches: %s us' %
              (M, N, K, tg, by_g / tg / 1e3 / hbm, ' + '.join('%.1f' % t for t in t_each)))


if __name__ == '__main__':
    main()
