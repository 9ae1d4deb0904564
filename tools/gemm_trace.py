#!/usr/bin/env python
"""Per-role event clocks of CTA 0 of the TMA-fed GEMM (debug build of the library with -DWS_TRACE, selected with SUBGNN_B200_LIB):
    SUBGNN_B200_LIB=build/libsubgnn_trace.so python tools/gemm_trace.py [M N K]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from subgnn_b200 import _abi  # noqa: E402


def read():
    buf = (C.c_longlong * 4096)()
    n = (C.c_int * 4)()
    _abi.lib.subgnn_ws_trace_read(buf, n)
    return buf, n


def dump(tag):
    buf, n = read()
    names = ['producer', 'mma', 'converter', 'epilogue']
    ev = []
    for r in range(4):
        for i in range(min(n[r], 512)):
            ev.append((buf[r * 1024 + 2 * i + 1], names[r], buf[r * 1024 + 2 * i]))
    ev.sort()
    t0 = ev[0][0]
    print('==== %s: %d events' % (tag, len(ev)))
    for t, role, k in ev:
        print('%8d  %-9s %d' % (t - t0, role, k))


def main():
    M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (10000, 512, 64)
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    x, w, bias = torch.randn(M, K, device=dev, generator=g), torch.randn(N, K, device=dev, generator=g), torch.randn(N, device=dev, generator=g)
    y, dy = torch.empty(M, N, device=dev), torch.randn(M, N, device=dev, generator=g)
    dw = torch.zeros(N, K, device=dev)
    st = _abi.stream_ptr()
    gd = _abi.gemm_desc
    fwd = gd(_abi.GEMM_FWD, x.data_ptr(), K, w.data_ptr(), K, y.data_ptr(), N, M, N, K, bias=bias.data_ptr())
    bww = gd(_abi.GEMM_BWD_WEIGHT, dy.data_ptr(), N, x.data_ptr(), K, dw.data_ptr(), K, M, N, K)
    for name, d in (('fwd', fwd), ('bwd_weight', bww)):
        for rep in range(2):
            _abi.gemm_group([d], st)
            torch.cuda.synchronize()
            if rep == 0:
                read()                                   # discard the cold run
        dump(name)


if __name__ == '__main__':
    main()
