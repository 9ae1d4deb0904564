#!/usr/bin/env python
"""Launches each dense product of one LSTM layer once on the TMA-fed kernel (for ncu):
    ncu --set full --import-source on --clock-control none -k regex:tc_gemm_ws -o gpurun_out/gemm_ws python tools/gemm_one.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from subgnn_b200 import _abi  # noqa: E402


def main():
    M, N, K, T = 10000, 512, 64, 10
    if len(sys.argv) > 3:
        M, N, K = (int(v) for v in sys.argv[1:4])
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    x, w, bias = torch.randn(M, K, device=dev, generator=g), torch.randn(N, K, device=dev, generator=g), torch.randn(N, device=dev, generator=g)
    y, dy = torch.empty(M, N, device=dev), torch.randn(M, N, device=dev, generator=g)
    dx, dw = torch.zeros(M, K, device=dev), torch.zeros(N, K, device=dev)
    H = N // 8
    out_h, dwhh = torch.randn(M + 1, 2 * H, device=dev, generator=g), torch.zeros(2, 4 * H, H, device=dev)
    st = _abi.stream_ptr()
    gd = _abi.gemm_desc
    fwd = gd(_abi.GEMM_FWD, x.data_ptr(), K, w.data_ptr(), K, y.data_ptr(), N, M, N, K, bias=bias.data_ptr())
    bwi = gd(_abi.GEMM_BWD_INPUT, dy.data_ptr(), N, w.data_ptr(), K, dx.data_ptr(), K, M, N, K)
    bww = gd(_abi.GEMM_BWD_WEIGHT, dy.data_ptr(), N, x.data_ptr(), K, dw.data_ptr(), K, M, N, K)
    hh = [gd(_abi.GEMM_BWD_WEIGHT_SHIFT, dy.data_ptr() + 4 * d * 4 * H, N, out_h.data_ptr() + 4 * d * H, 2 * H, dwhh.data_ptr() + 4 * d * 4 * H * H, H,
             M, 4 * H, H, shift=sh, period=T) for d, sh in ((0, -1), (1, 1))]
    for rep in range(2):
        for descs in ([fwd], [bwi], [bww], [bwi, bww] + hh):
            _abi.gemm_group(descs, st)
            torch.cuda.synchronize()


if __name__ == '__main__':
    main()
