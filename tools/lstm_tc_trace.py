#!/usr/bin/env python
"""Per-phase clocks of the tensor-core recurrence (lstm_tc.cu built with -DTC_TRACE into build/libsubgnn_trace.so):

    SUBGNN_B200_LIB=build/libsubgnn_trace.so python tools/lstm_tc_trace.py

role 0 = MMA thread (0 step start, 1 fences done, 2 MMAs issued + commit), role 1 = epilogue thread 0 (0 enter, 1 accumulator ready,
2 TMEM read, 3 activations done, 4 DSMEM stores issued, 5 proxy fence, 6 cluster arrive, 7 global stores issued, 8 next G loads
issued, 9 cluster wait passed).  Prints cycles relative to the MMA thread's step start."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from subgnn_b200 import _abi  # noqa: E402
from subgnn_b200._abi import call, ptr  # noqa: E402


def main():
    n_seq, T, H = 1000, 10, 64
    dev = 'cuda'
    G = torch.randn(n_seq * T, 8 * H, device=dev) * 0.5
    whh = torch.randn(2, 4 * H, H, device=dev) * 0.1
    OUT, CS = torch.zeros(n_seq * T + 1, 2 * H, device=dev), torch.zeros(n_seq * T, 2 * H, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        call('subgnn_lstm_recur_fwd_tc', ptr(G), ptr(whh), ptr(OUT), ptr(CS), n_seq, T, H, T, T, None, 0.0, 0, 0, None, st)
    torch.cuda.synchronize()
    buf = (C.c_longlong * (2 * 16 * 64))()
    _abi.lib.subgnn_tc_trace_read.restype = C.c_int
    _abi.lib.subgnn_tc_trace_read(buf)
    get = lambda role, s, tag: buf[(role * 64 + s) * 16 + tag]
    for s in range(T):
        t0 = get(0, s, 0)
        print('step %2d  mma: fence %5d issue+commit %5d | epi: enter %6d ready %5d tmem %5d act %5d dsmem %5d pfence %5d arrive %5d gstore %5d gload %5d wait %5d | next step start %5d' %
              ((s, get(0, s, 1) - t0, get(0, s, 2) - t0) + tuple(get(1, s, k) - t0 for k in range(10)) + ((get(0, s + 1, 0) - t0) if s + 1 < T else 0,)))


if __name__ == '__main__':
    main()
