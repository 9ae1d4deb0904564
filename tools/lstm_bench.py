#!/usr/bin/env python
"""Times the walk-encoder LSTM recurrence kernels alone (CUDA events, L2-resident inputs as inside the step) for the
benchmark shapes; SUBGNN_LSTM_TILE=<n> overrides the sequences-per-CTA choice (tuning aid)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from subgnn_b200._abi import call, ptr  # noqa: E402

SHAPES = {'ppi_bp': (1000, 10, 64), 'hpo_metab': (720, 10, 128), 'em_user': (700, 23, 128), 'density': (420, 10, 32)}


def run(name, n_seq, T, H, reps=30):
    dev = 'cuda'
    torch.manual_seed(0)
    M = n_seq * T
    G0 = torch.randn(M, 8 * H, device=dev) * 0.5
    whh = torch.randn(2, 4 * H, H, device=dev) * 0.1
    b = torch.zeros(2, 4 * H, device=dev)
    whh_t, bsum = torch.zeros(2 * H * 4 * H, device=dev), torch.zeros(8 * H, device=dev)
    OUT, CS, dOUT = torch.zeros(M + 1, 2 * H, device=dev), torch.zeros(M, 2 * H, device=dev), torch.randn(M, 2 * H, device=dev)
    db = torch.zeros(2, 2, 4 * H, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    call('subgnn_lstm_prep', ptr(whh), ptr(b), ptr(b), ptr(whh_t), ptr(bsum), H, st)
    tf, tb = [], []
    for i in range(reps + 3):
        G = G0.clone()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        call('subgnn_lstm_recur_fwd', ptr(G), ptr(whh_t), ptr(OUT), ptr(CS), n_seq, T, H, T, T, st)
        e[1].record()
        call('subgnn_lstm_recur_bwd', ptr(G), ptr(whh), ptr(OUT), ptr(CS), ptr(dOUT), n_seq, T, H, T, T, 1, ptr(db[0]), ptr(db[1]), st)
        e[2].record()
        torch.cuda.synchronize()
        if i >= 3:
            tf.append(e[0].elapsed_time(e[1]))
            tb.append(e[1].elapsed_time(e[2]))
    ttc = []
    if H == 64:                                # tensor-core forward recurrence (lstm_tc.cu), same inputs
        for i in range(reps + 3):
            G = G0.clone()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            e[0].record()
            call('subgnn_lstm_recur_fwd_tc', ptr(G), ptr(whh), ptr(OUT), ptr(CS), n_seq, T, H, T, T, None, 0.0, 0, 0, None, st)
            e[1].record()
            torch.cuda.synchronize()
            if i >= 3:
                ttc.append(e[0].elapsed_time(e[1]))
        ttc.sort()
        print('%-10s tensor-core forward recurrence (tcgen05, 4-CTA clusters): %.1f us (%.2f us/step)' %
              (name, 1e3 * ttc[len(ttc) // 2], 1e3 * ttc[len(ttc) // 2] / T), flush=True)
    tf.sort(), tb.sort()
    print('%-10s n_seq=%d T=%d H=%d tile=%s  fwd %.1f us (%.2f us/step)  bwd %.1f us (%.2f us/step)' %
          (name, n_seq, T, H, os.environ.get('SUBGNN_LSTM_TILE', 'auto'), 1e3 * tf[len(tf) // 2], 1e3 * tf[len(tf) // 2] / T,
           1e3 * tb[len(tb) // 2], 1e3 * tb[len(tb) // 2] / T), flush=True)


if __name__ == '__main__':
    for name in (sys.argv[1:] or list(SHAPES)):
        run(name, *SHAPES[name])
