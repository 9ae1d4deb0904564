"""The tensor-core forward recurrence (lstm_tc.cu: tcgen05 3xTF32, 4-CTA clusters, DSMEM hand-over of h) against the register-tiled
FFMA2 kernel (lstm_reg.cu) on the same inputs — both implement nn.LSTM's recurrence (SubGNN.py:60-88) on precomputed input
projections: gate activations (written back into G), cell states, outputs and the fused inter-layer dropout copy.
Tolerance: 3xTF32 carries ~2^-22 relative error per product; rtol 1e-5 / atol 2e-6 over T = 10 .. 23 steps (stated)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(fn, G0, whh, n_seq, T, H, sf, sr, p):
    from subgnn_b200._abi import call, ptr
    dev = 'cuda'
    M = n_seq * T
    G = G0.clone()
    OUT, CS, X = torch.zeros(M + 1, 2 * H, device=dev), torch.zeros(M, 2 * H, device=dev), torch.zeros(M, 2 * H, device=dev)
    step = torch.full((1,), 3, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    if fn == 'tc':
        call('subgnn_lstm_recur_fwd_tc', ptr(G), ptr(whh), ptr(OUT), ptr(CS), n_seq, T, H, sf, sr, ptr(X), p, 1234, 9, ptr(step), st)
    else:
        b = torch.zeros(2, 4 * H, device=dev)
        whh_t, bsum = torch.zeros(2 * H * 4 * H, device=dev), torch.zeros(8 * H, device=dev)
        call('subgnn_lstm_prep', ptr(whh), ptr(b), ptr(b), ptr(whh_t), ptr(bsum), H, st)
        call('subgnn_lstm_recur_fwd_drop', ptr(G), ptr(whh_t), ptr(OUT), ptr(CS), n_seq, T, H, sf, sr, ptr(X), p, 1234, 9, ptr(step), st)
    torch.cuda.synchronize()
    return G, OUT[:M], CS, X


@pytest.mark.parametrize('n_seq,T,sf,sr,p', [(1000, 10, 10, 10, 0.0), (1000, 10, 10, 1, 0.2), (300, 23, 23, 23, 0.2), (5, 4, 4, 4, 0.0), (129, 7, 7, 0, 0.0)])
def test_tensor_core_recurrence_matches_register_kernel(n_seq, T, sf, sr, p):
    H = 64
    gen = torch.Generator(device='cuda').manual_seed(n_seq + T)
    G0 = torch.randn(n_seq * T, 8 * H, device='cuda', generator=gen) * 0.7
    whh = torch.randn(2, 4 * H, H, device='cuda', generator=gen) * 0.15
    a = _run('reg', G0, whh, n_seq, T, H, sf, sr, p)
    b = _run('tc', G0, whh, n_seq, T, H, sf, sr, p)
    # compare only what the kernels define: direction d is written for its taken steps (forward: t < sf; reverse: t >= T - sr)
    t_idx = torch.arange(n_seq * T, device='cuda') % T
    for name, x, y, width in (('gates', a[0], b[0], 4 * H), ('h', a[1], b[1], H), ('c', a[2], b[2], H), ('dropout(h)', a[3], b[3], H)):
        for d, taken in ((0, t_idx < sf), (1, t_idx >= T - sr)):
            xs, ys = x[taken][:, d * width:(d + 1) * width], y[taken][:, d * width:(d + 1) * width]
            np.testing.assert_allclose(ys.cpu().numpy(), xs.cpu().numpy(), rtol=1e-5, atol=2e-6, err_msg='%s dir %d' % (name, d))
