"""tcgen05 3xTF32 GEMMs vs a float64 reference and vs the FFMA kernels (fp32-class accuracy required)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _call(name, *args):
    from subgnn_b200._abi import call, ptr, stream_ptr
    call(name, *[ptr(a) if isinstance(a, torch.Tensor) else a for a in args], stream_ptr())


@pytest.mark.parametrize('M,N,K,gather,relu', [(300, 512, 64, False, False), (1000, 512, 128, True, False), (130, 200, 40, True, True),
                                                (128, 128, 32, False, False), (77, 24, 8, False, True)])
def test_tc_linear_fwd(M, N, K, gather, relu):
    g = torch.Generator().manual_seed(M + N + K)
    n_table = 500
    x = torch.randn(n_table if gather else M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.3
    b = torch.randn(N, generator=g)
    ids = torch.randint(0, n_table, (M,), generator=g, dtype=torch.int32) if gather else None
    rows = x[ids.long()] if gather else x
    want = rows.double().numpy() @ w.double().numpy().T + b.double().numpy()
    if relu:
        want = np.maximum(want, 0)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    idd = ids.cuda() if gather else None
    y_tc = torch.zeros(M, N, device='cuda')
    y_ff = torch.zeros(M, N, device='cuda')
    _call('subgnn_tc_linear_fwd', xd, K, idd, wd, K, bd, y_tc, N, M, N, K, int(relu))
    _call('subgnn_linear_fwd', xd, K, idd, wd, K, bd, y_ff, N, M, N, K, int(relu))
    torch.cuda.synchronize()
    scale = np.abs(want).max()
    err_tc = np.abs(y_tc.cpu().numpy() - want).max() / scale
    err_ff = np.abs(y_ff.cpu().numpy() - want).max() / scale
    assert err_ff < 2e-6
    assert err_tc < 4e-6, (err_tc, err_ff)


@pytest.mark.parametrize('M,N,K,mode', [(300, 512, 64, 'store'), (1000, 256, 128, 'acc'), (333, 200, 40, 'scatter'), (64, 24, 8, 'store')])
def test_tc_linear_bwd_input(M, N, K, mode):
    g = torch.Generator().manual_seed(M * 3 + N + K)
    dy = torch.randn(M, N, generator=g)
    w = torch.randn(N, K, generator=g) * 0.3
    want = dy.double().numpy() @ w.double().numpy()
    dyd, wd = dy.cuda(), w.cuda()
    if mode == 'scatter':
        n_table = 50
        ids = torch.randint(0, n_table, (M,), generator=g, dtype=torch.int32)
        ref = np.zeros((n_table, K))
        np.add.at(ref, ids.numpy(), want)
        ref[0] = 0                                         # PAD row is skipped
        out = torch.zeros(n_table, K, device='cuda')
        _call('subgnn_tc_linear_bwd_input', dyd, N, wd, K, out, K, ids.cuda(), M, N, K, 1)
    else:
        base = torch.randn(M, K, generator=g)
        out = base.cuda().clone()
        ref = want + (base.double().numpy() if mode == 'acc' else 0)
        _call('subgnn_tc_linear_bwd_input', dyd, N, wd, K, out, K, None, M, N, K, 1 if mode == 'acc' else 0)
    torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err < 5e-6, err


@pytest.mark.parametrize('M,N,K,gather', [(10000, 512, 64, True), (5000, 256, 128, False), (700, 96, 40, True), (33, 32, 8, False)])
def test_tc_linear_bwd_weight(M, N, K, gather):
    g = torch.Generator().manual_seed(M + 7 * N + K)
    dy = torch.randn(M, N, generator=g)
    n_table = 400
    x = torch.randn(n_table if gather else M, K, generator=g)
    ids = torch.randint(0, n_table, (M,), generator=g, dtype=torch.int32) if gather else None
    rows = x[ids.long()] if gather else x
    want_w = dy.double().numpy().T @ rows.double().numpy()
    want_b = dy.double().numpy().sum(0)
    dw = torch.zeros(N, K, device='cuda')
    db = torch.zeros(N, device='cuda')
    _call('subgnn_tc_linear_bwd_weight', dy.cuda(), N, x.cuda(), K, ids.cuda() if gather else None, dw, K, db, M, N, K)
    torch.cuda.synchronize()
    assert np.abs(dw.cpu().numpy() - want_w).max() / np.abs(want_w).max() < 5e-6
    assert np.abs(db.cpu().numpy() - want_b).max() / np.abs(want_b).max() < 5e-6


# ---- the TMA-fed warp-specialised grouped kernel (tcgemm_ws.cu) ----------------------------------------------------------------
def _ptr_off(t, n_floats=0):
    return t.data_ptr() + 4 * n_floats


def test_dense_calls_run_on_the_tma_kernel():
    """the dense (no gather) entry points must land on tc_gemm_ws_kernel — the kernel the benchmarked step runs"""
    from subgnn_b200 import _abi
    _abi.variant_log(reset=True)
    x, w, y = torch.randn(300, 64, device='cuda'), torch.randn(512, 64, device='cuda'), torch.zeros(300, 512, device='cuda')
    _call('subgnn_tc_linear_fwd', x, 64, None, w, 64, None, y, 512, 300, 512, 64, 0)
    torch.cuda.synchronize()
    assert 'tc_gemm_ws_kernel<1>' in _abi.variant_log()
    want = x.cpu().double().numpy() @ w.cpu().double().numpy().T
    assert np.abs(y.cpu().numpy() - want).max() / np.abs(want).max() < 4e-6


@pytest.mark.parametrize('n_seq,T,H', [(37, 10, 64), (1000, 10, 64), (50, 23, 128), (9, 5, 32), (3, 1, 32)])
def test_recurrent_weight_gradient_shift(n_seq, T, H):
    """dW_hh[d] = sum over (sequence, t) of dG_t^T h_(t -/+ 1) with h = 0 beyond the sequence ends (nn.LSTM autograd, SubGNN.py:73):
    the B box is fetched one row up / down and the boundary rows of dG are zeroed in shared memory."""
    from subgnn_b200 import _abi
    g = torch.Generator().manual_seed(n_seq * 31 + T)
    M = n_seq * T
    dG = torch.randn(M, 8 * H, generator=g)
    OUT = torch.randn(M + 1, 2 * H, generator=g)
    OUT[M] = 0
    dGd, OUTd = dG.cuda(), OUT.cuda()
    dW = torch.zeros(2, 4 * H, H, device='cuda')
    descs = []
    for d, shift in ((0, -1), (1, 1)):
        descs.append(_abi.gemm_desc(_abi.GEMM_BWD_WEIGHT_SHIFT, _ptr_off(dGd, d * 4 * H), 8 * H, _ptr_off(OUTd, d * H), 2 * H, _ptr_off(dW, d * 4 * H * H), H,
                                    M, 4 * H, H, shift=shift, period=T))
    _abi.gemm_group(descs, _abi.stream_ptr())
    torch.cuda.synchronize()
    dG3, O3 = dG.double().view(n_seq, T, 8 * H), OUT[:M].double().view(n_seq, T, 2 * H)
    for d in range(2):
        gd, hd = dG3[:, :, d * 4 * H:(d + 1) * 4 * H], O3[:, :, d * H:(d + 1) * H]
        if T == 1:
            want = torch.zeros(4 * H, H, dtype=torch.float64)
        elif d == 0:
            want = torch.einsum('stn,stk->nk', gd[:, 1:], hd[:, :-1])
        else:
            want = torch.einsum('stn,stk->nk', gd[:, :-1], hd[:, 1:])
        got = dW[d].cpu().double()
        scale = max(float(want.abs().max()), 1.0)
        assert float((got - want).abs().max()) / scale < 5e-6, (d, float((got - want).abs().max()), scale)


def test_grouped_launch_of_a_layer_backward():
    """one launch = input gradient (row scatter into a table) + input-weight gradient + both recurrent-weight gradients of an LSTM
    layer, every product against float64; also the strided 'last rows' views (rows t = T-1 of every sequence)."""
    from subgnn_b200 import _abi
    g = torch.Generator().manual_seed(5)
    n_seq, T, H, D, n_table = 120, 10, 64, 64, 300
    M = n_seq * T
    dG = torch.randn(M, 8 * H, generator=g)
    X = torch.randn(M, D, generator=g)
    OUT = torch.randn(M + 1, 2 * H, generator=g)
    W = torch.randn(8 * H, D, generator=g) * 0.2
    ids = torch.randint(0, n_table, (M,), generator=g, dtype=torch.int32)
    dev = lambda t: t.cuda()
    dGd, Xd, OUTd, Wd, idd = dev(dG), dev(X), dev(OUT), dev(W), dev(ids)
    dE = torch.zeros(n_table, D, device='cuda')
    dWih = torch.zeros(8 * H, D, device='cuda')
    dWhh = torch.zeros(2, 4 * H, H, device='cuda')
    descs = [_abi.gemm_desc(_abi.GEMM_BWD_INPUT, dGd.data_ptr(), 8 * H, Wd.data_ptr(), D, dE.data_ptr(), D, M, 8 * H, D, scatter_ids=idd.data_ptr(), accumulate=1),
             _abi.gemm_desc(_abi.GEMM_BWD_WEIGHT, dGd.data_ptr(), 8 * H, Xd.data_ptr(), D, dWih.data_ptr(), D, M, 8 * H, D)]
    for d, shift in ((0, -1), (1, 1)):
        descs.append(_abi.gemm_desc(_abi.GEMM_BWD_WEIGHT_SHIFT, _ptr_off(dGd, d * 4 * H), 8 * H, _ptr_off(OUTd, d * H), 2 * H, _ptr_off(dWhh, d * 4 * H * H), H,
                                    M, 4 * H, H, shift=shift, period=T))
    _abi.variant_log(reset=True)
    _abi.gemm_group(descs, _abi.stream_ptr())
    torch.cuda.synchronize()
    assert 'tc_gemm_ws_kernel<4>' in _abi.variant_log()
    dx = dG.double() @ W.double()
    ref = torch.zeros(n_table, D, dtype=torch.float64).index_add_(0, ids.long(), dx)
    ref[0] = 0
    rel = lambda got, want: float((got.cpu().double() - want).abs().max()) / float(want.abs().max())
    assert rel(dE, ref) < 5e-6
    assert rel(dWih, dG.double().T @ X.double()) < 5e-6
    dG3, O3 = dG.double().view(n_seq, T, 8 * H), OUT[:M].double().view(n_seq, T, 2 * H)
    assert rel(dWhh[0], torch.einsum('stn,stk->nk', dG3[:, 1:, :4 * H], O3[:, :-1, :H])) < 5e-6
    assert rel(dWhh[1], torch.einsum('stn,stk->nk', dG3[:, :-1, 4 * H:], O3[:, 1:, H:])) < 5e-6
    # strided views: the rows t = T-1 of every sequence, reverse-direction gate columns ('last' aggregator, SubGNN.py:83)
    y = torch.zeros(M, 8 * H, device='cuda')
    bias = torch.randn(8 * H, generator=g).cuda()
    d1 = _abi.gemm_desc(_abi.GEMM_FWD, _ptr_off(Xd, (T - 1) * D), T * D, _ptr_off(Wd, 4 * H * D), D, _ptr_off(y, (T - 1) * 8 * H + 4 * H), T * 8 * H,
                        n_seq, 4 * H, D, bias=_ptr_off(bias, 4 * H))
    d0 = _abi.gemm_desc(_abi.GEMM_FWD, Xd.data_ptr(), D, Wd.data_ptr(), D, y.data_ptr(), 8 * H, M, 4 * H, D, bias=bias.data_ptr())
    _abi.gemm_group([d0, d1], _abi.stream_ptr(), max_ctas=40)
    torch.cuda.synchronize()
    want = torch.zeros(M, 8 * H, dtype=torch.float64)
    full = X.double() @ W.double().T + bias.cpu().double()
    want[:, :4 * H] = full[:, :4 * H]
    last = torch.arange(n_seq) * T + T - 1
    want[last, 4 * H:] = full[last, 4 * H:]
    assert float((y.cpu().double() - want).abs().max()) / float(want.abs().max()) < 5e-6
