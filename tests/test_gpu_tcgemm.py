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
