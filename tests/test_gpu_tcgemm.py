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
