"""Optimizer entry points (optim.cu) against torch.optim.Adam + torch.nn.utils.clip_grad_norm_ — the pair the reference runs
(SubGNN.py:1156-1164 Adam(lr); pytorch-lightning gradient_clip_val = grad_clip, train_config.py:109-158) — on a flat arena:
the two-kernel path (subgnn_grad_sumsq + subgnn_adam_step) and the one-launch path (subgnn_clip_adam_step: grid barrier between
the norm and the update).  fp32; tolerance rtol 1e-5 / atol 1e-7 on the parameters after 3 steps (stated), and the two paths
must agree with each other to the same bound at every size, including sizes that are not a multiple of 4, sizes below one CTA
and slices longer than the kernel's register window."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(entry, p0, grads, lr, clip, scale):
    from subgnn_b200._abi import call, ptr
    dev = 'cuda'
    n = p0.numel()
    p = p0.clone().to(dev)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    norms = []
    st = torch.cuda.current_stream().cuda_stream
    for g in grads:
        g = g.to(dev)
        sumsq = torch.zeros(1, device=dev)
        call('subgnn_inc_step', ptr(step), st)
        if entry == 'fused':
            call('subgnn_clip_adam_step', ptr(p), ptr(g), ptr(m), ptr(v), n, lr, 0.9, 0.999, 1e-8, ptr(step), ptr(sumsq), clip, scale, st)
        else:
            call('subgnn_grad_sumsq', ptr(g), n, ptr(sumsq), st)
            call('subgnn_adam_step', ptr(p), ptr(g), ptr(m), ptr(v), n, lr, 0.9, 0.999, 1e-8, ptr(step), ptr(sumsq), clip, scale, st)
        torch.cuda.synchronize()
        norms.append(float(torch.sqrt(sumsq).item()))
    return p.cpu(), m.cpu(), v.cpu(), norms


@pytest.mark.parametrize('n', [1, 3, 1000, 4099, 148 * 512 * 4 + 5, 1_580_000, 148 * 512 * 4 * 7 + 2])
@pytest.mark.parametrize('clip', [0.25, 0.0])
def test_clip_adam_paths_match_torch(n, clip):
    gen = torch.Generator().manual_seed(n % 1000 + 7)
    p0 = torch.randn(n, generator=gen)
    grads = [torch.randn(n, generator=gen) * (0.01 if i == 1 else 1.0) for i in range(3)]     # step 1 stays below the clip norm
    lr, scale = 1e-3, 0.5
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=lr)
    ref_norms = []
    for g in grads:
        ref.grad = g.clone() * scale                        # grad_scale: the 1 / world_size average applied before the norm
        ref_norms.append(float(torch.linalg.vector_norm(ref.grad.double())))
        if clip > 0:
            torch.nn.utils.clip_grad_norm_([ref], clip)
        opt.step()
    out = {}
    for entry in ('split', 'fused'):
        p, m, v, norms = _run(entry, p0, grads, lr, clip, scale)
        out[entry] = (p, m, v)
        np.testing.assert_allclose(np.array(norms) * scale, ref_norms, rtol=1e-5, err_msg=entry + ' gradient norm')
        np.testing.assert_allclose(p.numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-7, err_msg=entry + ' parameters')
    for a, b, what in zip(out['split'], out['fused'], ('p', 'm', 'v')):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-5, atol=1e-7, err_msg='split vs fused ' + what)


def test_fused_step_is_repeatable_bit_for_bit():
    """the fused kernel's norm is a fixed-order sum (per-CTA partials, same order in every CTA): two runs give identical bits"""
    n = 1_234_567
    gen = torch.Generator().manual_seed(3)
    p0 = torch.randn(n, generator=gen)
    grads = [torch.randn(n, generator=gen) for _ in range(2)]
    a = _run('fused', p0, grads, 1e-3, 0.25, 1.0)
    b = _run('fused', p0, grads, 1e-3, 0.25, 1.0)
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y)
    assert a[3] == b[3]
