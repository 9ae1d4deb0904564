"""GPU parity of the fused step engine against (a) the committed golden runs of the UNMODIFIED reference
(tests/golden/model_*.npz) and (b) the CPU oracle on the same inputs.

Tolerance (stated): fp32, rtol 1e-4 / atol 1e-5 on logits, loss, gradients and weights after 4 Adam steps
(different summation order than torch; no reduced precision anywhere on the path)."""
import numpy as np
import pytest
import torch

from tests.util import golden_model, state_from

pytestmark = pytest.mark.gpu

CONFIGS = ['all_L1_max', 'all_L2_sum', 'S_L2_sumagg', 'NP_L2_trainable']


def make_engine(name):
    from subgnn_b200.engine import Engine
    hp, prepared, raw = golden_model(name)
    eng = Engine(hp, prepared, device='cuda', seed=1)
    eng.arena.load_state_dict(state_from(raw, 'init/'))
    if eng.arena.cc_tables:
        eng.tables_for('val')
        eng.snapshot_eval_cc_tables('val')
    return hp, prepared, raw, eng


@pytest.mark.parametrize('name', CONFIGS)
@pytest.mark.parametrize('use_graph', [False, True])
def test_training_matches_reference_golden(name, use_graph):
    hp, prepared, raw, eng = make_engine(name)
    for it in range(4):
        idx = raw['step/%d/idx' % it]
        loss = eng.train_step(idx, use_graph=use_graph)
        c = eng.context('train', len(idx), True)
        torch.cuda.synchronize()
        np.testing.assert_allclose(c.logits.cpu().numpy(), raw['step/%d/logits' % it], rtol=1e-4, atol=1e-5, err_msg='logits step %d' % it)
        np.testing.assert_allclose(float(loss.item()), float(raw['step/%d/loss' % it]), rtol=1e-4, err_msg='loss step %d' % it)
        gn = float(torch.sqrt(c.sumsq).item())
        np.testing.assert_allclose(gn, float(raw['step/%d/grad_norm' % it]), rtol=2e-4, err_msg='grad norm step %d' % it)
        if it == 0:
            coef = min(1.0, hp['grad_clip'] / (float(raw['step/0/grad_norm']) + 1e-6))
            for k in eng.arena.entries:
                if 'grad0/' + k in raw:
                    got = eng.arena.view(k, 'grads').cpu().numpy() * coef
                    np.testing.assert_allclose(got, raw['grad0/' + k], rtol=2e-4, atol=2e-6, err_msg='grad ' + k)
    final = state_from(raw, 'final/')
    for k in eng.arena.entries:
        np.testing.assert_allclose(eng.arena.view(k).cpu().numpy(), final[k].numpy(), rtol=1e-4, atol=1e-5, err_msg='final ' + k)


@pytest.mark.parametrize('name', CONFIGS)
def test_validation_forward_matches_reference_golden(name):
    hp, prepared, raw, eng = make_engine(name)
    eng.arena.load_state_dict(state_from(raw, 'final/'))
    n_val = len(prepared['labels']['val'])
    logits, loss = eng.forward('val', np.arange(n_val), training=False)
    torch.cuda.synchronize()
    np.testing.assert_allclose(logits.cpu().numpy(), raw['val/logits'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(float(loss.item()), float(raw['val/loss']), rtol=1e-4)


def test_engine_matches_oracle_with_hop_table_resolution():
    """Same as above but similarities are resolved on the GPU from the uint8 hop table (production path),
    and the comparison is against the CPU oracle stepping the same batches."""
    from oracle.model import OracleSubGNN
    from subgnn_b200.engine import Engine
    from subgnn_b200.graph import DeviceGraph
    hp, prepared, raw = golden_model('all_L2_sum')
    g = DeviceGraph.from_edges(prepared['n_nodes'], prepared['edges'], one_indexed=False).set_hop_table(prepared['hop'])
    p2 = dict(prepared)
    p2['NP_sim'] = None
    eng = Engine(hp, p2, device='cuda', graph=g, seed=3)
    eng.arena.load_state_dict(state_from(raw, 'init/'))
    torch.manual_seed(0)
    m = OracleSubGNN(hp, prepared)
    m.load_state_dict(state_from(raw, 'init/'))
    opt = torch.optim.Adam(m.parameters(), lr=hp['learning_rate'])
    rnd = np.random.RandomState(0)
    n_train = len(prepared['labels']['train'])
    for it in range(3):
        idx = rnd.choice(n_train, size=7, replace=False)
        loss_o, logits_o = m.training_step(m.make_batch('train', idx))
        opt.zero_grad()
        loss_o.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), hp['grad_clip'])
        opt.step()
        loss = eng.train_step(idx)
        c = eng.context('train', 7, True)
        torch.cuda.synchronize()
        np.testing.assert_allclose(c.logits.cpu().numpy(), logits_o.detach().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(float(loss.item()), float(loss_o.detach()), rtol=1e-4)
    sd = m.state_dict()
    for k in eng.arena.entries:
        np.testing.assert_allclose(eng.arena.view(k).cpu().numpy(), sd[k].numpy(), rtol=1e-4, atol=1e-5, err_msg=k)
