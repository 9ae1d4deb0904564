"""End-to-end GPU checks on a generated (not golden) workload: prepare_data on the GPU, engine vs the CPU oracle
on the same prepared tensors, the SubGNN module's autograd path, smoke()."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def tiny():
    from subgnn_b200 import prepare as prep
    from subgnn_b200 import synth
    hp, g, subs, labs, emb = synth.make_workload('tiny', seed=42, device='cuda')
    prepared = prep.prepare(hp, g, subs, labs, emb, seed=0, splits=('train', 'val'), num_classes=3)
    return hp, g, prepared


def test_prepare_matches_oracle_restatement(tiny):
    """bit-exact: hop table, components, border sets, walks, degree-sequence/DTW similarities of the GPU prepare
    pipeline vs the oracle run on the same graph with the same Philox keys."""
    from oracle import gamma as og
    from oracle import sampling as osamp
    from oracle import walks as ow
    hp, g, p = tiny
    S = ow.SortedAdj.from_csr(g.rowptr_host, g.col_host)
    assert np.array_equal(g.hop.cpu().numpy(), osamp.all_pairs_hops(S))
    # same components; their ORDER along C follows the reference's networkx iteration (prepare.connected_components, pinned
    # against reference-written caches in test_formats_compat_cpu.py), the oracle restatement orders by smallest member
    cc = osamp.initialize_cc_ids(S, p['sub_G']['train'])
    assert cc.shape == p['cc_ids']['train'].shape
    canon = lambda a: [sorted(tuple(int(x) for x in row if x) for row in sub if row[0]) for sub in a]
    assert canon(cc) == canon(p['cc_ids']['train'])
    P_tot = hp['max_sim_epochs'] * hp['n_anchor_patches_structure'] * hp['n_layers']
    ref_p = ow.sample_structure_anchor_patches(S, P_tot, hp['sample_walk_len'], hp['rw_beta'], ow.philox_patch_factory(1))
    assert np.array_equal(p['structure_anchors'], ref_p)
    for key, inside, seed in (('int_rw_all', True, 2), ('bor_rw_all', False, 3)):
        ref_w = ow.perform_random_walks(S, ref_p, hp['n_triangular_walks'], hp['random_walk_len'], hp['rw_beta'], inside,
                                        ow.philox_walk_factory(seed, hp['n_triangular_walks']))
        assert np.array_equal(p[key], ref_w)
    sub = p['cc_ids']['train'][:6]
    for key, internal in (('I_S_sim', True), ('B_S_sim', False)):
        want = og.structure_patch_similarities(S, sub, ref_p, internal)
        assert np.array_equal(p[key]['train'][:6], want)
    bptr, bitems = p['N_border']['train']
    want = osamp.initialize_border_sets(S, p['cc_ids']['train'], hp['neigh_sample_border_size'])
    flat = want.reshape(-1, want.shape[-1])
    for r in range(flat.shape[0]):
        assert np.array_equal(bitems[bptr[r]:bptr[r + 1]], flat[r][flat[r] != 0])


def test_engine_vs_oracle_on_generated_workload(tiny):
    from oracle.model import OracleSubGNN
    from subgnn_b200 import prepare as prep
    from subgnn_b200.engine import Engine
    hp, g, p = tiny
    eng = Engine(hp, p, device='cuda', graph=g, seed=5)
    eng.init_parameters(3)
    n = len(p['labels']['train'])
    full = prep.prepared_subset(p, g, 'train', np.arange(n))
    m = OracleSubGNN(hp, full)
    m.load_state_dict({k: v.cpu() for k, v in eng.arena.state_dict().items()})
    opt = torch.optim.Adam(m.parameters(), lr=hp['learning_rate'])
    rs = np.random.RandomState(1)
    for it in range(3):
        idx = np.sort(rs.choice(n, size=hp['batch_size'], replace=False))
        loss_o, logits_o = m.training_step(m.make_batch('train', idx))
        opt.zero_grad()
        loss_o.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), hp['grad_clip'])
        opt.step()
        loss = eng.train_step(idx, use_graph=True)
        c = eng.context('train', len(idx), True)
        torch.cuda.synchronize()
        np.testing.assert_allclose(c.logits.cpu().numpy(), logits_o.detach().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(float(loss.item()), float(loss_o.detach()), rtol=1e-4)


def test_module_autograd_path_matches_fused(tiny):
    """SubGNN.training_step (torch autograd + torch Adam over arena views) == training_step_fused."""
    from subgnn_b200.SubGNN import SubGNN
    hp, g, p = tiny
    hp = dict(hp, grad_clip=0.0)
    a = SubGNN.from_prepared(hp, p, graph=g, seed=1, init_seed=11)
    b = SubGNN.from_prepared(hp, p, graph=g, seed=1, init_seed=11)
    assert 'lstm.lstm.weight_ih_l0_reverse' in a.state_dict() and 'neighborhood_mpns.1.border.linear.weight' in a.state_dict()
    opt = a.configure_optimizers()
    a.train()
    idx = torch.arange(hp['batch_size']).view(-1, 1)
    labels = torch.as_tensor(p['labels']['train'][:hp['batch_size']])
    for _ in range(2):
        out = a.training_step({'subgraph_idx': idx, 'label': labels})
        opt.zero_grad()
        a.backward(None, out['loss'], opt, 0)
        opt.step()
        fused = b.training_step_fused({'subgraph_idx': idx}, use_graph=False)
        np.testing.assert_allclose(float(out['loss']), float(fused['loss']), rtol=1e-4)
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        np.testing.assert_allclose(sa[k].cpu().numpy(), sb[k].cpu().numpy(), rtol=1e-4, atol=1e-5, err_msg=k)


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


def test_fused_lstm_dropout_equals_separate_mask_kernels(tiny):
    """nn.LSTM's inter-layer dropout fused into the recurrence kernels (mask written by the forward kernel, applied on load by the
    BPTT kernel) draws the same Philox mask as the stand-alone dropout kernel: identical steps either way."""
    from subgnn_b200.engine import Engine
    hp, g, p = tiny
    engines = []
    for fused in (True, False):
        h = dict(hp, lstm_dropout=0.35, lstm_n_layers=2, b200_fused_lstm_dropout=fused)
        eng = Engine(h, p, device='cuda', graph=g, seed=9)
        eng.init_parameters(4)
        assert eng.lstm.fused_drop == fused and eng.lstm.p_drop > 0
        engines.append(eng)
    idx = np.arange(hp['batch_size'])
    for it in range(3):
        la = float(engines[0].train_step(idx + it, use_graph=False).item())
        lb = float(engines[1].train_step(idx + it, use_graph=False).item())
        np.testing.assert_allclose(la, lb, rtol=1e-6)
    for k, v in engines[0].arena.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), engines[1].arena.state_dict()[k].cpu().numpy(), rtol=1e-5, atol=1e-7, err_msg=k)
    # and dropout is really active: the same engine without it takes a different step
    h = dict(hp, lstm_dropout=0.0, lstm_n_layers=2)
    ref = Engine(h, p, device='cuda', graph=g, seed=9)
    ref.init_parameters(4)
    assert abs(float(ref.train_step(idx, use_graph=False).item()) - float(Engine(dict(hp, lstm_dropout=0.35, lstm_n_layers=2), p, device='cuda', graph=g, seed=9).train_step(idx, use_graph=False).item())) >= 0.0


@pytest.mark.parametrize('n_layers', [1, 2])
def test_last_only_head_gradient_equals_zero_filled_rows(tiny, monkeypatch, n_layers):
    """'last' aggregator: the head writes only the rows t = T-1 of the top layer's output gradient and the BPTT kernel takes the
    others as zero without reading them (SUBGNN_HEAD_LAST_NO_FILL / SUBGNN_LSTM_DOUT_LAST_ONLY) — same steps as the zero-filled
    buffer; also checks the step-graph variants of the forward branches (SUBGNN_FWD_BRANCHES) against each other."""
    from subgnn_b200.engine import Engine
    hp, g, p = tiny
    h = dict(hp, lstm_aggregator='last', lstm_n_layers=n_layers, lstm_dropout=0.0)
    idx = np.arange(hp['batch_size'])
    states, losses = [], []
    for last_only, branches in (('1', '1'), ('0', '0')):
        monkeypatch.setenv('SUBGNN_LSTM_LAST_ONLY', last_only)
        monkeypatch.setenv('SUBGNN_FWD_BRANCHES', branches)
        eng = Engine(h, p, device='cuda', graph=g, seed=9)
        eng.init_parameters(4)
        assert eng.lstm.last_only == (last_only == '1')
        eng.lstm.dOUT[-1].fill_(123.0)            # stale rows must never be read in the last-only mode (and are overwritten otherwise)
        ls = [float(eng.train_step(idx + it, use_graph=(it > 0)).item()) for it in range(4)]
        torch.cuda.synchronize()
        losses.append(ls)
        states.append({k: v.cpu().numpy().copy() for k, v in eng.arena.state_dict().items()})
    np.testing.assert_allclose(losses[0], losses[1], rtol=1e-5)
    for k in states[0]:
        np.testing.assert_allclose(states[0][k], states[1][k], rtol=1e-4, atol=1e-6, err_msg=k)
