"""GPU parity of the module-level drop-ins (SG_MPN, LSTM walk encoder, get_anchor_patches, gamma helpers) against the
CPU oracle's restatement of the same reference modules.  fp32 tolerance rtol 1e-4 / atol 1e-5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mpn_inputs(B=5, C=3, A=7, D=40, n_opt=23, structure=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    cc_ids = torch.randint(1, 20, (B, C, 4), generator=g)
    cc_ids[1, 2] = 0
    cc_ids[3, 1:] = 0                                    # padded components
    cc_mask = cc_ids[:, :, 0] != 0
    cc_embeds = torch.randn(B, C, D, generator=g)
    anchor_embeds = torch.randn(B, C, A, D, generator=g)
    sims = torch.rand(B, C, n_opt, generator=g)
    if structure:
        patches = torch.randint(1, n_opt, (A, 5), generator=g).unsqueeze(0).unsqueeze(0).repeat(B, C, 1, 1)
        patches[~cc_mask] = 0
        sim_index = torch.randint(0, n_opt, (A,), generator=g).tolist()
    else:
        patches = torch.randint(1, n_opt + 1, (B, C, A, 1), generator=g)
        patches[0, 0, 2] = 0                             # a PAD anchor inside a valid component
        patches[~cc_mask] = 0
        sim_index = None
    mask = patches != 0
    return cc_ids, cc_mask, cc_embeds, anchor_embeds, sims, patches, mask, sim_index


@pytest.mark.parametrize('structure', [False, True])
@pytest.mark.parametrize('use_proj', [True, False])
def test_sg_mpn_forward_backward(structure, use_proj):
    from oracle.model import SG_MPN as OracleMPN
    from subgnn_b200.subgraph_mpn import SG_MPN
    hp = {'node_embed_size': 40, 'use_mpn_projection': use_proj}
    cc_ids, cc_mask, cc, ae, sims, patches, mask, sim_index = _mpn_inputs(structure=structure)
    torch.manual_seed(1)
    ref = OracleMPN(hp)
    mine = SG_MPN(hp)
    mine.load_state_dict(ref.state_dict())
    cc_r, ae_r = cc.clone().requires_grad_(True), ae.clone().requires_grad_(True)
    out_r, pos_r = ref(sims, cc_ids, cc_r, cc_mask, patches, ae_r, mask, sim_index)
    cc_g, ae_g = cc.cuda().requires_grad_(True), ae.cuda().requires_grad_(True)
    out_g, pos_g = mine(None, sims.cuda(), cc_ids.cuda(), cc_g, cc_mask.cuda(), patches.cuda(), ae_g, mask.cuda(), sim_index)
    np.testing.assert_allclose(out_g.detach().cpu().numpy(), out_r.detach().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(pos_g.detach().cpu().numpy(), pos_r.detach().numpy(), rtol=1e-4, atol=1e-5)
    w1, w2 = torch.randn_like(out_r), torch.randn_like(pos_r)
    ((out_r * w1).sum() + (pos_r * w2).sum()).backward()
    ((out_g * w1.cuda()).sum() + (pos_g * w2.cuda()).sum()).backward()
    np.testing.assert_allclose(cc_g.grad.cpu().numpy(), cc_r.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ae_g.grad.cpu().numpy(), ae_r.grad.numpy(), rtol=1e-4, atol=1e-5)
    for (n, p_r), (_, p_g) in zip(ref.named_parameters(), mine.named_parameters()):
        if p_r.grad is None:
            assert p_g.grad is None or float(p_g.grad.abs().max()) == 0.0, n
        else:
            np.testing.assert_allclose(p_g.grad.cpu().numpy(), p_r.grad.numpy(), rtol=1e-4, atol=1e-5, err_msg=n)


@pytest.mark.parametrize('layers,agg,H', [(1, 'last', 32), (2, 'last', 24), (2, 'sum', 64), (1, 'sum', 128)])
def test_lstm_walk_encoder(layers, agg, H):
    from oracle.model import LSTM as OracleLSTM
    from subgnn_b200.SubGNN import LSTM
    torch.manual_seed(3)
    ref = OracleLSTM(H, H, dropout=0.0, num_layers=layers, aggregator=agg)
    mine = LSTM(H, H, dropout=0.0, num_layers=layers, aggregator=agg).cuda()
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(37, 9, H)
    x[5, 4:] = 0                                         # PAD steps are zero vectors
    xr, xg = x.clone().requires_grad_(True), x.cuda().requires_grad_(True)
    yr, yg = ref(xr), mine(xg)
    np.testing.assert_allclose(yg.detach().cpu().numpy(), yr.detach().numpy(), rtol=1e-4, atol=1e-5)
    w = torch.randn_like(yr)
    (yr * w).sum().backward()
    (yg * w.cuda()).sum().backward()
    np.testing.assert_allclose(xg.grad.cpu().numpy(), xr.grad.numpy(), rtol=2e-4, atol=2e-5)
    for (n, p_r), (_, p_g) in zip(ref.named_parameters(), mine.named_parameters()):
        np.testing.assert_allclose(p_g.grad.cpu().numpy(), p_r.grad.numpy(), rtol=2e-4, atol=2e-5, err_msg=n)


def test_get_anchor_patches_and_gamma_api():
    """reference-named entry points: get_anchor_patches (all channels) vs the oracle; gamma helpers vs the oracle."""
    import networkx as nx
    from oracle import gamma as og
    from oracle import walks as ow
    from oracle.model import OracleSubGNN
    from subgnn_b200 import anchor_patch_samplers as aps
    from subgnn_b200 import gamma
    from subgnn_b200.SubGNN import LSTM
    from tests.util import golden_model, state_from
    hp, p, raw = golden_model('all_L2_sum')
    m = OracleSubGNN(hp, p)
    m.load_state_dict(state_from(raw, 'init/'))
    batch = m.make_batch('train', raw['step/0/idx'])
    cc_ids, sub_idx = batch['cc_ids'], batch['subgraph_idx']
    cc_mask = (cc_ids != 0)[:, :, 0]
    node_matrix = torch.nn.Embedding.from_pretrained(m.node_embeddings.weight.detach().clone(), padding_idx=0).cuda()
    lstm = LSTM(hp['node_embed_size'], hp['node_embed_size'], dropout=0.0, num_layers=hp['lstm_n_layers'], aggregator=hp['lstm_aggregator']).cuda()
    lstm.load_state_dict(m.lstm.state_dict())
    t = lambda d: {s: {l: torch.as_tensor(v) for l, v in dd.items()} for s, dd in d.items()}
    struct = {l: (torch.as_tensor(v[0]), v[1], torch.as_tensor(v[2]), torch.as_tensor(v[3])) for l, v in p['anchors_structure'].items()}
    pos_ext = {l: torch.as_tensor(v) for l, v in p['anchors_pos_ext'].items()}
    for channel in ('neighborhood', 'position', 'structure'):
        for inside in (True, False):
            want = m.get_anchor_patches('train', sub_idx, cc_ids, cc_mask, 1, channel, inside)
            got = aps.get_anchor_patches('train', hp, None, node_matrix, sub_idx, cc_ids.cuda(), cc_mask.cuda(), lstm, t(p['anchors_neigh_int']),
                                         t(p['anchors_neigh_border']), t(p['anchors_pos_int']), pos_ext, struct, 1, channel, inside, torch.device('cuda'))
            assert torch.equal(got[0].cpu(), want[0]) and torch.equal(got[1].cpu(), want[1])
            np.testing.assert_allclose(got[2].detach().cpu().numpy(), want[2].detach().numpy(), rtol=1e-4, atol=1e-5)
    with pytest.raises(Exception):
        aps.get_anchor_patches('train', hp, None, node_matrix, sub_idx, cc_ids.cuda(), cc_mask.cuda(), lstm, None, None, None, None, None, 0, 'bogus', True)
    # gamma: degree sequences through a networkx graph handle, calc_dist / calc_dtw
    G = nx.Graph()
    G.add_nodes_from(range(1, p['n_nodes'] + 1))
    G.add_edges_from((int(u) + 1, int(v) + 1) for u, v in p['edges'])
    S = ow.SortedAdj(p['n_nodes'], [(int(u) + 1, int(v) + 1) for u, v in p['edges']])
    nodes = torch.tensor([5, 9, 9, 0, 12, 3, 0])
    for internal in (True, False):
        assert gamma.get_degree_sequence(G, nodes, None, internal) == og.get_degree_sequence(S, nodes.numpy(), internal)
    assert gamma.calc_dist(3, 1) == og.calc_dist(3, 1)
    x, y = [0, 1, 1, 2, 4, 4, 7], [1, 1, 2, 2, 3, 5, 5, 6, 9]
    assert gamma.calc_dtw(x, y) == float(np.float32(og.calc_dtw(x, y)))
    assert gamma.calc_dtw([], y) == 0.0
