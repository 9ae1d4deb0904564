"""Pins the oracle (oracle/) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import json
import random

import networkx as nx
import numpy as np
import pytest
import torch

from oracle import gamma as og
from oracle import sampling as osamp
from oracle import walks as ow
from oracle.model import OracleSubGNN
from oracle.rng import MTStream, philox4x32_10
from tests.util import golden_model, load_npz, state_from


def nx_graph(n_nodes, edges):
    G = nx.Graph()
    G.add_nodes_from(range(1, n_nodes + 1))
    G.add_edges_from((int(u) + 1, int(v) + 1) for u, v in edges)
    return G


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10, counter=0, key=0 and the all-ones vector
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)


def test_walks_match_reference_streams():
    g = load_npz('walks_golden.npz')
    hp = json.loads(str(g['hparams_json']))
    G = ow.NxAdapter(nx_graph(int(g['n_nodes']), g['edges']))
    n_samples = 2 * hp['n_anchor_patches_structure'] * hp['n_layers']
    for seed in (0, 1, 2):
        np.random.seed(seed)
        random.seed(seed)
        np.random.choice(G.all_nodes(), n_samples, replace=True)   # anchor_patch_samplers.py:222 (drawn, unused)
        patches = ow.sample_structure_anchor_patches(G, n_samples, hp['sample_walk_len'], hp['rw_beta'], lambda i: MTStream())
        assert np.array_equal(patches, g['patches/%d' % seed])
        np.random.seed(seed + 100)
        random.seed(seed + 100)
        irw = ow.perform_random_walks(G, patches, hp['n_triangular_walks'], hp['random_walk_len'], hp['rw_beta'], True, lambda p, w: MTStream())
        assert np.array_equal(irw, g['int_rw/%d' % seed])
        np.random.seed(seed + 200)
        random.seed(seed + 200)
        brw = ow.perform_random_walks(G, patches, hp['n_triangular_walks'], hp['random_walk_len'], hp['rw_beta'], False, lambda p, w: MTStream())
        assert np.array_equal(brw, g['bor_rw/%d' % seed])


def test_border_nodes_match_reference():
    g = load_npz('walks_golden.npz')
    S = ow.SortedAdj(int(g['n_nodes']), [(int(u) + 1, int(v) + 1) for u, v in g['edges']])
    for i in range(4):
        nodes = g['border_patch/%d' % i]
        got = sorted(ow.border_nodes(S, S.patch(nodes).order))
        assert got == g['border_nodes/%d' % i].tolist()


def test_degree_sequences_bit_exact():
    g = load_npz('gamma_golden.npz')
    S = ow.SortedAdj(int(g['n_nodes']), [(int(u) + 1, int(v) + 1) for u, v in g['edges']])
    oi = np.concatenate([[0], np.cumsum(g['seq_int_len'])])
    ob = np.concatenate([[0], np.cumsum(g['seq_bor_len'])])
    for r, row in enumerate(g['rows']):
        assert og.get_degree_sequence(S, row, True) == g['seq_int'][oi[r]:oi[r + 1]].tolist()
        assert og.get_degree_sequence(S, row, False) == g['seq_bor'][ob[r]:ob[r + 1]].tolist()


def test_dtw_similarities():
    g = load_npz('gamma_golden.npz')
    oi = np.concatenate([[0], np.cumsum(g['seq_int_len'])])
    ob = np.concatenate([[0], np.cumsum(g['seq_bor_len'])])
    n = g['dtw_sims'].shape[0]
    for i in range(n):
        for j in range(n):
            x = g['seq_int'][oi[i]:oi[i + 1]].tolist()
            y = g['seq_bor'][ob[j]:ob[j + 1]].tolist()
            assert og.calc_dtw(x, y) == g['dtw_sims'][i, j]


def test_fastdtw_known_answers():
    # hand-derived: constant vs. step sequences, base cases, radius-1 >= exact
    assert og.calc_dist(3, 1) == 1.0 and og.calc_dist(0, 0) == 0.0 and og.calc_dist(1, 3) == 1.0
    # x = [2]*5, y=[0,0,1,1,1]: every alignment pays d(2,0)=2 twice and d(2,1)=.5 three times = 5.5
    d, _ = og.fastdtw([2, 2, 2, 2, 2], [0, 0, 1, 1, 1])
    assert d == 5.5
    assert og.calc_dtw([2, 2, 2, 2, 2], [0, 0, 1, 1, 1]) == 1 / 6.5
    assert og.fastdtw([1], [1, 1, 1])[0] == 0.0                     # length-1 base case (exact)
    assert og.fastdtw([1, 3], [3])[0] == 1.0                        # d(1,3) + d(3,3)
    assert og.fastdtw([0, 1, 2, 3, 4, 5, 6], [0, 1, 2, 3, 4, 5, 6])[0] == 0.0
    rnd = random.Random(5)
    for _ in range(300):
        x = sorted(rnd.randint(0, 30) for _ in range(rnd.randint(1, 25)))
        y = sorted(rnd.randint(0, 60) for _ in range(rnd.randint(1, 50)))
        fa, ex = og.fastdtw(x, y)[0], og.dtw_exact(x, y)[0]
        assert fa >= ex - 1e-12
        if len(x) < 3 or len(y) < 3:
            assert fa == ex


def test_exact_dtw_against_an_independent_formulation():
    """og.dtw_exact (row-by-row DP, the oracle of SUBGNN_DTW_EXACT / _THREAD) against a memoised top-down recursion over the
    definition D(i, j) = d(x_i, y_j) + min(D(i-1, j), D(i, j-1), D(i-1, j-1)) — a different evaluation order of the same sums."""
    import functools
    rnd = random.Random(11)
    for _ in range(200):
        x = sorted(rnd.randint(0, 40) for _ in range(rnd.randint(1, 18)))
        y = sorted(rnd.randint(0, 400) for _ in range(rnd.randint(1, 30)))

        @functools.lru_cache(maxsize=None)
        def D(i, j):
            c = og.calc_dist(x[i], y[j])
            if i == 0 and j == 0:
                return c
            best = float('inf')
            if i > 0:
                best = min(best, D(i - 1, j))
            if j > 0:
                best = min(best, D(i, j - 1))
            if i > 0 and j > 0:
                best = min(best, D(i - 1, j - 1))
            return c + best

        assert og.dtw_exact(x, y)[0] == D(len(x) - 1, len(y) - 1)
        assert og.calc_dtw(x, y, 'exact') == 1.0 / (1.0 + D(len(x) - 1, len(y) - 1))


def canon_cc(cc):
    out = []
    for sub in cc:
        comps = [tuple(sorted(int(n) for n in c if n != 0)) for c in sub]
        out.append(sorted(c for c in comps if c))
    return out


def test_cc_ids_border_sets_spmin():
    g = load_npz('sampling_golden.npz')
    S = ow.SortedAdj(int(g['n_nodes']), [(int(u) + 1, int(v) + 1) for u, v in g['edges']])
    offs = np.concatenate([[0], np.cumsum(g['sub_len'])])
    subs = [g['sub_flat'][offs[i]:offs[i + 1]].tolist() for i in range(len(g['sub_len']))]
    cc = osamp.initialize_cc_ids(S, subs)
    assert cc.shape == g['cc_ids'].shape
    assert canon_cc(cc) == canon_cc(g['cc_ids'])
    for k in (1, 2):
        got = osamp.initialize_border_sets(S, g['cc_ids'], k)
        assert np.array_equal(got, g['border_k%d' % k])
    assert np.array_equal(osamp.all_pairs_hops(S), g['hop'])
    sim = osamp.shortest_path_similarities(g['hop'], g['cc_ids'])
    assert np.array_equal(sim, g['sp_sim'])


def test_neighborhood_sampling_reference_stream():
    g = load_npz('sampling_golden.npz')
    torch.manual_seed(123)
    got = osamp.sample_neighborhood_anchor_patch_torch(torch.from_numpy(g['cc_ids']), 5)
    assert np.array_equal(got.numpy(), g['N_in_seed123'])


@pytest.mark.parametrize('name', ['all_L1_max', 'all_L2_sum', 'S_L2_sumagg', 'NP_L2_trainable'])
def test_model_matches_reference(name):
    hp, prepared, raw = golden_model(name)
    torch.manual_seed(0)
    m = OracleSubGNN(hp, prepared)
    missing = m.load_state_dict(state_from(raw, 'init/'), strict=True)
    m.snapshot_eval_cc_tables()
    opt = torch.optim.Adam(m.parameters(), lr=hp['learning_rate'])
    m.train()
    for it in range(4):
        batch = m.make_batch('train', raw['step/%d/idx' % it])
        loss, logits = m.training_step(batch)
        np.testing.assert_allclose(logits.detach().numpy(), raw['step/%d/logits' % it], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(float(loss.detach()), float(raw['step/%d/loss' % it]), rtol=1e-5)
        opt.zero_grad()
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(m.parameters(), hp['grad_clip'])
        np.testing.assert_allclose(float(gn), float(raw['step/%d/grad_norm' % it]), rtol=1e-4)
        if it == 0:
            for k, prm in m.named_parameters():
                if 'grad0/' + k in raw:
                    np.testing.assert_allclose(prm.grad.numpy(), raw['grad0/' + k], rtol=1e-4, atol=1e-6, err_msg=k)
        opt.step()
    final = state_from(raw, 'final/')
    for k, v in m.state_dict().items():
        np.testing.assert_allclose(v.numpy(), final[k].numpy(), rtol=1e-4, atol=1e-5, err_msg=k)
    m.eval()
    with torch.no_grad():
        vb = m.make_batch('val', np.arange(len(prepared['labels']['val'])))
        logits = m.forward('val', vb)
    np.testing.assert_allclose(logits.numpy(), raw['val/logits'], rtol=1e-4, atol=1e-5)
