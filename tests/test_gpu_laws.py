"""Laws (distributional / structural properties) of the stochastic and third-party-defined stages on the GPU:

  * dropout (SubGNN.py:306-310 F.dropout on the readout MLP, nn.LSTM inter-layer dropout :73): keep rate 1 - p, survivors scaled
    by exactly 1 / (1 - p), masks of the two MLP layers independent, a fresh mask every training step — also on the autograd path;
  * fastdtw (gamma.py:54-59): radius-1 fastdtw distance >= exact DTW distance on every pair, equality whenever either sequence
    is shorter than 3 (fastdtw's own base case), on > 5 000 random pairs — what IS proven while the third-party package's
    tie-breaking stays unpinned;
  * anchor-patch walks (anchor_patch_samplers.py:49-158): unique-node patch-size histogram, per-node visit frequency and walk
    length of the GPU Philox sampler against the reference-stream (numpy MT / python random) sampler, for full-graph, inside-patch
    and border walks;
  * init_anchors_* containers (anchor_patch_samplers.py:248-328): shapes, membership and cross-process reproducibility.
"""
import random

import networkx as nx
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------- dropout
@pytest.fixture(scope='module')
def wide():
    """tiny-shaped workload with enough subgraphs for a 256-sample batch and the benchmark's 64-wide MLP."""
    from subgnn_b200 import prepare as prep
    from subgnn_b200 import synth
    hp, g, subs, labs, emb = synth.make_workload('tiny', seed=42, device='cuda', n_sub=400)
    hp = dict(hp, linear_hidden_dim_1=64, linear_hidden_dim_2=64, lstm_n_layers=2)
    prepared = prep.prepare(hp, g, subs, labs, emb, seed=0, splits=('train',), num_classes=3)
    return hp, g, prepared


def _binom_tol(p, n, z=5.0):
    return z * np.sqrt(p * (1 - p) / n)


def test_readout_dropout_law(wide):
    from subgnn_b200.engine import Engine
    hp, g, p = wide
    pd = 0.3681
    eng = Engine(dict(hp, lin_dropout=pd, lstm_dropout=0.0), p, device='cuda', graph=g, seed=77)
    eng.init_parameters(5)
    B = 256
    idx = np.arange(B)
    eng.forward('train', idx, training=False)
    ce = eng.context('train', B, False)
    torch.cuda.synchronize()
    h1_eval = ce.H1.cpu().numpy().copy()                         # relu(W1 z + b1), no dropout
    masks1, masks2 = [], []
    for step in range(2):
        eng.forward('train', idx, training=True)
        ct = eng.context('train', B, True)
        torch.cuda.synchronize()
        h1 = ct.H1.cpu().numpy().copy()                          # post-relu post-dropout
        h2 = ct.H2.cpu().numpy().copy()
        live1 = h1_eval > 1e-6
        keep1 = h1 != 0
        assert (keep1 & ~live1).sum() <= 2                       # dropout never revives a relu-dead unit (split-K atomics: a unit at ~0 may flip)
        np.testing.assert_allclose(h1[keep1 & live1], (h1_eval / (1 - pd))[keep1 & live1], rtol=1e-4, atol=1e-6)   # survivors scaled by 1/(1-p)
        rate1 = keep1[live1].mean()
        assert abs(rate1 - (1 - pd)) < _binom_tol(pd, live1.sum()), rate1
        # second layer: its input is the DROPPED h1; pre-dropout value recomputed on the host from the same weights
        W2 = eng.arena.view('lin2.weight').cpu().numpy().astype(np.float64)
        b2 = eng.arena.view('lin2.bias').cpu().numpy().astype(np.float64)
        pre2 = np.maximum(h1.astype(np.float64) @ W2.T + b2, 0.0)
        live2 = pre2 > 1e-5
        keep2 = h2 != 0
        assert not np.any(keep2 & (pre2 < -1e-5))
        np.testing.assert_allclose(h2[keep2 & live2], (pre2 / (1 - pd))[keep2 & live2], rtol=1e-4, atol=1e-6)
        rate2 = keep2[live2].mean()
        assert abs(rate2 - (1 - pd)) < _binom_tol(pd, live2.sum()), rate2
        masks1.append((keep1, live1))
        masks2.append((keep2, live2))
    # independence of the two layers' masks (same (sample, unit) index space: h1 == h2 == 64): P(keep1 & keep2) = (1-p)^2
    both = masks1[0][1] & masks2[0][1]
    joint = (masks1[0][0] & masks2[0][0])[both].mean()
    assert abs(joint - (1 - pd) ** 2) < _binom_tol((1 - pd) ** 2, both.sum()), joint
    agree = (masks1[0][0] == masks2[0][0])[both].mean()
    assert agree < 0.62                                          # identical masks would agree everywhere; independent ones on 1-2p(1-p) = 0.535
    # a fresh mask every training forward (the autograd path used to freeze the step counter)
    live = masks1[0][1]
    same = (masks1[0][0] == masks1[1][0])[live].mean()
    assert same < 0.62, same


def test_lstm_interlayer_dropout_law(wide):
    from subgnn_b200.engine import Engine
    hp, g, p = wide
    pd = 0.2353
    for fused in (True, False):
        eng = Engine(dict(hp, lin_dropout=0.0, lstm_dropout=pd, b200_fused_lstm_dropout=fused), p, device='cuda', graph=g, seed=78)
        eng.init_parameters(5)
        keeps = []
        for step in range(2):
            eng.train_step(np.arange(hp['batch_size']), use_graph=False)
            torch.cuda.synchronize()
            ls = eng.lstm
            M = ls.n_seq * ls.T
            out0 = ls.OUT[0][:M].cpu().numpy()
            x1 = ls.X[1][:M].cpu().numpy()
            live = np.abs(out0) > 1e-7
            keep = x1 != 0
            np.testing.assert_allclose(x1[keep], (out0 / (1 - pd))[keep], rtol=1e-5)
            rate = keep[live].mean()
            assert abs(rate - (1 - pd)) < _binom_tol(pd, live.sum()), (fused, rate)
            keeps.append((keep, live))
        both = keeps[0][1] & keeps[1][1]
        same = (keeps[0][0] == keeps[1][0])[both].mean()
        assert same < 1 - 2 * pd * (1 - pd) + 0.02, same          # a new mask each step


def test_autograd_training_steps_draw_fresh_masks_and_match_fused(wide):
    """SubGNN.training_step (autograd path) advances the dropout counter like the fused step: with the same seeds the two paths
    see the same masks step after step, i.e. identical losses (engine.py:758)."""
    from subgnn_b200.SubGNN import SubGNN
    hp, g, p = wide
    h = dict(hp, lin_dropout=0.3, lstm_dropout=0.25, grad_clip=0.0)
    a = SubGNN.from_prepared(h, p, graph=g, seed=1, init_seed=11)
    b = SubGNN.from_prepared(h, p, graph=g, seed=1, init_seed=11)
    opt = a.configure_optimizers()
    a.train()
    B = hp['batch_size']
    idx = torch.arange(B).view(-1, 1)
    labels = torch.as_tensor(p['labels']['train'][:B])
    losses = []
    for it in range(3):
        out = a.training_step({'subgraph_idx': idx, 'label': labels})
        opt.zero_grad()
        a.backward(None, out['loss'], opt, 0)
        opt.step()
        fused = b.training_step_fused({'subgraph_idx': idx}, use_graph=False)
        np.testing.assert_allclose(float(out['loss']), float(fused['loss']), rtol=2e-4, err_msg='step %d' % it)
        losses.append(float(out['loss']))
    assert len(set(round(x, 6) for x in losses)) == 3
    # two forwards of the same context before a backward would read the later forward's buffers: refused loudly
    o1 = a.training_step({'subgraph_idx': idx, 'label': labels})
    a.training_step({'subgraph_idx': idx, 'label': labels})
    with pytest.raises(RuntimeError):
        o1['loss'].backward()


@pytest.mark.parametrize('H,layers,pd', [(32, 1, 0.0), (64, 2, 0.0), (20, 2, 0.0), (64, 2, 0.3)])
def test_lstm_module_two_calls_before_backward(H, layers, pd):
    """The reference calls the shared LSTM 2 * n_layers times per step before loss.backward() (anchor_patch_samplers.py:429 via
    get_anchor_patches): every outstanding call keeps its own saved activations (SubGNN.py:87).  H = 20 runs the
    general-H recurrence kernels (lstm.cu), the others the register-tiled ones."""
    from oracle.model import LSTM as OracleLSTM
    from subgnn_b200.SubGNN import LSTM
    torch.manual_seed(4)
    ref = OracleLSTM(H, H, dropout=0.0, num_layers=layers, aggregator='last')
    mine = LSTM(H, H, dropout=pd, num_layers=layers, aggregator='last').cuda()
    mine.load_state_dict(ref.state_dict())
    xs = [torch.randn(21, 7, H), torch.randn(21, 7, H), torch.randn(13, 5, H)]
    ws = [torch.randn(x.shape[0], H) for x in xs]
    if pd > 0:
        # with dropout there is no torch oracle for the mask: check that backward of call 1 after call 2 equals backward of call 1
        # run alone with the same counter (the mask a call's backward applies is the one ITS forward drew)
        mine.train()
        xg = xs[0].cuda().requires_grad_(True)
        y1 = mine(xg)
        g_alone = torch.autograd.grad((y1 * ws[0].cuda()).sum(), [xg] + list(mine.parameters()))
        mine._step.zero_()
        xg2 = xs[0].cuda().requires_grad_(True)
        y1b = mine(xg2)
        _ = mine(xs[1].cuda().requires_grad_(True))            # second outstanding call, other counter value, other buffers
        g_both = torch.autograd.grad((y1b * ws[0].cuda()).sum(), [xg2] + list(mine.parameters()))
        np.testing.assert_allclose(y1b.detach().cpu().numpy(), y1.detach().cpu().numpy(), rtol=1e-6)
        for u, v in zip(g_alone, g_both):
            np.testing.assert_allclose(v.cpu().numpy(), u.cpu().numpy(), rtol=1e-5, atol=1e-7)
        return
    xr = [x.clone().requires_grad_(True) for x in xs]
    xg = [x.cuda().requires_grad_(True) for x in xs]
    loss_r = sum((ref(x) * w).sum() for x, w in zip(xr, ws))
    loss_g = sum((mine(x) * w.cuda()).sum() for x, w in zip(xg, ws))
    np.testing.assert_allclose(float(loss_g), float(loss_r), rtol=1e-4)
    loss_r.backward()
    loss_g.backward()
    for a, b in zip(xg, xr):
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.numpy(), rtol=2e-4, atol=2e-5)
    for (n, p_r), (_, p_g) in zip(ref.named_parameters(), mine.named_parameters()):
        np.testing.assert_allclose(p_g.grad.cpu().numpy(), p_r.grad.numpy(), rtol=2e-4, atol=3e-5, err_msg=n)
    assert sum(len(v) for v in mine._runners.values()) == 3      # all three runners returned to the pool
    with torch.no_grad():
        mine(xs[0].cuda())
    assert sum(len(v) for v in mine._runners.values()) == 3      # a no-grad call borrows and returns one


# ---------------------------------------------------------------------------------------------------- fastdtw properties
def test_fastdtw_r1_properties_on_random_pairs():
    from oracle import gamma as og
    from subgnn_b200 import ops
    rnd = np.random.RandomState(7)
    nA, nB, max_a, max_b = 80, 90, 24, 50                         # 7 200 pairs
    la = rnd.randint(1, max_a + 1, size=nA).astype(np.int32)
    lb = rnd.randint(1, max_b + 1, size=nB).astype(np.int32)
    la[:12] = rnd.randint(1, 3, size=12)                          # plenty of length-1/2 rows: fastdtw's exact base case
    lb[:12] = rnd.randint(1, 3, size=12)
    A = np.zeros((nA, max_a), dtype=np.int32)
    B = np.zeros((nB, max_b), dtype=np.int32)
    for i in range(nA):
        A[i, :la[i]] = np.sort(rnd.randint(0, 40, size=la[i]))    # ordered degree sequences (gamma.py:47)
    for j in range(nB):
        B[j, :lb[j]] = np.sort(rnd.randint(0, 300, size=lb[j]))
    dev = [torch.from_numpy(x).cuda() for x in (A, la, B, lb)]
    fast = ops.dtw_batch(*dev, ops.DTW_FASTDTW_R1, max_len_a=max_a, max_len_b=max_b).cpu().numpy()
    exact = ops.dtw_batch(*dev, ops.DTW_EXACT, max_len_a=max_a, max_len_b=max_b).cpu().numpy()
    assert fast.shape == (nA, nB) and np.all(fast > 0) and np.all(fast <= 1)
    assert np.all(fast <= exact)                                  # sim = 1/(1+d): d_fast >= d_exact on every pair
    short = (la[:, None] < 3) | (lb[None, :] < 3)
    assert short.sum() > 1500 and np.array_equal(fast[short], exact[short])
    differ = (fast != exact).mean()
    assert 0.0 < differ < 0.2, differ                             # the approximation is real but rare (~1 % of pairs)
    for i, j in zip(rnd.choice(nA, 40), rnd.choice(nB, 40)):      # and both equal the oracle's restatement
        x, y = A[i, :la[i]].tolist(), B[j, :lb[j]].tolist()
        assert fast[i, j] == np.float32(og.calc_dtw(x, y, 'fastdtw_r1')) and exact[i, j] == np.float32(og.calc_dtw(x, y, 'exact'))
    # symmetry of the exact DTW under swapping the roles of the two sequence sets (dist is symmetric)
    swapped = ops.dtw_batch(dev[2], dev[3], dev[0], dev[1], ops.DTW_EXACT, max_len_a=max_b, max_len_b=max_a).cpu().numpy()
    np.testing.assert_allclose(swapped.T, exact, rtol=1e-6)


# ---------------------------------------------------------------------------------------------------- walk laws
def _graph(n=300, m=4, seed=1):
    G = nx.barabasi_albert_graph(n, m, seed=seed)
    edges = [(u + 1, v + 1) for u, v in G.edges()]
    return n, edges


def _tv(a, b):
    a, b = a / a.sum(), b / b.sum()
    return 0.5 * np.abs(a - b).sum()


@pytest.mark.parametrize('mode', ['full', 'inside', 'border'])
def test_walk_laws_match_reference_stream_sampler(mode):
    """patch-size (unique nodes per walk) histogram, per-node visit frequency, walk-length law: GPU Philox sampler vs the
    reference-stream sampler (oracle walk logic, pinned bit-exact to the reference under MT streams in test_oracle_pins.py)."""
    from oracle import walks as ow
    from oracle.rng import MTStream
    from subgnn_b200 import ops
    from subgnn_b200.graph import DeviceGraph
    N, edges = _graph()
    S, g = ow.SortedAdj(N, edges), DeviceGraph.from_edges(N, edges)
    np.random.seed(11)
    random.seed(11)
    beta = 0.65
    if mode == 'full':
        n, L = 4000, 20
        ref = ow.sample_structure_anchor_patches(S, n, L, beta, lambda i: MTStream())
        got = ops.walk_full(g, n, L, beta, 2024).cpu().numpy()
        ref = np.pad(ref, ((0, 0), (0, L - ref.shape[1])))
    else:
        patches = ow.sample_structure_anchor_patches(S, 160, 30, beta, ow.philox_patch_factory(9))
        W, L = 25, 10
        ref = ow.perform_random_walks(S, patches, W, L, beta, mode == 'inside', lambda p_, w_: MTStream()).reshape(-1, L)
        got = ops.walk_patch(g, torch.from_numpy(patches).cuda(), W, L, beta, mode == 'border', 4242).cpu().numpy().reshape(-1, L)
    assert ref.shape == got.shape
    # walk length law
    len_r, len_g = (ref > 0).sum(1), (got > 0).sum(1)
    assert _tv(np.bincount(len_r, minlength=L + 1).astype(float), np.bincount(len_g, minlength=L + 1).astype(float)) < 0.04
    # patch-size law: number of distinct nodes per walk
    uniq = lambda a: np.array([len(set(r[r > 0].tolist())) for r in a])
    u_r, u_g = uniq(ref), uniq(got)
    assert abs(u_r.mean() - u_g.mean()) < 0.25, (u_r.mean(), u_g.mean())
    assert _tv(np.bincount(u_r, minlength=L + 1).astype(float), np.bincount(u_g, minlength=L + 1).astype(float)) < 0.06
    # per-node visit frequency (nodes pooled into 12 quantile groups of the reference's frequency so every cell is populated)
    v_r = np.bincount(ref[ref > 0], minlength=N + 1).astype(float)
    v_g = np.bincount(got[got > 0], minlength=N + 1).astype(float)
    order = np.argsort(-v_r, kind='stable')
    groups = np.array_split(order, 12)
    c_r = np.array([v_r[gp].sum() for gp in groups])
    c_g = np.array([v_g[gp].sum() for gp in groups])
    assert _tv(c_r, c_g) < 0.05, (c_r / c_r.sum(), c_g / c_g.sum())
    top = order[:10]                                               # the hubs individually
    assert np.abs(v_r[top] / v_r.sum() - v_g[top] / v_g.sum()).max() < 0.012
    # P(triangular move | both classes non-empty) = beta, measured on the GPU walks
    if mode == 'full':
        tri = non = 0
        for row in got[:1500]:
            row = row[row > 0]
            for t in range(2, len(row)):
                prev, cur, nxt = int(row[t - 2]), int(row[t - 1]), int(row[t])
                nb = S.neighbors(cur)
                has_tri = any(S.has_edge(prev, x) for x in nb)
                has_non = any(not S.has_edge(prev, x) for x in nb)
                if has_tri and has_non:
                    if S.has_edge(prev, nxt):
                        tri += 1
                    else:
                        non += 1
        frac = tri / (tri + non)
        assert abs(frac - beta) < _binom_tol(beta, tri + non), frac


# ---------------------------------------------------------------------------------------------------- init_anchors_* containers
def test_init_anchor_containers_follow_the_reference_contract():
    """a7: init_anchors_neighborhood / pos_int / pos_ext / structure (anchor_patch_samplers.py:248-328): container layout,
    shapes, every drawn id a member of the set it must come from, the F10 PAD law, `indices` consistent with the sliced tensors,
    and the same result for the same hparams['seed'] from a fresh call-counter state (no per-process hash salt)."""
    from oracle import sampling as osamp
    from oracle import walks as ow
    from subgnn_b200 import anchor_patch_samplers as aps
    from subgnn_b200.graph import DeviceGraph
    N, edges = _graph(200, 3, seed=5)
    S, g = ow.SortedAdj(N, edges), DeviceGraph.from_edges(N, edges)
    rnd = random.Random(3)
    subs = {k: [sorted(set(rnd.randrange(1, N + 1) for _ in range(rnd.randint(2, 9)))) for _ in range(n)] for k, n in (('train', 14), ('val', 5), ('test', 4))}
    cc = {k: torch.from_numpy(osamp.initialize_cc_ids(S, v)) for k, v in subs.items()}
    bor = {k: torch.from_numpy(osamp.initialize_border_sets(S, cc[k].numpy(), 1)) for k in subs}
    hp = {'seed': 13, 'n_layers': 2, 'n_anchor_patches_N_in': 6, 'n_anchor_patches_N_out': 9, 'n_anchor_patches_pos_in': 5,
          'n_anchor_patches_pos_out': 11, 'n_anchor_patches_structure': 4, 'n_triangular_walks': 3, 'random_walk_len': 5,
          'sample_walk_len': 8, 'rw_beta': 0.65, 'structure_patch_type': 'triangular_random_walk'}

    def run():
        aps.reset_call_counters()
        ai, ab = aps.init_anchors_neighborhood('all', hp, g, 'cuda', cc['train'], cc['val'], cc['test'], bor['train'], bor['val'], bor['test'])
        pi = aps.init_anchors_pos_int('train_val', hp, g, 'cuda', subs['train'], subs['val'], subs['test'])
        pe = aps.init_anchors_pos_ext(hp, g, 'cuda')
        patches = aps.sample_structure_anchor_patches(hp, g, 'cuda', 3)
        irw = aps.perform_random_walks(hp, g, patches, inside=True)
        brw = aps.perform_random_walks(hp, g, patches, inside=False)
        st = aps.init_anchors_structure(hp, patches, irw, brw)
        return ai, ab, pi, pe, patches, irw, brw, st

    ai, ab, pi, pe, patches, irw, brw, st = run()
    assert set(ai) == set(ab) == {'train', 'val', 'test'} and set(pi) == {'train', 'val'} and set(pe) == set(st) == {0, 1}
    for k in subs:
        for l in range(2):
            a_in, a_out = ai[k][l], ab[k][l]
            assert a_in.dtype == torch.int64 and not a_in.is_cuda
            assert a_in.shape == (len(subs[k]), cc[k].shape[1], 6) and a_out.shape == (len(subs[k]), cc[k].shape[1], 9)
            for s in range(len(subs[k])):
                for c in range(cc[k].shape[1]):
                    comp = set(cc[k][s, c].tolist()) - {0}
                    bset = set(bor[k][s, c].tolist()) - {0}
                    assert set(a_in[s, c].tolist()) - {0} <= comp and set(a_out[s, c].tolist()) - {0} <= bset
                    if not comp:
                        assert int(a_in[s, c].abs().sum()) == 0 and int(a_out[s, c].abs().sum()) == 0
                    elif len(comp) == cc[k].shape[2]:
                        assert 0 not in a_in[s, c].tolist()          # a full-width row is never PAD (F10)
    assert not torch.equal(ai['train'][0], ai['train'][1])           # a fresh draw per layer
    for k in ('train', 'val'):
        for l in range(2):
            assert pi[k][l].shape == (len(subs[k]), 5)
            for s, nodes in enumerate(subs[k]):
                assert set(pi[k][l][s].tolist()) <= set(nodes)
    for l in range(2):
        assert pe[l].shape == (11,) and int(pe[l].min()) >= 1 and int(pe[l].max()) <= N
        pa, indices, wi, wb = st[l]
        assert isinstance(indices, list) and len(indices) == 4 and all(0 <= i < patches.shape[0] for i in indices)
        assert (indices * 3) == [*indices, *indices, *indices]       # subgraph_mpn.py:88 relies on list * int tiling
        assert torch.equal(pa, patches[indices]) and torch.equal(wi, irw[indices]) and torch.equal(wb, brw[indices])
        assert wi.shape == (4, 3, 5)
    # walks stay inside / on the border of their patch
    for pch, walks_in in zip(patches.tolist(), irw.tolist()):
        members = set(pch) - {0}
        assert all(set(w) - {0} <= members for w in walks_in)
    again = run()
    for a, b in zip((ai, ab, pi, pe), again[:4]):
        for k in a:
            if isinstance(a[k], dict):
                for l in a[k]:
                    assert torch.equal(a[k][l], b[k][l])
            else:
                assert torch.equal(a[k], b[k])
    assert torch.equal(patches, again[4]) and torch.equal(irw, again[5]) and torch.equal(brw, again[6])
    assert all(st[l][1] == again[7][l][1] for l in st)
    assert aps._seed({'seed': 13}, 'trw') != aps._seed({'seed': 14}, 'trw')
