"""GPU parity of the setup-phase kernels (walks, samplers, border sets, SP-min, degree sequences, DTW)
against the CPU oracle — bit-exact for ids/integers, exact fp32 equality for similarities."""
import random

import networkx as nx
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def build_graph(n=300, m=4, seed=1):
    G = nx.barabasi_albert_graph(n, m, seed=seed)
    G.add_edge(0, n)          # a pendant node
    G.add_node(n + 1)         # an isolated node
    N = n + 2
    edges = [(u + 1, v + 1) for u, v in G.edges()]
    return N, edges


@pytest.fixture(scope='module')
def env():
    from oracle import walks as ow
    from subgnn_b200.graph import DeviceGraph
    N, edges = build_graph()
    return N, edges, ow.SortedAdj(N, edges), DeviceGraph.from_edges(N, edges)


def test_walk_full_bit_exact(env):
    from oracle import walks as ow
    from subgnn_b200 import ops
    N, edges, S, g = env
    for seed, L in ((5, 50), (77, 12)):
        got = ops.walk_full(g, 64, L, 0.65, seed).cpu().numpy()
        ref = ow.sample_structure_anchor_patches(S, 64, L, 0.65, ow.philox_patch_factory(seed))
        want = np.zeros_like(got)
        want[:, :ref.shape[1]] = ref
        assert np.array_equal(got, want)


@pytest.mark.parametrize('border', [False, True])
def test_walk_patch_bit_exact(env, border):
    from oracle import walks as ow
    from subgnn_b200 import ops
    N, edges, S, g = env
    patches = ow.sample_structure_anchor_patches(S, 40, 30, 0.65, ow.philox_patch_factory(3))
    patches[7, :] = 0                                      # an all-PAD patch
    patches[8, 1:] = 0
    patches[8, 0] = N                                      # the isolated node alone
    got = ops.walk_patch(g, torch.from_numpy(patches).cuda(), 5, 10, 0.65, border, 11).cpu().numpy()
    ref = ow.perform_random_walks(S, patches, 5, 10, 0.65, not border, ow.philox_walk_factory(11, 5))
    assert np.array_equal(got, ref)


def test_walk_distribution_matches_reference_sampler(env):
    """Distributional parity with the reference-stream sampler: visited-degree histogram and walk lengths."""
    from oracle import walks as ow
    from oracle.rng import MTStream
    from subgnn_b200 import ops
    N, edges, S, g = env
    np.random.seed(0)
    random.seed(0)
    ref = ow.sample_structure_anchor_patches(S, 1500, 20, 0.65, lambda i: MTStream())
    got = ops.walk_full(g, 1500, 20, 0.65, 123).cpu().numpy()
    deg = np.array([0] + [len(S.adj[i]) for i in range(1, N + 1)])
    bins = [0, 1, 4, 6, 9, 15, 30, 1000]
    h_ref, _ = np.histogram(deg[ref[ref > 0]], bins=bins)
    h_got, _ = np.histogram(deg[got[got > 0]], bins=bins)
    p_ref, p_got = h_ref / h_ref.sum(), h_got / h_got.sum()
    assert np.abs(p_ref - p_got).max() < 0.02
    assert abs((ref > 0).sum(1).mean() - (got > 0).sum(1).mean()) < 0.2


def _cc(N, n_sub=40, seed=2):
    rnd = random.Random(seed)
    return [sorted(set(rnd.randrange(1, N + 1) for _ in range(rnd.randint(1, 12)))) for _ in range(n_sub)]


def test_border_sets_and_sampling(env):
    from oracle import sampling as osamp
    from subgnn_b200 import ops
    from subgnn_b200.graph import ragged_from_padded
    N, edges, S, g = env
    cc_ids = osamp.initialize_cc_ids(S, _cc(N))
    flat = cc_ids.reshape(-1, cc_ids.shape[-1])
    ptr, items = ragged_from_padded(flat)
    dptr, ditems = torch.from_numpy(ptr).cuda(), torch.from_numpy(items).cuda()
    for k in (1, 2):
        want = osamp.initialize_border_sets(S, cc_ids, k).reshape(flat.shape[0], -1)
        optr, out = ops.border_khop(g, dptr, ditems, k)
        optr, out = optr.cpu().numpy(), out.cpu().numpy()
        for r in range(flat.shape[0]):
            assert np.array_equal(out[optr[r]:optr[r + 1]], want[r][want[r] != 0])
    # neighbourhood sampling (same Philox law as the oracle) and position sampling
    got = ops.sample_rows(dptr, ditems, flat.shape[1], 7, True, 99, 3, False).cpu().numpy()
    want = osamp.sample_neighborhood_anchor_patch_philox(cc_ids, 7, 99, 3).reshape(flat.shape[0], 7)
    assert np.array_equal(got, want)
    pools = _cc(N, 9, seed=8)
    pptr = np.concatenate([[0], np.cumsum([len(p) for p in pools])]).astype(np.int32)
    pit = np.concatenate(pools).astype(np.int32)
    got = ops.sample_rows(torch.from_numpy(pptr).cuda(), torch.from_numpy(pit).cuda(), 0, 11, False, 5, 1, True).cpu().numpy()
    assert np.array_equal(got, osamp.sample_position_philox(None, pools, 11, 5, 1))


def test_neighborhood_sampling_law(env):
    """chi-square style check of the F10 law against the reference's randn/argmax construction."""
    from oracle import sampling as osamp
    from subgnn_b200 import ops
    rows = np.zeros((1, 3, 4), dtype=np.int64)
    rows[0, 0, :2] = [5, 9]          # k=2 of width 4 -> P(PAD)=1/4
    rows[0, 1, :4] = [1, 2, 3, 4]    # full row -> never PAD
    # row 2 all PAD
    A = 20000
    torch.manual_seed(0)
    ref = osamp.sample_neighborhood_anchor_patch_torch(torch.from_numpy(rows), A).numpy()[0]
    from subgnn_b200.graph import ragged_from_padded
    ptr, items = ragged_from_padded(rows.reshape(3, 4))
    got = ops.sample_rows(torch.from_numpy(ptr).cuda(), torch.from_numpy(items).cuda(), 4, A, True, 1, 0, False).cpu().numpy()
    for r in range(3):
        for v in np.unique(np.concatenate([ref[r], got[r]])):
            assert abs((ref[r] == v).mean() - (got[r] == v).mean()) < 0.015
    assert abs((got[0] == 0).mean() - 0.25) < 0.01 and (got[1] == 0).sum() == 0 and (got[2] != 0).sum() == 0


def test_sp_min(env):
    from oracle import sampling as osamp
    from subgnn_b200 import ops
    from subgnn_b200.graph import ragged_from_padded
    N, edges, S, g = env
    hop = osamp.all_pairs_hops(S)
    g.set_hop_table(hop)
    cc_ids = osamp.initialize_cc_ids(S, _cc(N, 25, seed=4))
    flat = cc_ids.reshape(-1, cc_ids.shape[-1])
    ptr, items = ragged_from_padded(flat)
    dptr, ditems = torch.from_numpy(ptr).cuda(), torch.from_numpy(items).cuda()
    want = osamp.shortest_path_similarities(hop, cc_ids).reshape(flat.shape[0], -1)
    got = ops.sp_min_dense(g.hop, dptr, ditems).cpu().numpy()
    assert np.array_equal(got, want)
    rnd = np.random.RandomState(0)
    anchors = rnd.randint(0, N + 1, size=(flat.shape[0], 9)).astype(np.int32)     # includes PAD anchors
    got = ops.sp_min_gather(g.hop, dptr, ditems, torch.from_numpy(anchors).cuda()).cpu().numpy()
    ref = np.where(anchors > 0, np.take_along_axis(want, np.maximum(anchors - 1, 0).astype(np.int64), axis=1), 0)
    assert np.array_equal(got, ref)


def test_degree_sequences_and_dtw(env):
    from oracle import gamma as og
    from subgnn_b200 import ops
    N, edges, S, g = env
    rnd = np.random.RandomState(1)
    rows = np.zeros((60, 40), dtype=np.int64)
    for r in range(59):
        L = rnd.randint(1, 40)
        rows[r, :L] = rnd.randint(1, N + 1, size=L)      # duplicates on purpose
        if r % 7 == 0:
            rows[r, L // 2] = 0                           # PAD in the middle
    seqs = {}
    for internal in (True, False):
        seq, ln = ops.degree_seq(g, torch.from_numpy(rows).cuda(), internal)
        seq, ln = seq.cpu().numpy(), ln.cpu().numpy()
        for r in range(rows.shape[0]):
            want = og.get_degree_sequence(S, rows[r], internal)
            assert ln[r] == len(want) and seq[r, :ln[r]].tolist() == want
        seqs[internal] = (seq, ln)
    # DTW all pairs: internal sequences of the first 20 rows (as components) x border sequences (as patches)
    sa, la = seqs[True]
    sb, lb = seqs[False]
    A = torch.from_numpy(sa[:20]).cuda()
    LA = torch.from_numpy(la[:20]).cuda()
    B = torch.from_numpy(sb).cuda()
    LB = torch.from_numpy(lb).cuda()
    for mode, name in ((ops.DTW_FASTDTW_R1, 'fastdtw_r1'), (ops.DTW_EXACT, 'exact'), (ops.DTW_EXACT_THREAD, 'exact')):
        got = ops.dtw_batch(A, LA, B, LB, mode).cpu().numpy()
        for i in range(20):
            for j in range(rows.shape[0]):
                want = np.float32(og.calc_dtw(sa[i, :la[i]].tolist(), sb[j, :lb[j]].tolist(), name))
                assert got[i, j] == want, (i, j, mode)


@pytest.mark.parametrize('max_a', [1, 3, 4, 7, 13, 32, 33, 100, 150, 256, 300])
def test_dtw_exact_wavefront_equals_thread_mapping(max_a):
    """SUBGNN_DTW_EXACT (sub-warp wavefront, DP carried in registers through shuffles; every lanes-per-pair / rows-per-lane
    instantiation) against SUBGNN_DTW_EXACT_THREAD (row-by-row, one thread per pair) and the oracle: bit-identical fp32
    similarities, empty sequences and ragged lengths included."""
    from oracle import gamma as og
    from subgnn_b200 import ops
    rnd = np.random.RandomState(100 + max_a)
    nA, nB, max_b = 37, 53, 50
    la = rnd.randint(0, max_a + 1, size=nA).astype(np.int32)
    la[0], la[-1] = max_a, 0
    lb = rnd.randint(0, max_b + 1, size=nB).astype(np.int32)
    lb[0], lb[1], lb[2] = max_b, 0, 1
    A = np.zeros((nA, max_a), dtype=np.int32)
    B = np.zeros((nB, max_b), dtype=np.int32)
    for i in range(nA):
        A[i, :la[i]] = np.sort(rnd.randint(0, 60, size=la[i]))
    for j in range(nB):
        B[j, :lb[j]] = np.sort(rnd.randint(0, 400, size=lb[j]))
    dev = [torch.from_numpy(x).cuda() for x in (A, la, B, lb)]
    wave = ops.dtw_batch(*dev, ops.DTW_EXACT, max_len_a=max_a, max_len_b=max_b).cpu().numpy()          # rows bucketed by length
    wave_one = ops.dtw_batch(*dev, ops.DTW_EXACT, max_len_a=max_a, max_len_b=max_b, bucketed=False).cpu().numpy()   # one launch
    thread = ops.dtw_batch(*dev, ops.DTW_EXACT_THREAD, max_len_a=max_a, max_len_b=max_b).cpu().numpy()
    assert np.array_equal(wave, thread) and np.array_equal(wave_one, thread)
    assert np.all(wave[la == 0] == 0) and np.all(wave[:, lb == 0] == 0)
    for i in rnd.choice(nA, 6, replace=False):
        for j in rnd.choice(nB, 6, replace=False):
            want = np.float32(og.calc_dtw(A[i, :la[i]].tolist(), B[j, :lb[j]].tolist(), 'exact'))
            assert wave[i, j] == want, (i, j)


def test_dtw_golden_reference_pairs():
    """The committed reference-generated DTW similarities (tests/golden/gamma_golden.npz)."""
    from subgnn_b200 import ops
    from tests.util import load_npz
    g = load_npz('gamma_golden.npz')
    oi = np.concatenate([[0], np.cumsum(g['seq_int_len'])])
    ob = np.concatenate([[0], np.cumsum(g['seq_bor_len'])])
    n = g['dtw_sims'].shape[0]
    A = np.zeros((n, 16), dtype=np.int32)
    B = np.zeros((n, 16), dtype=np.int32)
    for i in range(n):
        A[i, :g['seq_int_len'][i]] = g['seq_int'][oi[i]:oi[i + 1]]
        B[i, :g['seq_bor_len'][i]] = g['seq_bor'][ob[i]:ob[i + 1]]
    got = ops.dtw_batch(torch.from_numpy(A).cuda(), torch.from_numpy(g['seq_int_len'][:n].astype(np.int32)).cuda(),
                        torch.from_numpy(B).cuda(), torch.from_numpy(g['seq_bor_len'][:n].astype(np.int32)).cuda()).cpu().numpy()
    assert np.array_equal(got, g['dtw_sims'].astype(np.float32))
