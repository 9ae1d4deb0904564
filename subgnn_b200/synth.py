"""Synthetic workloads of the named benchmark shapes (BASELINE.json configs).

The reference's generator (prepare_dataset/prepare_dataset.py) cannot run on Python >= 3.11 (random.sample on sets,
SURVEY F14) and the real datasets are not available offline, so the shapes are restated here: Barabasi-Albert base
graphs of the stated node / edge counts (seed 42, config_prepare_dataset.py:15), subgraphs with the stated size and
component statistics, labels uniform over the class count, N(0,1) node embeddings, hyper-parameters from the
reference's best_model_hyperparameters/*.json (values inlined below with their source file).
"""
import numpy as np

COMMON = {"structure_patch_type": "triangular_random_walk", "lstm_aggregator": "last", "resample_anchor_patches": False,
          "freeze_node_embeds": False, "use_mpn_projection": True, "compute_similarities": False, "rw_beta": 0.65,
          "max_sim_epochs": 5, "trainable_cc": False, "use_neighborhood": True, "use_position": True, "use_structure": True,
          "linear_hidden_dim_1": 64, "linear_hidden_dim_2": 64}

WORKLOADS = {
    # best_model_hyperparameters/density/all_density_hyperparams.json
    'density': dict(graph=('ba', 5000, 5), n_sub=250, n_classes=3,
                    hp=dict(node_embed_size=32, batch_size=64, n_layers=1, n_anchor_patches_N_in=10, n_anchor_patches_N_out=43,
                            neigh_sample_border_size=2, n_anchor_patches_pos_in=57, n_anchor_patches_pos_out=183,
                            n_anchor_patches_structure=42, n_triangular_walks=5, random_walk_len=10, sample_walk_len=50,
                            lstm_n_layers=1, cc_aggregator='max', learning_rate=0.0002951850045886519, grad_clip=0.1929946246623414,
                            lin_dropout=0.2522849803237359, lstm_dropout=0.34069986114743744)),
    # best_model_hyperparameters/cutratio/S_cutratio_hyperparams.json (structure channel only)
    'cutratio': dict(graph=('ba', 5000, 5), n_sub=250, n_classes=3,
                     hp=dict(use_neighborhood=False, use_position=False, node_embed_size=64, batch_size=128, n_layers=4,
                             n_anchor_patches_N_in=0, n_anchor_patches_N_out=0, neigh_sample_border_size=1, n_anchor_patches_pos_in=0,
                             n_anchor_patches_pos_out=0, n_anchor_patches_structure=28, n_triangular_walks=5, random_walk_len=10,
                             sample_walk_len=50, lstm_n_layers=2, cc_aggregator='sum', learning_rate=1e-3, grad_clip=0.2,
                             lin_dropout=0.2, lstm_dropout=0.2)),
    # best_model_hyperparameters/ppi_bp/hyperparams.json with all three channels switched on (BASELINE.json config 3)
    'ppi_bp': dict(graph=('ba', 17080, 19), n_sub=1591, n_classes=6,
                   hp=dict(node_embed_size=64, batch_size=32, n_layers=4, n_anchor_patches_N_in=12, n_anchor_patches_N_out=58,
                           neigh_sample_border_size=1, n_anchor_patches_pos_in=66, n_anchor_patches_pos_out=178,
                           n_anchor_patches_structure=25, n_triangular_walks=5, random_walk_len=10, sample_walk_len=50,
                           lstm_n_layers=2, cc_aggregator='sum', learning_rate=0.00015124982405268227, grad_clip=0.20580294622501977,
                           lin_dropout=0.36813148347966873, lstm_dropout=0.2353272396179573)),
    # best_model_hyperparameters/hpo_metab/hyperparams.json, all channels (BASELINE.json config 4)
    'hpo_metab': dict(graph=('ba', 14587, 225), n_sub=2400, n_classes=6,
                      hp=dict(node_embed_size=128, batch_size=64, n_layers=4, n_anchor_patches_N_in=13, n_anchor_patches_N_out=34,
                              neigh_sample_border_size=2, n_anchor_patches_pos_in=56, n_anchor_patches_pos_out=90,
                              n_anchor_patches_structure=18, n_triangular_walks=5, random_walk_len=10, sample_walk_len=50,
                              lstm_n_layers=2, cc_aggregator='sum', learning_rate=1e-3, grad_clip=0.25, lin_dropout=0.3, lstm_dropout=0.2)),
    # best_model_hyperparameters/em_user/hyperparams.json, all channels (BASELINE.json config 5)
    'em_user': dict(graph=('ba', 57333, 80), n_sub=324, n_classes=2,
                    hp=dict(node_embed_size=128, batch_size=32, n_layers=1, n_anchor_patches_N_in=16, n_anchor_patches_N_out=32,
                            neigh_sample_border_size=2, n_anchor_patches_pos_in=48, n_anchor_patches_pos_out=77,
                            n_anchor_patches_structure=35, n_triangular_walks=10, random_walk_len=23, sample_walk_len=22,
                            lstm_n_layers=1, cc_aggregator='sum', learning_rate=1e-3, grad_clip=0.25, lin_dropout=0.3, lstm_dropout=0.2)),
    # tiny shape for smoke tests
    'tiny': dict(graph=('ba', 400, 4), n_sub=60, n_classes=3,
                 hp=dict(node_embed_size=16, batch_size=8, n_layers=2, n_anchor_patches_N_in=4, n_anchor_patches_N_out=6,
                         neigh_sample_border_size=1, n_anchor_patches_pos_in=5, n_anchor_patches_pos_out=9,
                         n_anchor_patches_structure=6, n_triangular_walks=3, random_walk_len=6, sample_walk_len=12, max_sim_epochs=2,
                         lstm_n_layers=2, cc_aggregator='sum', learning_rate=1e-3, grad_clip=0.25, lin_dropout=0.0, lstm_dropout=0.0,
                         linear_hidden_dim_1=16, linear_hidden_dim_2=12)),
}


def hparams(name):
    hp = dict(COMMON)
    hp.update(WORKLOADS[name]['hp'])
    return hp


def ba_edges(n, m, seed=42):
    """Barabasi-Albert edge list (0-indexed) — networkx's generator, kept as the SURVEY's concrete graphs use it."""
    import networkx as nx
    G = nx.barabasi_albert_graph(n, m, seed=seed)
    return np.array(G.edges(), dtype=np.int64)


def _neighbors(rp, col, u):
    return col[rp[u]:rp[u + 1]]


def make_subgraphs(name, n_nodes, rp, col, n_sub, rs):
    """Node-id lists (1-indexed) with the component statistics of SURVEY §8d."""
    subs = []
    for _ in range(n_sub):
        nodes = set()
        if name in ('density', 'cutratio', 'tiny'):
            # 20-node BFS subgraph; DENSITY: keep each node w.p. so that edge removal fragments it (F16: ~3 components)
            size = 20 if name != 'tiny' else int(rs.randint(4, 12))
            src = int(rs.randint(n_nodes))
            order, seen, qi = [src], {src}, 0
            while len(order) < size and qi < len(order):
                for v in _neighbors(rp, col, order[qi]):
                    v = int(v)
                    if v not in seen:
                        seen.add(v)
                        order.append(v)
                        if len(order) >= size:
                            break
                qi += 1
            nodes = set(order)
            if name in ('density', 'tiny'):
                extra = rs.randint(n_nodes, size=int(rs.randint(0, 4)))          # a few detached nodes -> extra components
                nodes.update(int(x) for x in extra)
        elif name == 'ppi_bp':
            for _c in range(int(rs.poisson(6)) + 1):
                src = int(rs.randint(n_nodes))
                nodes.add(src)
                if rs.rand() >= 0.8:
                    nb = _neighbors(rp, col, src)
                    k = min(len(nb), int(rs.randint(1, 6)))
                    nodes.update(int(x) for x in rs.choice(nb, size=k, replace=False))
        elif name == 'hpo_metab':
            size = max(3, int(round(rs.normal(14.4, 6.2))))
            n_cc = int(rs.choice([1, 2, 3], p=[0.55, 0.36, 0.09]))
            per = max(1, size // n_cc)
            deg = rp[1:] - rp[:-1]
            for _c in range(n_cc):
                src = int(rs.randint(n_nodes))
                nodes.add(src)
                nb = _neighbors(rp, col, src)
                low = nb[np.argsort(deg[nb], kind='stable')[:per - 1]]           # seed + its lowest-degree neighbours
                nodes.update(int(x) for x in low)
        elif name == 'em_user':
            for _c in range(max(1, int(round(rs.normal(52, 15))))):
                src = int(rs.randint(n_nodes))
                nodes.add(src)
                k = int(rs.geometric(0.5)) - 1                                   # component size 1 + Geometric(mean 2) - 1
                nb = _neighbors(rp, col, src)
                if k > 0:
                    nodes.update(int(x) for x in nb[np.argsort(rp[nb + 1] - rp[nb], kind='stable')[:k]])
        else:
            raise KeyError(name)
        subs.append(sorted(n + 1 for n in nodes))
    return subs


def make_workload(name, seed=42, device='cuda', scale_subgraphs=1, graph=None, n_sub=None):
    """-> (hparams, DeviceGraph, subgraphs{split}, labels{split}, embeddings).
    graph=('ba', n, m) / n_sub override the base graph and the subgraph count (reduced-size instances of a named shape for the
    parity tests: every hyper-parameter, hence every kernel instantiation, stays the benchmark's)."""
    from .graph import DeviceGraph
    w = dict(WORKLOADS[name])
    if graph is not None:
        w['graph'] = graph
    if n_sub is not None:
        w['n_sub'] = n_sub
    hp = hparams(name)
    _, n, m = w['graph']
    edges = ba_edges(n, m, seed)
    g = DeviceGraph.from_edges(n, edges, device=device, one_indexed=False)
    rs = np.random.RandomState(seed)
    n_sub = w['n_sub'] * scale_subgraphs
    subs = make_subgraphs(name, n, g.rowptr_host.astype(np.int64), g.col_host.astype(np.int64), n_sub, rs)
    labels = rs.randint(w['n_classes'], size=n_sub)
    labels[:w['n_classes']] = np.arange(w['n_classes'])
    n_train, n_val = int(0.8 * n_sub), int(0.1 * n_sub)
    split = {'train': slice(0, n_train), 'val': slice(n_train, n_train + n_val), 'test': slice(n_train + n_val, n_sub)}
    subgraphs = {k: subs[v] for k, v in split.items()}
    labs = {k: labels[v] for k, v in split.items()}
    D = hp['node_embed_size']
    emb = np.zeros((n + 1, D), dtype=np.float32)
    emb[1:] = rs.standard_normal((n, D)).astype(np.float32)
    return hp, g, subgraphs, labs, emb
