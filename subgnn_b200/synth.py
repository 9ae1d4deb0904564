"""Synthetic workloads of the named benchmark shapes (BASELINE.json configs).

The reference's generator (prepare_dataset/prepare_dataset.py) cannot run on Python >= 3.11 (random.sample on sets,
SURVEY F14) and the real datasets are not available offline.  DENSITY and CUTRATIO (BASELINE configs[0], [1]) are the
reference generator's recipe restated step by step (generate_property_dataset: BFS / planted subgraphs, edge editing toward
the drawn density / cut-ratio target, largest-component relabelling, equal-count label bins); the three real-data shapes
are Barabasi-Albert base graphs of the stated node / edge counts (seed 42, config_prepare_dataset.py:15) with subgraphs
of the stated size and component statistics and uniform labels.  N(0,1) node embeddings; hyper-parameters from the
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


# ---- DENSITY / CUTRATIO: restatement of prepare_dataset/prepare_dataset.py ------------------------------------------
# The reference's SyntheticGraph cannot run on python >= 3.11 (random.sample over sets / node views); the recipe is
# restated on plain adjacency sets: same steps, constants and order of operations; draws come from random.Random(seed) over SORTED
# candidate lists where the reference samples python sets (stream identity is unattainable either way).
DENSITY_RANGE, DENSITY_EPSILON = [0.05, 0.25, 0.45], 0.01              # config_prepare_dataset.py:38-39
CUT_RATIO_RANGE, CUT_RATIO_EPSILON = [0.005, 0.0125, 0.02], 0.001      # :40-41
MAX_TRIES = 100                                                        # :46


def _bfs_nodes(adj, start, depth_limit, limit):
    """prepare_dataset.py:311-313: [start] + BFS tree edge targets (depth <= depth_limit), truncated to `limit` nodes; neighbours
    are visited in adjacency (insertion) order like nx.bfs_edges."""
    order, seen, frontier = [start], {start}, [start]
    for _ in range(depth_limit):
        nxt = []
        for u in frontier:
            for v in adj[u]:
                if v not in seen:
                    seen.add(v)
                    order.append(v)
                    nxt.append(v)
        frontier = nxt
    return order[:limit]


def generate_property_dataset(kind, n=5000, m=5, n_subgraphs=250, n_subgraph_nodes=20, seed=42):
    """DENSITY (README.md:59-74) or CUTRATIO (:76-91) synthetic data set of the reference's generator:
    barabasi_albert_graph(n, m, seed) (prepare_dataset.py:49-52), subgraphs by BFS (:288-327; density) or by planting a complete
    graph on randomly chosen nodes (:469-516; cut ratio), per-subgraph edge editing toward a randomly drawn target value
    (:552-617), restriction to the largest connected component with consecutive relabelling (:619-633), labels by binning the
    ACHIEVED property values into len(range) equal-count bins (:641-693, generate_bins :712-728), 80/10/10 split (:756-779).
    -> (edges int64 (E, 2) 0-indexed, subgraphs {split: [node lists, 1-indexed]}, labels {split: int array}, values (n_kept,))"""
    import random as _random
    import networkx as nx
    assert kind in ('density', 'cutratio')
    rnd = _random.Random(seed)
    G = nx.barabasi_albert_graph(n, m, seed=seed)
    adj = {u: dict.fromkeys(G.neighbors(u)) for u in G.nodes()}        # insertion-ordered neighbour sets

    def add_edge(u, v):
        if u != v:
            adj[u][v] = None
            adj[v][u] = None

    def remove_edge(u, v):
        adj[u].pop(v, None)
        adj[v].pop(u, None)

    nodes_all = sorted(adj)
    subgraphs = []
    if kind == 'density':
        for _ in range(n_subgraphs):                                   # _get_subgraphs_by_bfs, 1 connected component, max_depth 3
            start = rnd.choice(nodes_all)
            subgraphs.append(_bfs_nodes(adj, start, 3, n_subgraph_nodes))
    else:
        for _ in range(n_subgraphs):                                   # _get_subgraphs_by_planting: K_n composed onto sampled nodes
            ids = rnd.sample(nodes_all, n_subgraph_nodes)
            for i, u in enumerate(ids):
                for v in ids[i + 1:]:
                    add_edge(u, v)
            subgraphs.append(ids)
    n_nodes = len(adj)
    for s in subgraphs:                                                # _modify_graph_for_desired_subgraph_properties
        sset, slist = set(s), sorted(set(s))
        if kind == 'density':
            target = rnd.choice(DENSITY_RANGE)
            for _try in range(MAX_TRIES):
                edges = sorted((u, v) for u in slist for v in adj[u] if v in sset and u < v)
                k = len(slist)
                dens = 2.0 * len(edges) / (k * (k - 1)) if k > 1 else 0.0
                if abs(dens - target) < DENSITY_EPSILON:
                    break
                if dens > target:
                    remove_edge(*rnd.choice(edges))                    # :571-573
                else:
                    add_edge(*rnd.sample(slist, 2))                    # :575-577
        else:
            target = rnd.choice(CUT_RATIO_RANGE)
            outside = [u for u in nodes_all if u not in sset]
            for _try in range(MAX_TRIES):
                boundary = sorted((u, v) for u in slist for v in adj[u] if v not in sset)
                ratio = len(boundary) / (len(slist) * (n_nodes - len(slist)))
                if abs(ratio - target) < CUT_RATIO_EPSILON:
                    break
                if ratio > target:
                    remove_edge(*rnd.choice(boundary))                 # :604-606
                else:
                    add_edge(rnd.choice(slist), rnd.choice(outside))   # :608-611
    # _relabel_nodes: largest connected component, consecutive ids in node order; subgraphs lose the removed nodes
    seen, best = set(), []
    for src in nodes_all:
        if src in seen:
            continue
        comp, stack = [], [src]
        seen.add(src)
        while stack:
            u = stack.pop()
            comp.append(u)
            for v in adj[u]:
                if v not in seen:
                    seen.add(v)
                    stack.append(v)
        if len(comp) > len(best):
            best = comp
    keep = set(best)
    mapping = {u: i for i, u in enumerate(u for u in nodes_all if u in keep)}
    edges = np.array(sorted((mapping[u], mapping[v]) for u in keep for v in adj[u] if v in keep and u < v), dtype=np.int64)
    subs = [[mapping[u] for u in s if u in keep] for s in subgraphs]
    nbr = {}
    for u, v in edges:
        nbr.setdefault(int(u), set()).add(int(v))
        nbr.setdefault(int(v), set()).add(int(u))
    values = []
    for s in subs:                                                     # generate_subgraph_labels: the ACHIEVED values
        sset = set(s)
        if kind == 'density':
            e = sum(1 for u in sset for v in nbr.get(u, ()) if v in sset and u < v)
            k = len(sset)
            values.append(2.0 * e / (k * (k - 1)) if k > 1 else 0.0)
        else:
            b = sum(1 for u in sset for v in nbr.get(u, ()) if v not in sset)
            values.append(b / (len(sset) * (len(mapping) - len(sset))))
    values = np.asarray(values)
    n_bins = len(DENSITY_RANGE if kind == 'density' else CUT_RATIO_RANGE)
    srt = np.sort(values)
    cuts = (len(srt) / float(n_bins)) * np.arange(1, n_bins + 1)       # generate_bins :724-726
    bins = np.unique(np.array([srt[int(b) - 1] for b in cuts]))
    bins = np.delete(bins, len(bins) - 1)
    labels = np.digitize(values, bins=bins)
    idx = list(range(len(subs)))                                       # generate_mask :756-779
    train = set(rnd.sample(idx, int(len(idx) * 0.8)))
    rest = [i for i in idx if i not in train]
    val = set(rnd.sample(rest, len(rest) // 2))
    split_of = ['train' if i in train else ('val' if i in val else 'test') for i in idx]
    out_s = {k: [] for k in ('train', 'val', 'test')}
    out_l = {k: [] for k in ('train', 'val', 'test')}
    for i, sp in enumerate(split_of):
        out_s[sp].append(sorted(u + 1 for u in subs[i]))
        out_l[sp].append(int(labels[i]))
    return edges, out_s, {k: np.asarray(v, dtype=np.int64) for k, v in out_l.items()}, values


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
    if name in ('density', 'cutratio') and scale_subgraphs == 1:
        # BASELINE configs[0] / [1]: the reference generator's own recipe (edge editing toward the target property, binned labels)
        edges, subgraphs, labs, _ = generate_property_dataset(name, n=n, m=m, n_subgraphs=w['n_sub'], seed=seed)
        n = int(edges.max()) + 1
        g = DeviceGraph.from_edges(n, edges, device=device, one_indexed=False)
        rs = np.random.RandomState(seed)
        D = hp['node_embed_size']
        emb = np.zeros((n + 1, D), dtype=np.float32)
        emb[1:] = rs.standard_normal((n, D)).astype(np.float32)
        return hp, g, subgraphs, labs, emb
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
