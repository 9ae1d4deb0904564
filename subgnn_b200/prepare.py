"""prepare_data on the GPU: everything SubGNN.prepare_data (SubGNN.py:1024-1063) computes once before training,
produced by the kernels of libsubgnn_b200 instead of Python/networkx loops.

The result is the ``prepared`` dict shared by the engine, the SubGNN module and (in tests / bench) the CPU oracle:

  embeddings        fp32 (N+1, D), row 0 = 0                         SubGNN.py:562-568
  num_classes, multilabel
  cc_ids[split]     int64 (n_sub, C, Lcc)                            SubGNN.py:575-607 initialize_cc_ids
  sub_G[split]      list of node-id lists                            SubGNN.py:528-559
  labels[split]     int64 (n_sub,)
  N_border[split]   (ptr int64, ids int32) ragged k-hop border sets  SubGNN.py:673-700  (reference: padded matrix)
  NP_sim            None — similarities are resolved from graph.hop  SubGNN.py:752-781  (reference: dense (n_sub,C,N) slab)
  structure_anchors int64 (P_tot, Ls); int_rw_all / bor_rw_all (P_tot, W, T)   SubGNN.py:889-919
  I_S_sim / B_S_sim[split] fp32 (n_sub, C, P_tot)                    SubGNN.py:783-833
  anchors_neigh_int / anchors_neigh_border[split][l] (n_sub, C, A)   anchor_patch_samplers.py:248-279
  anchors_pos_int[split][l] (n_sub, A_pi); anchors_pos_ext[l] (A_pb) anchor_patch_samplers.py:281-314
  anchors_structure[l] = (patches, indices, int_rw, bor_rw)          anchor_patch_samplers.py:316-328
"""
import numpy as np
import torch

from . import ops
from .graph import DeviceGraph, ragged_from_padded


def connected_components(g, nodes):
    """Components of the subgraph induced on ``nodes`` (1-indexed ids) — SubGNN.py:590-591.  Host side: the
    subgraphs are tiny (10-200 nodes).

    Component ORDER follows the reference, because the similarity / border-set caches are indexed by it along C:
    networkx yields components in the order ``for v in subgraph`` first meets them, and a subgraph view over fewer
    than half of the base graph iterates ``set(nodes)`` (networkx FilterAtlas.__iter__), i.e. CPython set order of
    the ids inserted in file order.  Members are stored ascending (the reference stores python-set order; every
    consumer — pooling, min-hop, border set, degree sequence — is order-free within a component)."""
    seen_order = list(set(int(n) for n in nodes))
    if 2 * len(seen_order) >= g.n_nodes:                   # view iterates the base graph instead: node insertion order
        rank = getattr(g, 'insertion_rank', None)
        seen_order.sort(key=(lambda n: rank[n - 1]) if rank is not None else None)
    idx = {n: i for i, n in enumerate(seen_order)}
    parent = list(range(len(seen_order)))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    rp, col = g.rowptr_host, g.col_host
    arr = np.sort(np.asarray(seen_order, dtype=np.int64))
    for i, n in enumerate(seen_order):
        nb = col[rp[n - 1]:rp[n]] + 1
        hit = nb[np.isin(nb, arr, assume_unique=True)]
        for m in hit:
            a, b = find(i), find(idx[int(m)])
            if a != b:
                parent[max(a, b)] = min(a, b)              # root = member met first in iteration order
    comps = {}
    for i, n in enumerate(seen_order):
        comps.setdefault(find(i), []).append(n)
    return [sorted(comps[k]) for k in sorted(comps)]


def initialize_cc_ids(g, subgraphs):
    """SubGNN.py:575-607 -> int64 (n_sub, max_n_cc, max_len_cc), PAD = 0."""
    cc = [connected_components(g, s) for s in subgraphs]
    C = max(len(c) for c in cc)
    L = max(len(x) for c in cc for x in c)
    out = np.zeros((len(subgraphs), C, L), dtype=np.int64)
    for s, comps in enumerate(cc):
        for c, comp in enumerate(comps):
            out[s, c, :len(comp)] = comp
    return out


def _dev_ragged(rows, device):
    ptr, items = ragged_from_padded(rows)
    return torch.from_numpy(ptr).to(device), torch.from_numpy(items).to(device)


def structure_similarities(g, cc_ids, patches, mode=ops.DTW_FASTDTW_R1):
    """SubGNN.py:783-833 for one split: (internal, border) fp32 (n_sub, C, P_tot); padded components -> 0."""
    dev = g.device
    cc = np.asarray(cc_ids)
    n_sub, C, Lcc = cc.shape
    flat = torch.from_numpy(cc.reshape(n_sub * C, Lcc)).to(dev)
    pt = (patches if isinstance(patches, torch.Tensor) else torch.as_tensor(np.asarray(patches))).to(dev)
    out = []
    for internal in (True, False):
        sa, la = ops.degree_seq(g, flat, internal)
        sb, lb = ops.degree_seq(g, pt, internal)
        sims = ops.dtw_batch(sa, la, sb, lb, mode, max_len_a=Lcc, max_len_b=pt.shape[1])
        out.append(sims.view(n_sub, C, -1))
    return out[0], out[1]


def prepare(hp, g, subgraphs, labels, embeddings, seed=0, splits=('train', 'val'), num_classes=None, dtw_mode=ops.DTW_FASTDTW_R1,
            to_host=True, cache=None, multilabel=False, shared=None):
    """subgraphs / labels: dict split -> list of node-id lists / int array (multilabel: (n_sub, K) 0/1 indicator rows).
    g: DeviceGraph (hop table is computed here if a position/neighbourhood channel needs it).  cache: optional
    formats.SimilarityCache — every product the reference caches under <task>/similarities/ is loaded from there when
    present (unless hp['compute_similarities']) and written there otherwise, under the reference's file names.
    shared: a previously prepared dict whose split-independent parts (structure patches / walks / anchor choice,
    P-border anchors) are reused — prepare_test_data (SubGNN.py:994-1022).  Returns the ``prepared`` dict."""
    dev = g.device
    L = hp['n_layers']
    rs = np.random.RandomState(seed)
    so = 7 if shared is not None else 0          # Philox stream offset of a later-prepared split (test)
    p = {'embeddings': np.asarray(embeddings, dtype=np.float32), 'multilabel': bool(multilabel), 'cc_ids': {}, 'labels': {}, 'sub_G': {},
         'NP_sim': None, 'N_border': {}, 'n_nodes': g.n_nodes}
    if num_classes is None:
        num_classes = (np.asarray(labels[splits[0]]).shape[1] if multilabel else
                       int(np.concatenate([np.asarray(labels[s]).reshape(-1) for s in labels]).max()) + 1)
    p['num_classes'] = int(num_classes)
    if (hp['use_position'] or hp['use_neighborhood']) and g.hop is None:
        ops.hop_table(g)
    for s in splits:
        p['cc_ids'][s] = initialize_cc_ids(g, subgraphs[s])
        lab = np.asarray(labels[s], dtype=np.int64)
        p['labels'][s] = lab.reshape(len(subgraphs[s]), -1) if multilabel else lab.reshape(-1)
        p['sub_G'][s] = [list(map(int, x)) for x in subgraphs[s]]
    cpu = (lambda t: t.cpu().numpy()) if to_host else (lambda t: t)
    load = (lambda name: cache.load(name)) if cache is not None else (lambda name: None)
    save = (lambda name, arr, **kw: cache.save(name, arr.cpu().numpy() if isinstance(arr, torch.Tensor) else arr, **kw)) \
        if cache is not None else (lambda name, arr, **kw: False)
    if hp['use_structure']:
        P_tot = hp['max_sim_epochs'] * hp['n_anchor_patches_structure'] * L                      # anchor_patch_samplers.py:220
        if hp['structure_patch_type'] != 'triangular_random_walk':
            raise NotImplementedError(hp['structure_patch_type'])                                # anchor_patch_samplers.py:233
        if hp.get('structure_similarity_fn', 'dtw') != 'dtw':
            raise NotImplementedError(hp['structure_similarity_fn'])                             # SubGNN.py:826
        if shared is not None:
            patches = torch.as_tensor(np.asarray(shared['structure_anchors'])).to(dev, torch.int32)
            p['structure_anchors'], p['int_rw_all'], p['bor_rw_all'] = shared['structure_anchors'], shared['int_rw_all'], shared['bor_rw_all']
        else:
            got = load(cache.struc_patches()) if cache is not None else None                     # SubGNN.py:893-899
            if got is not None:
                patches = torch.from_numpy(np.ascontiguousarray(got)).to(dev, torch.int32)
            else:
                patches = ops.walk_full(g, P_tot, hp['sample_walk_len'], hp['rw_beta'], seed * 7919 + 1)
                keep = int((patches != 0).sum(dim=0).ne(0).sum().item())                         # pad to the longest walk (:237)
                patches = patches[:, :max(keep, 1)].contiguous()
                if cache is not None:
                    save(cache.struc_patches(), patches.long())
            rw = {}
            for inside, sd in ((False, 3), (True, 2)):                                           # border first, SubGNN.py:903-919
                got = load(cache.walks(inside)) if cache is not None else None
                if got is not None:
                    rw[inside] = torch.from_numpy(np.ascontiguousarray(got)).to(dev, torch.int32)
                else:
                    rw[inside] = ops.walk_patch(g, patches, hp['n_triangular_walks'], hp['random_walk_len'], hp['rw_beta'], not inside,
                                                seed * 7919 + sd)
                    if cache is not None:
                        save(cache.walks(inside), rw[inside].long())
            p['structure_anchors'], p['int_rw_all'], p['bor_rw_all'] = cpu(patches.long()), cpu(rw[True].long()), cpu(rw[False].long())
        p['I_S_sim'], p['B_S_sim'] = {}, {}
        for s in splits:
            got_i = load(cache.struc_sim(True, s)) if cache is not None else None                # SubGNN.py:933-975
            got_b = load(cache.struc_sim(False, s)) if cache is not None else None
            if got_i is None or got_b is None:
                i_s, b_s = structure_similarities(g, p['cc_ids'][s], patches, dtw_mode)
                if cache is not None and got_i is None:
                    save(cache.struc_sim(True, s), i_s)
                if cache is not None and got_b is None:
                    save(cache.struc_sim(False, s), b_s)
            p['I_S_sim'][s] = np.asarray(got_i, dtype=np.float32) if got_i is not None else cpu(i_s)
            p['B_S_sim'][s] = np.asarray(got_b, dtype=np.float32) if got_b is not None else cpu(b_s)
        if shared is not None:
            p['anchors_structure'] = shared['anchors_structure']
        else:
            p['anchors_structure'] = {}
            for l in range(L):                                                                   # anchor_patch_samplers.py:316-328
                idx = rs.choice(P_tot, hp['n_anchor_patches_structure'], replace=True)
                p['anchors_structure'][l] = (p['structure_anchors'][idx], idx.tolist(), p['int_rw_all'][idx], p['bor_rw_all'][idx])
    if (hp['use_position'] or hp['use_neighborhood']) and cache is not None:                     # SubGNN.py:844-874
        dense = {}
        for s in splits:
            got = load(cache.np_sim(s))
            n_sub, C, Lcc = p['cc_ids'][s].shape
            if got is not None and got.shape == (n_sub, C, g.n_nodes):
                dense[s] = np.asarray(got, dtype=np.float32)
            elif 4 * n_sub * C * g.n_nodes <= cache.dense_limit_bytes and cache.write:
                # the reference's dense slab, written for its benefit only: the engine resolves similarities from the hop table
                save(cache.np_sim(s), dense_np_sim(g, p['cc_ids'][s]), dense=True)
        if len(dense) == len(splits):
            p['NP_sim'] = dense
    if hp['use_neighborhood']:
        p['anchors_neigh_int'], p['anchors_neigh_border'] = {}, {}
        for si, s in enumerate(splits):
            cc = p['cc_ids'][s]
            n_sub, C, Lcc = cc.shape
            rptr, ritems = _dev_ragged(cc.reshape(n_sub * C, Lcc), dev)
            got = load(cache.border_set(s)) if cache is not None else None                       # SubGNN.py:726-742
            if got is not None and got.shape[:2] == (n_sub, C):
                bp, bi = ragged_from_padded(np.asarray(got).reshape(n_sub * C, -1))
                bptr, bitems = torch.from_numpy(bp.astype(np.int64)).to(dev), torch.from_numpy(bi).to(dev)
            else:
                bptr, bitems = ops.border_khop(g, rptr, ritems, hp['neigh_sample_border_size'])
                if cache is not None:
                    from .formats import pad_ragged
                    save(cache.border_set(s), pad_ragged(bptr.cpu().numpy(), bitems.cpu().numpy(), (n_sub, C)), dense=True)
            p['N_border'][s] = (cpu(bptr), cpu(bitems))
            width_b = int((bptr[1:] - bptr[:-1]).max().item()) if n_sub else 0
            p['anchors_neigh_int'][s], p['anchors_neigh_border'][s] = {}, {}
            for l in range(L):                                                                   # anchor_patch_samplers.py:275-278
                a_in = ops.sample_rows(rptr, ritems, Lcc, hp['n_anchor_patches_N_in'], True, seed, 100 * (si + so) + 2 * l, False)
                a_out = ops.sample_rows(bptr.to(torch.int32), bitems, width_b, hp['n_anchor_patches_N_out'], True, seed, 100 * (si + so) + 2 * l + 1, False)
                p['anchors_neigh_int'][s][l] = cpu(a_in.view(n_sub, C, -1).long())
                p['anchors_neigh_border'][s][l] = cpu(a_out.view(n_sub, C, -1).long())
    if hp['use_position']:
        p['anchors_pos_int'], p['anchors_pos_ext'] = {}, {}
        allptr = torch.tensor([0, g.n_nodes], dtype=torch.int32, device=dev)
        for l in range(L):                                                                       # anchor_patch_samplers.py:306-314
            if shared is not None:
                p['anchors_pos_ext'][l] = shared['anchors_pos_ext'][l]
                continue
            p['anchors_pos_ext'][l] = cpu(ops.sample_rows(allptr, g.all_nodes(), 0, hp['n_anchor_patches_pos_out'], False, seed, 1000 + l, True)
                                          .view(-1).long())
        for si, s in enumerate(splits):
            lens = np.array([len(x) for x in p['sub_G'][s]], dtype=np.int64)
            sptr = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)).to(dev)
            sitems = torch.from_numpy(np.concatenate(p['sub_G'][s]).astype(np.int32)).to(dev)
            p['anchors_pos_int'][s] = {}
            for l in range(L):                                                                   # anchor_patch_samplers.py:281-304
                p['anchors_pos_int'][s][l] = cpu(ops.sample_rows(sptr, sitems, 0, hp['n_anchor_patches_pos_in'], False, seed,
                                                                 2000 + 100 * (si + so) + l, True).long())
    return p


def dense_np_sim(g, cc_ids):
    """The reference's dense similarity slab (SubGNN.py:752-781) for a (small) set of subgraphs — used to feed the
    CPU oracle; the engine never materialises it."""
    cc = np.asarray(cc_ids)
    n_sub, C, Lcc = cc.shape
    rptr, ritems = _dev_ragged(cc.reshape(n_sub * C, Lcc), g.device)
    return ops.sp_min_dense(g.hop, rptr, ritems).view(n_sub, C, -1).cpu().numpy()


def prepared_subset(p, g, split, indices, as_split='train'):
    """Slices a prepared dict down to ``indices`` of ``split`` (renumbered 0..n-1) and adds the dense NP_sim slab
    for them, i.e. the inputs the reference's per-step path needs for exactly those subgraphs."""
    idx = np.asarray(indices, dtype=np.int64)
    q = {k: p[k] for k in ('embeddings', 'num_classes', 'multilabel', 'n_nodes')}
    q['cc_ids'] = {as_split: np.asarray(p['cc_ids'][split])[idx]}
    q['labels'] = {as_split: np.asarray(p['labels'][split])[idx]}
    q['sub_G'] = {as_split: [p['sub_G'][split][i] for i in idx]}
    for key in ('anchors_neigh_int', 'anchors_neigh_border', 'anchors_pos_int'):
        if key in p:
            q[key] = {as_split: {l: np.asarray(v)[idx] for l, v in p[key][split].items()}}
    for key in ('anchors_pos_ext', 'anchors_structure'):
        if key in p:
            q[key] = p[key]
    for key in ('I_S_sim', 'B_S_sim'):
        q[key] = {as_split: np.asarray(p[key][split])[idx]} if p.get(key) is not None else None
    q['NP_sim'] = {as_split: dense_np_sim(g, q['cc_ids'][as_split])} if g.hop is not None else None
    return q
