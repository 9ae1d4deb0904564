"""anchor_patch_samplers with the reference's entry points (SubGNN/anchor_patch_samplers.py), sampling on the GPU.

Same names, argument order and return conventions as the reference (CPU int64 tensors, PAD = 0); ``networkx_graph``
may be the reference's networkx graph (resolved once to a device CSR) or a DeviceGraph.  Random draws come from
counter-based Philox streams keyed by ``hparams['seed']`` (+ a per-call counter), not from numpy / python / torch
global state, so results are reproducible per call site but not stream-identical to the reference; the laws are
(tests/test_gpu_setup_kernels.py checks them against the reference-stream sampler).
"""
from collections import defaultdict

import numpy as np
import torch

from . import PAD_VALUE, ops
from .graph import ragged_from_padded, resolve_graph

_calls = defaultdict(int)
# fixed call-site table (NOT hash(site): CPython salts str hashes per process, which made the module-level samplers differ from run
# to run and between data-parallel ranks)
_SITES = {'trw': 1, 'prw_in': 2, 'prw_bor': 3, 'ssap': 4, 'snap_in': 5, 'snap_out': 6, 'spap': 7, 'iapi': 8, 'ias': 9}


def _seed(hparams, site):
    """distinct Philox key per call site and call count (a re-sample must differ from the first sample); a function of
    (hparams['seed'], site, number of earlier calls from that site in this process) only."""
    _calls[site] += 1
    return (int(hparams.get('seed', 0)) * 1000003 + _SITES[site] * 911) * 4099 + _calls[site]


def reset_call_counters():
    """start the per-site call counters again (a fresh process starts at zero: same seed => same samples)."""
    _calls.clear()


# ---- walks ---------------------------------------------------------------------------------------------
def triangular_random_walk(hparams, networkx_graph, anchor_patch_subgraph, walk_len, in_border_nodes, all_valid_nodes, inside):
    """anchor_patch_samplers.py:49-113 — one walk.  anchor_patch_subgraph: the full graph (patch sampling) or an
    iterable of the patch's node ids."""
    g = resolve_graph(networkx_graph)
    seed = _seed(hparams, 'trw')
    if anchor_patch_subgraph is networkx_graph:
        w = ops.walk_full(g, 1, walk_len, hparams['rw_beta'], seed)[0]
    else:
        nodes = list(anchor_patch_subgraph.nodes()) if hasattr(anchor_patch_subgraph, 'nodes') else list(anchor_patch_subgraph)
        patch = torch.tensor([nodes], dtype=torch.int64, device=g.device)
        w = ops.walk_patch(g, patch, 1, walk_len, hparams['rw_beta'], not inside, seed)[0, 0]
    return [int(v) for v in w.cpu().tolist() if v != PAD_VALUE]


def perform_random_walks(hparams, networkx_graph, anchor_patch_ids, inside):
    """anchor_patch_samplers.py:118-158 -> LongTensor (n_patches, n_triangular_walks, random_walk_len)."""
    g = resolve_graph(networkx_graph)
    out = ops.walk_patch(g, torch.as_tensor(anchor_patch_ids).to(g.device), hparams['n_triangular_walks'], hparams['random_walk_len'],
                         hparams['rw_beta'], not inside, _seed(hparams, 'prw_in' if inside else 'prw_bor'))
    return out.long().cpu()


def sample_structure_anchor_patches(hparams, networkx_graph, device, max_sim_epochs):
    """anchor_patch_samplers.py:210-243 -> LongTensor (n_samples, max patch length)."""
    g = resolve_graph(networkx_graph)
    n_samples = max_sim_epochs * hparams['n_anchor_patches_structure'] * hparams['n_layers']
    if hparams['structure_patch_type'] != 'triangular_random_walk':
        raise NotImplementedError                                                              # :232-233 (ego_graph patches are not built)
    p = ops.walk_full(g, n_samples, hparams['sample_walk_len'], hparams['rw_beta'], _seed(hparams, 'ssap'))
    keep = max(int((p != 0).sum(dim=0).ne(0).sum().item()), 1)                                  # :236-241 pad to the longest walk
    return p[:, :keep].long().cpu()


# ---- neighbourhood / position sampling -------------------------------------------------------------------
def sample_neighborhood_anchor_patch(hparams, networkx_graph, cc_ids, border_set, sample_inside=True):
    """anchor_patch_samplers.py:163-198 -> LongTensor (n_sub, max_n_cc, n_anchor_patches_N_in | _out)."""
    src = cc_ids if sample_inside else border_set
    n_sub, C, width = src.shape
    dev = torch.device('cuda')
    ptr, items = ragged_from_padded(np.asarray(src.cpu()).reshape(n_sub * C, width))
    A = hparams['n_anchor_patches_N_in'] if sample_inside else hparams['n_anchor_patches_N_out']
    out = ops.sample_rows(torch.from_numpy(ptr).to(dev), torch.from_numpy(items).to(dev), width, A, True,
                          _seed(hparams, 'snap_in' if sample_inside else 'snap_out'), 0, False)
    return out.view(n_sub, C, A).long().cpu()


def sample_position_anchor_patches(hparams, networkx_graph, subgraph=None):
    """anchor_patch_samplers.py:200-208 -> list of sampled node ids."""
    g = resolve_graph(networkx_graph)
    dev = g.device
    if not subgraph:
        pool, A = g.all_nodes(), hparams['n_anchor_patches_pos_out']
    else:
        pool, A = torch.tensor(list(subgraph), dtype=torch.int32, device=dev), hparams['n_anchor_patches_pos_in']
    ptr = torch.tensor([0, pool.numel()], dtype=torch.int32, device=dev)
    return ops.sample_rows(ptr, pool, 0, A, False, _seed(hparams, 'spap'), 0, True).view(-1).cpu().tolist()


def _split_sets(split, train, val, test):
    if split == 'all':
        return ['train', 'val', 'test'], [train, val, test]
    if split == 'train_val':
        return ['train', 'val'], [train, val]
    if split == 'test':
        return ['test'], [test]
    raise ValueError(split)


def init_anchors_neighborhood(split, hparams, networkx_graph, device, train_cc_ids, val_cc_ids, test_cc_ids, train_N_border, val_N_border,
                              test_N_border):
    """anchor_patch_samplers.py:248-279"""
    names, sets = _split_sets(split, train_cc_ids, val_cc_ids, test_cc_ids)
    _, borders = _split_sets(split, train_N_border, val_N_border, test_N_border)
    anchors_int, anchors_border = defaultdict(dict), defaultdict(dict)
    for name, cc, bs in zip(names, sets, borders):
        for n in range(hparams['n_layers']):
            anchors_int[name][n] = sample_neighborhood_anchor_patch(hparams, networkx_graph, cc, bs, sample_inside=True)
            anchors_border[name][n] = sample_neighborhood_anchor_patch(hparams, networkx_graph, cc, bs, sample_inside=False)
    return anchors_int, anchors_border


def init_anchors_pos_int(split, hparams, networkx_graph, device, train_cc_ids, val_cc_ids, test_cc_ids):
    """anchor_patch_samplers.py:281-304 — the 5th-7th arguments are lists of subgraph node lists."""
    g = resolve_graph(networkx_graph)
    names, sets = _split_sets(split, train_cc_ids, val_cc_ids, test_cc_ids)
    out = defaultdict(dict)
    for name, subs in zip(names, sets):
        lens = np.array([len(s) for s in subs], dtype=np.int64)
        ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)).to(g.device)
        items = torch.from_numpy(np.concatenate([np.asarray(s, dtype=np.int32) for s in subs])).to(g.device)
        for n in range(hparams['n_layers']):
            out[name][n] = ops.sample_rows(ptr, items, 0, hparams['n_anchor_patches_pos_in'], False, _seed(hparams, 'iapi'), n, True).long().cpu()
    return out


def init_anchors_pos_ext(hparams, networkx_graph, device):
    """anchor_patch_samplers.py:306-314"""
    return {n: torch.tensor(sample_position_anchor_patches(hparams, networkx_graph)) for n in range(hparams['n_layers'])}


def init_anchors_structure(hparams, structure_anchors, int_structure_anchor_rw, bor_structure_anchor_rw):
    """anchor_patch_samplers.py:316-328 (index choice with replacement; host side, a handful of integers)."""
    rs = np.random.RandomState(_seed(hparams, 'ias') % (2 ** 31))
    out = {}
    for n in range(hparams['n_layers']):
        indices = list(rs.choice(structure_anchors.shape[0], hparams['n_anchor_patches_structure'], replace=True))
        out[n] = (structure_anchors[indices, :], indices, int_structure_anchor_rw[indices, :, :], bor_structure_anchor_rw[indices, :, :])
    return out


# ---- per-step lookup (module-level compatibility path; the fused engine never materialises these) -----------
def embed_anchor_patch(node_matrix, anchor_patch_ids, device):
    """anchor_patch_samplers.py:404-411"""
    ids = anchor_patch_ids.to(device)
    return node_matrix(ids), (ids != PAD_VALUE).bool()


def aggregate_structure_anchor_patch(hparams, networkx_graph, lstm, node_matrix, anchor_patch_ids, all_patch_walks, inside, device):
    """anchor_patch_samplers.py:413-433: embed the pre-sampled walks, run the walk encoder, sum over the walks."""
    n_patches = anchor_patch_ids.shape[0]
    walk_embeds, _ = embed_anchor_patch(node_matrix, all_patch_walks, device)
    walk_hidden = lstm(walk_embeds.view(n_patches * hparams['n_triangular_walks'], hparams['random_walk_len'], hparams['node_embed_size']))
    return walk_hidden.view(n_patches, hparams['n_triangular_walks'], -1).sum(dim=1)


def get_anchor_patches(dataset_type, hparams, networkx_graph, node_matrix, subgraph_idx, cc_ids, cc_embed_mask, lstm, anchors_neigh_int,
                       anchors_neigh_border, anchors_pos_int, anchors_pos_ext, anchors_structure, layer_num, channel, inside, device=None):
    """anchor_patch_samplers.py:333-399 -> (anchor_patches, anchor_mask, anchor_embeds)."""
    batch_sz, max_n_cc, _ = cc_ids.shape
    dev = cc_ids.device
    idx = subgraph_idx.to('cpu')
    if channel == 'neighborhood':
        src = anchors_neigh_int if inside else anchors_neigh_border
        patches = src[dataset_type][layer_num][idx].squeeze(1).to(dev)
        embeds, mask = embed_anchor_patch(node_matrix, patches, dev)
        return patches.unsqueeze(-1), mask.unsqueeze(-1), embeds
    if channel == 'position':
        if inside:
            patches = anchors_pos_int[dataset_type][layer_num][idx].squeeze(1).unsqueeze(1).repeat(1, max_n_cc, 1).to(dev)
        else:
            patches = anchors_pos_ext[layer_num].unsqueeze(0).unsqueeze(0).repeat(batch_sz, max_n_cc, 1).to(dev)
        patches[~cc_embed_mask] = PAD_VALUE
        embeds, mask = embed_anchor_patch(node_matrix, patches, dev)
        return patches.unsqueeze(-1), mask.unsqueeze(-1), embeds
    if channel == 'structure':
        patches, indices, int_rw, bor_rw = anchors_structure[layer_num]
        embeds = aggregate_structure_anchor_patch(hparams, networkx_graph, lstm, node_matrix, patches, int_rw if inside else bor_rw,
                                                  inside=inside, device=dev)
        patches = patches.to(dev).unsqueeze(0).unsqueeze(0).repeat(batch_sz, max_n_cc, 1, 1)
        patches[~cc_embed_mask] = PAD_VALUE
        mask = (patches != PAD_VALUE).bool()
        embeds = embeds.unsqueeze(0).unsqueeze(0).repeat(batch_sz, max_n_cc, 1, 1)
        embeds[~cc_embed_mask] = PAD_VALUE
        return patches, mask, embeds
    raise Exception('An invalid channel has been entered.')                                    # :396
