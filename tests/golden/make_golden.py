"""Generates the committed golden vectors under tests/golden/ by running the UNMODIFIED
reference modules (/root/reference, imported under stubs by oracle/ref_loader.py) on tiny
seeded inputs.  Build-container only (the reference is absent on the GPU box).

    python tests/golden/make_golden.py

Outputs (all small, committed):
  walks_golden.npz     reference walks / structure patches / border nodes for fixed MT seeds
  gamma_golden.npz     reference degree sequences (+ DTW similarities through the restated fastdtw)
  sampling_golden.npz  reference cc ids, k-hop border sets, SP-min similarity, N-anchor samples
  model_<cfg>.npz      a full reference SubGNN run: prepared tensors, initial weights, per-step
                       loss / logits for K Adam steps, final weights
"""
import json
import os
import random
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
os.environ['CUDA_VISIBLE_DEVICES'] = ''

import networkx as nx  # noqa: E402
import torch  # noqa: E402

from oracle import ref_loader  # noqa: E402

OUT = Path(__file__).resolve().parent


def small_graph(n=60, m=3, seed=7, extra_pendants=4):
    """BA graph + a few pendant nodes + one isolated pair, nodes 0..N-1, edges sorted so that
    nx.read_edgelist inserts nodes in id order."""
    G = nx.barabasi_albert_graph(n, m, seed=seed)
    rng = random.Random(seed)
    for k in range(extra_pendants):
        G.add_edge(rng.randrange(n), n + k)
    edges = sorted((min(u, v), max(u, v)) for u, v in G.edges())
    edges.sort(key=lambda e: (e[1], e[0]))
    return G.number_of_nodes(), edges


def write_dataset(root, n_nodes, edges, subgraphs, labels, splits, D, seed):
    root = Path(root)
    (root / 'similarities').mkdir(parents=True, exist_ok=True)
    with open(root / 'edge_list.txt', 'w') as f:
        for u, v in edges:
            f.write('%d %d\n' % (u, v))
    with open(root / 'subgraphs.pth', 'w') as f:
        for s, l, sp in zip(subgraphs, labels, splits):
            f.write('-'.join(str(n) for n in s) + '\t' + str(l) + '\t' + sp + '\n')
    g = torch.Generator().manual_seed(seed)
    torch.save(torch.randn(n_nodes, D, generator=g), root / 'emb.pth')
    # hop table with the reference's conventions (0 = self / unreachable)
    G = nx.Graph()
    G.add_nodes_from(range(n_nodes))
    G.add_edges_from(edges)
    sp = np.zeros((n_nodes, n_nodes))
    for s, d in nx.all_pairs_shortest_path_length(G):
        for t, v in d.items():
            sp[s, t] = v
    np.save(root / 'shortest_path_matrix.npy', sp)
    return sp


def make_subgraphs(n_nodes, edges, n_sub, seed, n_classes=3):
    G = nx.Graph()
    G.add_nodes_from(range(n_nodes))
    G.add_edges_from(edges)
    rng = random.Random(seed)
    subs, labels, splits = [], [], []
    for i in range(n_sub):
        nodes = set()
        for _ in range(rng.randint(1, 3)):          # 1-3 BFS chunks => multi-component subgraphs
            src = rng.randrange(n_nodes)
            chunk = list(nx.bfs_tree(G, src, depth_limit=1).nodes())[:rng.randint(1, 5)]
            nodes.update(chunk)
        subs.append(sorted(nodes))
        labels.append('c%d' % rng.randrange(n_classes))
        splits.append('train' if i < int(0.7 * n_sub) else ('val' if i < int(0.85 * n_sub) else 'test'))
    # make sure every class label string appears (label ids are assigned in file order)
    for c in range(n_classes):
        labels[c] = 'c%d' % c
    return subs, labels, splits


BASE_HP = {
    "max_epochs": 1, "seed": 0, "structure_patch_type": "triangular_random_walk", "lstm_aggregator": "last",
    "n_processes": 1, "resample_anchor_patches": False, "freeze_node_embeds": False, "use_mpn_projection": True,
    "print_train_times": False, "compute_similarities": True, "sample_walk_len": 12, "n_triangular_walks": 3,
    "random_walk_len": 6, "rw_beta": 0.65, "learning_rate": 1e-2, "grad_clip": 0.25, "neigh_sample_border_size": 1,
    "n_anchor_patches_pos_out": 7, "n_anchor_patches_pos_in": 5, "n_anchor_patches_N_in": 4, "n_anchor_patches_N_out": 6,
    "n_anchor_patches_structure": 5, "linear_hidden_dim_1": 16, "linear_hidden_dim_2": 12, "lin_dropout": 0.0,
    "lstm_dropout": 0.0, "max_sim_epochs": 2, "auto_lr_find": False,
}

CONFIGS = {
    # name: overrides
    'all_L1_max': dict(use_neighborhood=True, use_position=True, use_structure=True, node_embed_size=8, batch_size=6,
                       n_layers=1, lstm_n_layers=1, cc_aggregator='max', trainable_cc=False),
    'all_L2_sum': dict(use_neighborhood=True, use_position=True, use_structure=True, node_embed_size=8, batch_size=5,
                       n_layers=2, lstm_n_layers=2, cc_aggregator='sum', trainable_cc=False, neigh_sample_border_size=2),
    'S_L2_sumagg': dict(use_neighborhood=False, use_position=False, use_structure=True, node_embed_size=8, batch_size=7,
                        n_layers=2, lstm_n_layers=2, cc_aggregator='sum', trainable_cc=False, lstm_aggregator='sum'),
    'NP_L2_trainable': dict(use_neighborhood=True, use_position=True, use_structure=False, node_embed_size=8, batch_size=6,
                            n_layers=2, lstm_n_layers=1, cc_aggregator='sum', trainable_cc=True),
}


def to_np(x):
    return x.detach().cpu().numpy().copy() if isinstance(x, torch.Tensor) else np.array(x)


def run_model_golden(ref, name, overrides, n_steps=4):
    hp = dict(BASE_HP)
    hp.update(overrides)
    tmp = Path(tempfile.mkdtemp(prefix='subgnn_golden_'))
    ref.config.PROJECT_ROOT = tmp
    n_nodes, edges = small_graph()
    subs, labels, splits = make_subgraphs(n_nodes, edges, 30, seed=11)
    write_dataset(tmp, n_nodes, edges, subs, labels, splits, hp['node_embed_size'], seed=5)
    torch.manual_seed(hp['seed'])
    np.random.seed(hp['seed'])
    random.seed(hp['seed'])
    model = ref.SubGNN.SubGNN(hp, 'edge_list.txt', 'subgraphs.pth', 'emb.pth', 'similarities',
                              'shortest_path_matrix.npy', 'degree_sequence.txt', 'ego_graphs.txt')
    model.prepare_data()
    # F15: torch>=2 keeps the transposed strides of the N anchors; the reference's .view needs contiguity
    if hp['use_neighborhood']:
        for d in (model.anchors_neigh_int, model.anchors_neigh_border):
            for split in d:
                for l in d[split]:
                    d[split][l] = d[split][l].contiguous()
    out = {'hparams_json': json.dumps(hp), 'n_nodes': n_nodes, 'edges': np.array(edges, dtype=np.int64),
           'num_classes': model.num_classes, 'embeddings': to_np(model.node_embeddings.weight)}
    for split in ('train', 'val'):
        out['cc_ids/' + split] = to_np(getattr(model, split + '_cc_ids'))
        out['labels/' + split] = to_np(getattr(model, split + '_sub_G_label'))
        subs_s = getattr(model, split + '_sub_G')
        out['sub_G_len/' + split] = np.array([len(s) for s in subs_s])
        out['sub_G_flat/' + split] = np.array([n for s in subs_s for n in s], dtype=np.int64)
        if hp['use_neighborhood']:
            out['N_border/' + split] = to_np(getattr(model, split + '_N_border'))
        if hp['use_neighborhood'] or hp['use_position']:
            out['NP_sim/' + split] = to_np(getattr(model, split + '_neigh_pos_similarities'))
        if hp['use_structure']:
            out['I_S_sim/' + split] = to_np(getattr(model, split + '_int_struc_similarities'))
            out['B_S_sim/' + split] = to_np(getattr(model, split + '_bor_struc_similarities'))
        for l in range(hp['n_layers']):
            if hp['use_neighborhood']:
                out['anchors_neigh_int/%s/%d' % (split, l)] = to_np(model.anchors_neigh_int[split][l])
                out['anchors_neigh_border/%s/%d' % (split, l)] = to_np(model.anchors_neigh_border[split][l])
            if hp['use_position']:
                out['anchors_pos_int/%s/%d' % (split, l)] = to_np(model.anchors_pos_int[split][l])
    out['hop'] = np.load(tmp / 'shortest_path_matrix.npy').astype(np.uint8)
    for l in range(hp['n_layers']):
        if hp['use_position']:
            out['anchors_pos_ext/%d' % l] = to_np(model.anchors_pos_ext[l])
        if hp['use_structure']:
            patches, indices, irw, brw = model.anchors_structure[l]
            out['anchors_structure/%d/patches' % l] = to_np(patches)
            out['anchors_structure/%d/indices' % l] = np.array(indices, dtype=np.int64)
            out['anchors_structure/%d/int_rw' % l] = to_np(irw)
            out['anchors_structure/%d/bor_rw' % l] = to_np(brw)
    if hp['use_structure']:
        out['structure_anchors'] = to_np(model.structure_anchors)
        out['int_rw_all'] = to_np(model.int_structure_anchor_random_walks)
        out['bor_rw_all'] = to_np(model.bor_structure_anchor_random_walks)
    for k, v in model.state_dict().items():
        out['init/' + k] = to_np(v)
    # K Adam steps over fixed batches (no shuffling), Lightning 0.7.1's loop restated:
    # training_step -> backward(retain_graph=True) -> clip_grad_norm_ -> optimizer.step
    opt = model.configure_optimizers()
    loader = model.train_dataloader()
    ds = loader.dataset
    n_train = len(ds)
    B = hp['batch_size']
    order = list(range(n_train))
    batches = [order[i:i + B] for i in range(0, n_train - B + 1, B)]
    model.train()
    step = 0
    for it in range(n_steps):
        idxs = batches[it % len(batches)]
        batch = model._pad_collate([ds[i] for i in idxs])
        res = model.training_step(batch, it)
        logits = model.forward('train', model.train_N_I_cc_embed, model.train_N_B_cc_embed, model.train_S_I_cc_embed,
                               model.train_S_B_cc_embed, model.train_P_I_cc_embed, model.train_P_B_cc_embed,
                               batch['subgraph_ids'], batch['cc_ids'], batch['subgraph_idx'], batch['NP_sim'],
                               batch['I_S_sim'], batch['B_S_sim'])
        opt.zero_grad()
        model.backward(None, res['loss'], opt, 0)
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), hp['grad_clip'])
        if it == 0:
            for k, prm in model.named_parameters():
                if prm.grad is not None:
                    out['grad0/' + k] = to_np(prm.grad)       # clipped gradient of step 0
        opt.step()
        out['step/%d/idx' % it] = np.array(idxs)
        out['step/%d/loss' % it] = to_np(res['loss'])
        out['step/%d/logits' % it] = to_np(logits)
        out['step/%d/grad_norm' % it] = to_np(gn)
        step += 1
    for k, v in model.state_dict().items():
        out['final/' + k] = to_np(v)
    # one validation forward
    model.eval()
    vds = model.val_dataloader().dataset
    vb = model._pad_collate([vds[i] for i in range(len(vds))])
    with torch.no_grad():
        r = model.validation_step(vb, 0)
    out['val/logits'] = to_np(r['val_logits'])
    out['val/loss'] = to_np(r['val_loss'])
    np.savez_compressed(OUT / ('model_%s.npz' % name), **out)
    shutil.rmtree(tmp)
    print('model golden', name, 'losses', [float(out['step/%d/loss' % i]) for i in range(n_steps)])


def run_walk_golden(ref):
    aps = ref.anchor_patch_samplers
    n_nodes, edges = small_graph(n=50, m=3, seed=3)
    G = nx.Graph()
    G.add_nodes_from(range(1, n_nodes + 1))
    G.add_edges_from((u + 1, v + 1) for u, v in edges)
    hp = {'n_anchor_patches_structure': 6, 'n_layers': 2, 'structure_patch_type': 'triangular_random_walk',
          'sample_walk_len': 14, 'rw_beta': 0.65, 'n_triangular_walks': 4, 'random_walk_len': 7}
    out = {'n_nodes': n_nodes, 'edges': np.array(edges, dtype=np.int64), 'hparams_json': json.dumps(hp)}
    for seed in (0, 1, 2):
        np.random.seed(seed)
        random.seed(seed)
        patches = aps.sample_structure_anchor_patches(hp, G, None, 2)
        out['patches/%d' % seed] = to_np(patches)
        np.random.seed(seed + 100)
        random.seed(seed + 100)
        out['int_rw/%d' % seed] = to_np(aps.perform_random_walks(hp, G, patches, inside=True))
        np.random.seed(seed + 200)
        random.seed(seed + 200)
        out['bor_rw/%d' % seed] = to_np(aps.perform_random_walks(hp, G, patches, inside=False))
    # border nodes of a few patches
    pts = to_np(patches)
    for i in range(4):
        nodes = pts[i][pts[i] != 0]
        sub = G.subgraph(nodes)
        b, non = ref.subgraph_utils.get_border_nodes(G, sub)
        out['border_nodes/%d' % i] = np.sort(np.asarray(b).reshape(-1))
        out['border_patch/%d' % i] = nodes
    np.savez_compressed(OUT / 'walks_golden.npz', **out)
    print('walk golden ok', {k: v.shape for k, v in out.items() if k.startswith('patches')})


def run_gamma_sampling_golden(ref):
    n_nodes, edges = small_graph(n=40, m=2, seed=9)
    G = nx.Graph()
    G.add_nodes_from(range(1, n_nodes + 1))
    G.add_edges_from((u + 1, v + 1) for u, v in edges)
    rng = np.random.RandomState(0)
    out = {'n_nodes': n_nodes, 'edges': np.array(edges, dtype=np.int64)}
    rows = []
    for i in range(24):
        L = rng.randint(1, 12)
        row = np.zeros(14, dtype=np.int64)
        row[:L] = rng.randint(1, n_nodes + 1, size=L)        # duplicates on purpose (F9)
        rows.append(row)
    rows.append(np.zeros(14, dtype=np.int64))                 # all PAD
    rows = np.stack(rows)
    out['rows'] = rows
    seq_i = [ref.gamma.get_degree_sequence(G, torch.tensor(r), None, internal=True) for r in rows]
    seq_b = [ref.gamma.get_degree_sequence(G, torch.tensor(r), None, internal=False) for r in rows]
    out['seq_int_len'] = np.array([len(s) for s in seq_i])
    out['seq_int'] = np.array([v for s in seq_i for v in s], dtype=np.int64)
    out['seq_bor_len'] = np.array([len(s) for s in seq_b])
    out['seq_bor'] = np.array([v for s in seq_b for v in s], dtype=np.int64)
    sims = np.zeros((len(rows) - 1, len(rows) - 1))
    for i in range(len(rows) - 1):
        for j in range(len(rows) - 1):
            sims[i, j] = ref.gamma.calc_dtw(seq_i[i], seq_b[j])
    out['dtw_sims'] = sims
    np.savez_compressed(OUT / 'gamma_golden.npz', **out)

    # cc ids / border sets / SP-min / N sampling
    out = {'n_nodes': n_nodes, 'edges': np.array(edges, dtype=np.int64)}
    subs = []
    r2 = random.Random(4)
    for i in range(12):
        subs.append(sorted(set(r2.randrange(1, n_nodes + 1) for _ in range(r2.randint(2, 9)))))
    out['sub_len'] = np.array([len(s) for s in subs])
    out['sub_flat'] = np.array([n for s in subs for n in s], dtype=np.int64)
    holder = type('H', (), {})()
    holder.networkx_graph = G
    cc_ids = ref.SubGNN.SubGNN.initialize_cc_ids(holder, subs)
    out['cc_ids'] = to_np(cc_ids)
    for k in (1, 2):
        sets = [[sorted(ref.subgraph_utils.get_component_border_neighborhood_set(G, comp, k, None)) for comp in sub] for sub in cc_ids]
        Lb = max(len(x) for row in sets for x in row)
        arr = np.zeros((cc_ids.shape[0], cc_ids.shape[1], Lb), dtype=np.int64)
        for s, row in enumerate(sets):
            for c, x in enumerate(row):
                arr[s, c, :len(x)] = x
        out['border_k%d' % k] = arr
    sp = np.zeros((n_nodes, n_nodes))
    for s, d in nx.all_pairs_shortest_path_length(G):
        for t, v in d.items():
            sp[s - 1, t - 1] = v
    out['hop'] = sp.astype(np.uint8)
    tmp = Path(tempfile.mkdtemp(prefix='subgnn_golden_'))
    sims = ref.SubGNN.SubGNN.compute_shortest_path_similarities(holder, tmp / 'x' / 'sim.npy', sp, cc_ids)
    out['sp_sim'] = to_np(sims)
    shutil.rmtree(tmp)
    hp = {'n_anchor_patches_N_in': 5, 'n_anchor_patches_N_out': 3}
    torch.manual_seed(123)
    out['N_in_seed123'] = to_np(ref.anchor_patch_samplers.sample_neighborhood_anchor_patch(hp, G, cc_ids, None, True).contiguous())
    np.savez_compressed(OUT / 'sampling_golden.npz', **out)
    print('gamma/sampling golden ok')


if __name__ == '__main__':
    ref = ref_loader.load(tempfile.gettempdir())
    run_walk_golden(ref)
    run_gamma_sampling_golden(ref)
    for name, ov in CONFIGS.items():
        run_model_golden(ref, name, ov)
