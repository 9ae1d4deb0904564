"""Generates tests/golden/task_tiny/: a complete SubGNN task directory in the reference's on-disk formats whose
similarity caches were WRITTEN BY THE UNMODIFIED REFERENCE (imported under stubs, oracle/ref_loader.py), plus
tests/golden/metrics_golden.json (the reference's epoch-end metrics for fixed logits / labels).  Build-container only.

    python tests/golden/make_task_golden.py

The GPU tests load this directory through subgnn_b200.SubGNN (reference caches must load directly), and
recompute every cached product with the CUDA kernels to compare with what the reference wrote.
"""
import json
import os
import random
import shutil
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))
os.environ['CUDA_VISIBLE_DEVICES'] = ''

import networkx as nx  # noqa: E402
import torch  # noqa: E402

from make_golden import make_subgraphs, small_graph  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = Path(__file__).resolve().parent
TASK = OUT / 'task_tiny'

HP = {
    "max_epochs": 2, "seed": 0, "structure_patch_type": "triangular_random_walk", "lstm_aggregator": "last", "n_processes": 1,
    "resample_anchor_patches": False, "freeze_node_embeds": False, "use_mpn_projection": True, "print_train_times": False,
    "compute_similarities": False, "sample_walk_len": 12, "n_triangular_walks": 3, "random_walk_len": 6, "rw_beta": 0.65,
    "learning_rate": 1e-2, "grad_clip": 0.25, "neigh_sample_border_size": 2, "n_anchor_patches_pos_out": 7, "n_anchor_patches_pos_in": 5,
    "n_anchor_patches_N_in": 4, "n_anchor_patches_N_out": 6, "n_anchor_patches_structure": 5, "linear_hidden_dim_1": 16,
    "linear_hidden_dim_2": 12, "lin_dropout": 0.0, "lstm_dropout": 0.0, "max_sim_epochs": 2, "auto_lr_find": False,
    "use_neighborhood": True, "use_position": True, "use_structure": True, "node_embed_size": 8, "batch_size": 6, "n_layers": 2,
    "lstm_n_layers": 1, "cc_aggregator": "sum", "trainable_cc": False, "embedding_type": "gin", "ff_attn": False,
}


def main():
    ref = ref_loader.load()
    if TASK.exists():
        shutil.rmtree(TASK)
    (TASK / 'similarities').mkdir(parents=True)
    n_nodes, edges = small_graph()
    subs, labels, splits = make_subgraphs(n_nodes, edges, 36, seed=23)
    with open(TASK / 'edge_list.txt', 'w') as f:
        for u, v in edges:
            f.write('%d %d\n' % (u, v))
    with open(TASK / 'subgraphs.pth', 'w') as f:
        for s, l, sp in zip(subs, labels, splits):
            f.write('-'.join(str(n) for n in s) + '\t' + str(l) + '\t' + sp + '\n')
    g = torch.Generator().manual_seed(5)
    torch.save(torch.randn(n_nodes, HP['node_embed_size'], generator=g), TASK / 'gin_embeddings.pth')
    G = nx.Graph()
    G.add_nodes_from(range(n_nodes))
    G.add_edges_from(edges)
    sp = np.zeros((n_nodes, n_nodes))
    for s, d in nx.all_pairs_shortest_path_length(G):
        for t, v in d.items():
            sp[s, t] = v
    np.save(TASK / 'shortest_path_matrix.npy', sp)
    # precompute_graph_metrics.py:34-59 layouts (snap is not installed: same content from networkx)
    with open(TASK / 'degree_sequence.txt', 'w') as f:
        json.dump({str(n): int(G.degree(n)) for n in sorted(G.nodes())}, f)
    with open(TASK / 'ego_graphs.txt', 'w') as f:
        json.dump({str(n): sorted(int(x) for x in G.neighbors(n)) for n in sorted(G.nodes())}, f)
    with open(TASK / 'hyperparams.json', 'w') as f:
        json.dump(HP, f, indent=1)

    ref.config.PROJECT_ROOT = OUT
    hp = dict(HP)
    torch.manual_seed(hp['seed'])
    np.random.seed(hp['seed'])
    random.seed(hp['seed'])
    # degree_sequence / ego_graphs paths point at absent files so that the reference takes its compute-from-graph branches
    # (the ego-graph dict holds 1-hop sets; with neigh_sample_border_size = 2 it must not be used)
    model = ref.SubGNN.SubGNN(hp, 'task_tiny/edge_list.txt', 'task_tiny/subgraphs.pth', 'task_tiny/gin_embeddings.pth', 'task_tiny/similarities',
                              'task_tiny/shortest_path_matrix.npy', 'task_tiny/none_degree.txt', 'task_tiny/none_ego.txt')
    model.prepare_data()
    model.prepare_test_data()
    written = sorted(p.name for p in (TASK / 'similarities').iterdir())
    print('reference wrote:', written)

    # ---- epoch-end metrics golden (SubGNN.py:408-504) ----
    rs = np.random.RandomState(3)
    cases = {}

    def run(kind, logits, labels, multilabel, n_batches=3):
        holder = type('H', (), {})()
        holder.multilabel = multilabel
        holder.multilabel_binarizer = object() if multilabel else None
        holder.hparams = {'trainable_cc': True, 'resample_anchor_patches': False}
        holder.metric_scores = []
        outs = []
        lg, lb = torch.tensor(logits, dtype=torch.float32), torch.tensor(labels)
        loss = ref.SubGNN.nn.BCEWithLogitsLoss() if multilabel else ref.SubGNN.nn.CrossEntropyLoss()
        for ch_l, ch_y in zip(torch.chunk(lg, n_batches), torch.chunk(lb, n_batches)):
            l_ = loss(ch_l, ch_y.type_as(ch_l)) if multilabel else loss(ch_l, ch_y)
            outs.append({kind + '_loss': l_, kind + '_acc': ref.subgraph_utils.calc_accuracy(ch_l, ch_y, multilabel_binarizer=holder.multilabel_binarizer),
                         kind + '_macro_f1': ref.subgraph_utils.calc_f1(ch_l, ch_y, avg_type='macro', multilabel_binarizer=holder.multilabel_binarizer),
                         kind + '_logits': ch_l, kind + '_labels': ch_y})
        fn = ref.SubGNN.SubGNN.validation_epoch_end if kind == 'val' else ref.SubGNN.SubGNN.test_epoch_end
        res = fn(holder, outs)
        return {k: float(v) for k, v in res['log'].items()}

    for name, K, multilabel in (('multiclass', 4, False), ('binary', 2, False), ('multilabel', 5, True)):
        n = 30
        logits = rs.randn(n, K).astype(np.float32) * 2
        if multilabel:
            labels = (rs.rand(n, K) < 0.4).astype(np.int64)
            labels[0], labels[1] = 1, 0                      # both classes present in every column
            labels[1, :] = 1 - labels[0, :]
        else:
            labels = rs.randint(K, size=n).astype(np.int64)
            labels[:K] = np.arange(K)
        cases[name] = {'logits': logits.tolist(), 'labels': labels.tolist(), 'multilabel': multilabel,
                       'val': run('val', logits, labels, multilabel), 'test': run('test', logits, labels, multilabel)}
    with open(OUT / 'metrics_golden.json', 'w') as f:
        json.dump(cases, f)
    print('metrics golden ok:', {k: sorted(v['val'])[:4] for k, v in cases.items()})


if __name__ == '__main__':
    main()
