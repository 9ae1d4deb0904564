"""Shared helpers for the tests: golden-file loading into the ``prepared`` dict layout."""
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / 'golden'


def load_npz(name):
    return dict(np.load(GOLDEN / name, allow_pickle=False))


def golden_model(name):
    """Returns (hparams, prepared, raw) for tests/golden/model_<name>.npz."""
    raw = load_npz('model_%s.npz' % name)
    hp = json.loads(str(raw['hparams_json']))
    L = hp['n_layers']
    p = {'num_classes': int(raw['num_classes']), 'multilabel': False, 'embeddings': raw['embeddings'],
         'n_nodes': int(raw['n_nodes']), 'edges': raw['edges'], 'hop': raw['hop'],
         'cc_ids': {}, 'labels': {}, 'sub_G': {}}
    for split in ('train', 'val'):
        p['cc_ids'][split] = raw['cc_ids/' + split]
        p['labels'][split] = raw['labels/' + split]
        lens, flat = raw['sub_G_len/' + split], raw['sub_G_flat/' + split]
        offs = np.concatenate([[0], np.cumsum(lens)])
        p['sub_G'][split] = [flat[offs[i]:offs[i + 1]].tolist() for i in range(len(lens))]
    for key in ('NP_sim', 'I_S_sim', 'B_S_sim', 'N_border'):
        if key + '/train' in raw:
            p[key] = {s: raw[key + '/' + s] for s in ('train', 'val')}
        else:
            p[key] = None
    if hp['use_neighborhood']:
        p['anchors_neigh_int'] = {s: {l: raw['anchors_neigh_int/%s/%d' % (s, l)] for l in range(L)} for s in ('train', 'val')}
        p['anchors_neigh_border'] = {s: {l: raw['anchors_neigh_border/%s/%d' % (s, l)] for l in range(L)} for s in ('train', 'val')}
    if hp['use_position']:
        p['anchors_pos_int'] = {s: {l: raw['anchors_pos_int/%s/%d' % (s, l)] for l in range(L)} for s in ('train', 'val')}
        p['anchors_pos_ext'] = {l: raw['anchors_pos_ext/%d' % l] for l in range(L)}
    if hp['use_structure']:
        p['anchors_structure'] = {l: (raw['anchors_structure/%d/patches' % l], raw['anchors_structure/%d/indices' % l].tolist(),
                                      raw['anchors_structure/%d/int_rw' % l], raw['anchors_structure/%d/bor_rw' % l]) for l in range(L)}
        p['structure_anchors'] = raw['structure_anchors']
    return hp, p, raw


def state_from(raw, prefix):
    import torch
    return {k[len(prefix):]: torch.from_numpy(np.array(v)) for k, v in raw.items() if k.startswith(prefix)}
