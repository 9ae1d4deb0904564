"""CPU-only checks of SURVEY 8f rows 1, 2 and 4: on-disk formats against a task directory whose caches the unmodified
reference wrote (tests/golden/task_tiny, generator: tests/golden/make_task_golden.py), epoch-end metrics against the
reference's numbers (tests/golden/metrics_golden.json), and the driver stand-ins (pytorch_lightning 0.7.1 protocol, optuna,
commentjson) — including the reference's own train_config.py run UNCHANGED on top of them with the reference model on the
CPU when /root/reference is present."""
import json
import os
import runpy
import shutil
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

GOLDEN = Path(__file__).resolve().parent / 'golden'
TASK = GOLDEN / 'task_tiny'


def _hp():
    return json.loads((TASK / 'hyperparams.json').read_text())


# ---- 8f-2: formats ---------------------------------------------------------------------------------------------------
def test_cache_names_are_the_ones_the_reference_wrote():
    from subgnn_b200.formats import SimilarityCache
    c = SimilarityCache(TASK / 'similarities', _hp())
    want = {c.struc_patches(), c.walks(True), c.walks(False)}
    for s in ('train', 'val', 'test'):
        want |= {c.border_set(s), c.np_sim(s), c.struc_sim(True, s), c.struc_sim(False, s)}
    assert want == {p.name for p in (TASK / 'similarities').iterdir()}
    hp = dict(_hp(), structure_similarity_fn='edit_distance')
    assert SimilarityCache('.', hp).struc_sim(True, 'train') == 'int_struc_12_triangular_random_walk_2_0_edit_distance_train_similarities.npy'
    assert c.load(c.struc_patches()).dtype == np.int64 and c.loaded == [c.struc_patches()]
    assert SimilarityCache(TASK / 'similarities', dict(_hp(), compute_similarities=True)).load(c.struc_patches()) is None


def test_readers_and_component_order_match_the_reference_caches():
    """components must come out in the reference's order: the (n_sub, C, N) similarity cache identifies the members of
    component c as the columns holding hop 0 (SubGNN.py:771 stores the raw min hop count; the tiny graph is connected)."""
    from subgnn_b200 import formats, prepare
    from subgnn_b200.graph import DeviceGraph
    edges = formats.read_edge_list(TASK / 'edge_list.txt')
    emb = formats.load_embeddings(TASK / 'gin_embeddings.pth')
    n = emb.shape[0]
    hop = formats.load_hop_table(TASK / 'shortest_path_matrix.npy', n)
    assert hop.dtype == np.uint8 and np.array_equal(hop, np.load(TASK / 'shortest_path_matrix.npy').astype(np.uint8))
    g = DeviceGraph.from_edges(n, edges, device='cpu', one_indexed=False)
    deg = formats.load_degree_dict(TASK / 'degree_sequence.txt')
    assert np.array_equal(deg, g.rowptr_host[1:] - g.rowptr_host[:-1])
    ego = json.loads((TASK / 'ego_graphs.txt').read_text())
    assert all(sorted(ego[str(i)]) == g.col_host[g.rowptr_host[i]:g.rowptr_host[i + 1]].tolist() for i in range(n))
    splits, multilabel, n_labels = formats.read_subgraphs(TASK / 'subgraphs.pth')
    assert not multilabel and n_labels == 3
    assert [len(splits[s][0]) for s in ('train', 'val', 'test')] == [25, 6, 5]       # val / test swapped: val is the larger one
    cache = formats.SimilarityCache(TASK / 'similarities', _hp())
    for s in ('train', 'val', 'test'):
        subs = [[x + 1 for x in nodes] for nodes in splits[s][0]]
        cc = prepare.initialize_cc_ids(g, subs)
        sim = np.load(TASK / 'similarities' / cache.np_sim(s))
        assert sim.shape == cc.shape[:2] + (n,)
        for i in range(cc.shape[0]):
            for c in range(cc.shape[1]):
                members = sorted(int(x) for x in cc[i, c] if x)
                ref_members = (np.nonzero(sim[i, c] == 0.0)[0] + 1).tolist() if members else []
                assert members == ref_members, (s, i, c)


def test_multilabel_subgraph_file_and_writers_round_trip(tmp_path):
    from subgnn_b200 import formats
    subs = [[0, 1, 2], [3], [4, 5], [6, 7, 8, 9], [1, 3]]
    labels = [['x', 'y'], ['y'], ['z', 'x'], ['z'], ['y', 'z']]
    formats.write_subgraphs(tmp_path / 'subgraphs.pth', subs, labels, ['train', 'train', 'val', 'test', 'val'])
    out, multilabel, n_labels = formats.read_subgraphs(tmp_path / 'subgraphs.pth')
    assert multilabel and n_labels == 3
    assert out['train'] == ([[0, 1, 2], [3]], [[0, 1], [1]]) and out['val'][1] == [[2, 0], [1, 2]] and out['test'] == ([[6, 7, 8, 9]], [[2]])
    formats.write_edge_list(tmp_path / 'e.txt', [(0, 1), (1, 2)])
    assert formats.read_edge_list(tmp_path / 'e.txt').tolist() == [[0, 1], [1, 2]]
    formats.write_embeddings(tmp_path / 'emb.pth', np.arange(6, dtype=np.float32).reshape(3, 2))
    assert formats.load_embeddings(tmp_path / 'emb.pth').tolist() == [[0, 1], [2, 3], [4, 5]]
    padded = formats.pad_ragged([0, 2, 2, 5], [7, 8, 1, 2, 3], (3,))
    assert padded.tolist() == [[7, 8, 0], [0, 0, 0], [1, 2, 3]]
    from oracle import ref_loader
    if ref_loader.available():
        ref = ref_loader.load()
        r = ref.subgraph_utils.read_subgraphs(tmp_path / 'subgraphs.pth')
        assert r[0] == out['train'][0] and r[1] == out['train'][1] and r[2] == out['val'][0] and r[3] == out['val'][1]
        r = ref.subgraph_utils.read_subgraphs(TASK / 'subgraphs.pth')
        mine = formats.read_subgraphs(TASK / 'subgraphs.pth')[0]
        assert r[0] == mine['train'][0] and r[1].tolist() == [l[0] for l in mine['train'][1]] and r[2] == mine['val'][0] and r[4] == mine['test'][0]


# ---- 8f-4: epoch-end metrics --------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', ['multiclass', 'binary', 'multilabel'])
def test_epoch_end_metrics_match_the_reference(case):
    from subgnn_b200 import SubGNN as sg
    gold = json.loads((GOLDEN / 'metrics_golden.json').read_text())[case]
    logits, labels = torch.tensor(gold['logits'], dtype=torch.float32), torch.tensor(gold['labels'])
    ml = gold['multilabel']
    for kind in ('val', 'test'):
        outs = []
        for ch_l, ch_y in zip(torch.chunk(logits, 3), torch.chunk(labels, 3)):
            loss = (torch.nn.functional.binary_cross_entropy_with_logits(ch_l, ch_y.float()) if ml else torch.nn.functional.cross_entropy(ch_l, ch_y))
            outs.append({kind + '_loss': loss, kind + '_acc': sg.calc_accuracy(ch_l, ch_y, ml), kind + '_macro_f1': sg.calc_f1(ch_l, ch_y, 'macro', ml),
                         kind + '_logits': ch_l, kind + '_labels': ch_y})
        holder = type('H', (), {'multilabel': ml, 'hparams': {}, 'metric_scores': []})()
        holder._epoch_end = lambda o, k, h=holder: sg.SubGNN._epoch_end(h, o, k)
        res = (sg.SubGNN.validation_epoch_end if kind == 'val' else sg.SubGNN.test_epoch_end)(holder, outs)
        got = {k: float(v) for k, v in res['log'].items()}
        assert set(got) == set(gold[kind])
        for k, v in gold[kind].items():
            assert got[k] == pytest.approx(v, rel=1e-6, abs=1e-7), k
        if kind == 'val':
            assert holder.metric_scores[-1]['val_micro_f1'].numpy() == pytest.approx(gold['val']['val_micro_f1'])   # train_config.py:203
        else:
            assert float(holder.test_results['test_auroc']) == pytest.approx(gold['test']['test_auroc'])


# ---- 8f-1: driver stand-ins ---------------------------------------------------------------------------------------------
def test_commentjson_lite():
    from subgnn_b200.compat import commentjson_lite as cj
    txt = '{\n "a": 1, // one\n "b": "x // not a comment # nor this", # two\n /* block\n comment */ "c": [1e-4, true, null]\n}'
    assert cj.loads(txt) == {'a': 1, 'b': 'x // not a comment # nor this', 'c': [1e-4, True, None]}


def test_optuna_lite_study(tmp_path):
    from subgnn_b200.compat import optuna_lite as op
    db = 'sqlite:///' + str(tmp_path / 'study' / 'db.sqlite')
    study = op.create_study(direction='maximize', sampler=op.GridSampler({'k': [1, 2, 3]}), storage=db, study_name='s', load_if_exists=True)
    seen = []

    def objective(trial):
        k = trial.suggest_int('k', 1, 3)
        lr = trial.suggest_float('lr', 1e-4, 1e-2, log=True)
        assert trial.suggest_float('lr', 0, 1) == lr                          # same name -> same value
        c = trial.suggest_categorical('c', ['sum', 'max'])
        assert 1e-4 <= lr <= 1e-2 and c in ('sum', 'max') and trial.suggest_int('n', 5, 9) in range(5, 10)
        seen.append(k)
        return float(k)

    study.optimize(objective, n_trials=3, n_jobs=4)
    assert sorted(seen) == [1, 2, 3] and study.best_value == 3.0 and study.best_params['k'] == 3
    again = op.create_study(direction='maximize', storage=db, study_name='s', load_if_exists=True)
    assert len(again.trials) == 3 and again.best_params == study.best_params
    import joblib
    joblib.dump(study, tmp_path / 'study.pkl')
    assert joblib.load(tmp_path / 'study.pkl').best_value == 3.0
    # median pruner through the Lightning callback
    st = op.create_study(direction='maximize', pruner=op.MedianPruner(n_startup_trials=2))

    class T:
        callback_metrics, current_epoch = {}, 0

    def obj(trial, curve):
        cb = op.PyTorchLightningPruningCallback(trial, 'val_micro_f1')
        for e, v in enumerate(curve):
            T.callback_metrics, T.current_epoch = {'val_micro_f1': torch.tensor(v)}, e
            cb.on_epoch_end(T, None)
        return curve[-1]

    st.optimize(lambda t: obj(t, [0.5, 0.6, 0.7]), n_trials=1)
    st.optimize(lambda t: obj(t, [0.5, 0.7, 0.8]), n_trials=1)
    st.optimize(lambda t: obj(t, [0.1, 0.1, 0.1]), n_trials=1)
    assert [t.state for t in st.trials] == ['COMPLETE', 'COMPLETE', 'PRUNED']


def test_trainer_protocol_with_a_toy_module(tmp_path):
    """hook order, gradient clipping, ModelCheckpoint top-k and the checkpoint layout train.py:307-316 restores from."""
    from subgnn_b200.compat import lightning as pl
    calls = []

    class Toy(pl.LightningModule):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Linear(3, 2)
            self.metric_scores = []

        def prepare_data(self):
            calls.append('prepare_data')
            g = torch.Generator().manual_seed(0)
            self.x, self.y = torch.randn(40, 3, generator=g), torch.randint(2, (40,), generator=g)

        def configure_optimizers(self):
            calls.append('configure_optimizers')
            return torch.optim.Adam(self.parameters(), lr=1e-2)

        def _loader(self):
            return [{'x': self.x[i:i + 10], 'y': self.y[i:i + 10]} for i in range(0, 40, 10)]

        train_dataloader = val_dataloader = test_dataloader = _loader

        def training_step(self, b, i):
            return {'loss': torch.nn.functional.cross_entropy(self.w(b['x']), b['y']), 'log': {}}

        def backward(self, trainer, loss, optimizer, idx):
            calls.append('backward')
            loss.backward()

        def validation_step(self, b, i):
            return {'val_loss': torch.nn.functional.cross_entropy(self.w(b['x']), b['y'])}

        def validation_epoch_end(self, outs):
            v = torch.stack([o['val_loss'] for o in outs]).mean()
            self.metric_scores.append({'val_loss': v})
            return {'avg_val_loss': v, 'log': {'val_loss': v, 'val_acc': torch.tensor(0.5)}}

        test_step = validation_step

        def test_epoch_end(self, outs):
            return {'log': {'test_loss': torch.stack([o['val_loss'] for o in outs]).mean()}}

    logger = pl.TensorBoardLogger(str(tmp_path), name='run', version='version_7')
    ck = pl.ModelCheckpoint(filepath=os.path.join(logger.log_dir, '{epoch}-{val_loss:.2f}-{val_acc:.2f}'), save_top_k=2, monitor='val_loss', mode='min')
    m = Toy()
    tr = pl.Trainer(max_epochs=5, gpus=0, gradient_clip_val=0.1, logger=logger, checkpoint_callback=ck, progress_bar_refresh_rate=0)
    tr.fit(m)
    assert calls[:2] == ['prepare_data', 'configure_optimizers'] and calls.count('backward') == 20
    files = sorted(f for f in os.listdir(logger.log_dir) if f.endswith('.ckpt'))
    assert len(files) == 2 and all(f.startswith('epoch=') and 'val_loss=' in f for f in files)
    ckpt = torch.load(os.path.join(logger.log_dir, files[-1]))
    assert set(ckpt['state_dict']) == set(m.state_dict()) and ckpt['epoch'] >= 1
    losses = [float(s['val_loss']) for s in m.metric_scores]
    assert losses[-1] < losses[0]
    rows = [json.loads(l) for l in open(os.path.join(logger.log_dir, 'metrics.jsonl'))]
    assert any('val_loss' in r for r in rows)
    assert 'test_loss' in tr.test(m)['log']


def _write_run_config(path, task, tb_dir, n_trials=2):
    cfg = {
        'data': {'task': task, 'embedding_type': 'gin'},
        'tb': {'tb_logging': True, 'dir': str(tb_dir), 'name': 'tiny_optuna', 'local': True},
        'no_gpu': True,
        'optuna': {'opt_n_trials': n_trials, 'opt_n_cores': 1, 'monitor_metric': 'val_micro_f1', 'opt_direction': 'maximize', 'sampler': 'random',
                   'pruning': True},
        'hyperparams_fix': {k: v for k, v in _hp().items() if k not in ('learning_rate', 'lin_dropout', 'cc_aggregator')},
        'hyperparams_optuna': {'learning_rate': {'type': 'suggest_float', 'args': [1e-3, 1e-2], 'kwargs': {'log': True}},
                               'lin_dropout': {'type': 'suggest_float', 'args': [0.0, 0.2]},
                               'cc_aggregator': {'type': 'suggest_categorical', 'args': [['sum', 'max']]}},
    }
    txt = json.dumps(cfg, indent=1).replace('"optuna": {', '"optuna": {   // search set-up (commentjson)\n')
    Path(path).write_text('# run config for the tiny task\n' + txt)


def test_reference_train_config_runs_unchanged_on_the_stand_ins(tmp_path, monkeypatch):
    """The unmodified /root/reference/SubGNN/train_config.py (optuna study -> pl.Trainer.fit -> metric_scores) runs on
    subgnn_b200.compat, here with the unmodified reference model on the CPU (the CUDA-backed model takes its place on
    the GPU box: tests/test_gpu_task_dir.py)."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('/root/reference not present')
    ref = ref_loader.load(project_root=tmp_path)
    from subgnn_b200 import compat
    assert set(compat.install(force=True)) == {'pytorch_lightning', 'optuna', 'commentjson'}
    import pytorch_lightning as pl
    pl.LightningModule = torch.nn.Module                     # the reference module was already imported over the bare stub
    shutil.copytree(TASK, tmp_path / 'task_tiny')
    _write_run_config(tmp_path / 'cfg.json', 'task_tiny', tmp_path / 'tb')
    monkeypatch.setattr(sys, 'argv', ['train_config.py', '-config_path', str(tmp_path / 'cfg.json')])
    monkeypatch.chdir(tmp_path)
    # F15 (torch >= 2): the reference's .view on the transposed N anchors needs contiguity — harness-side shim, as in make_golden.py
    aps = ref.anchor_patch_samplers
    orig = aps.init_anchors_neighborhood

    def contiguous_anchors(*a, **k):
        out = orig(*a, **k)
        for d in out:
            for s in d:
                for l in d[s]:
                    d[s][l] = d[s][l].contiguous()
        return out

    monkeypatch.setattr(ref.SubGNN, 'init_anchors_neighborhood', contiguous_anchors)
    ns = runpy.run_path(str(ref_loader.REF_ROOT / 'SubGNN' / 'train_config.py'), run_name='__main__')
    study_dir = tmp_path / 'tb' / 'tiny_optuna'
    assert (study_dir / 'optuna_study_sqlite.db').exists() and (study_dir / 'optuna_study.pkl').exists()
    versions = [d for d in study_dir.iterdir() if d.is_dir()]
    assert len(versions) == 2
    for v in versions:
        assert json.loads((v / 'hyperparams.json').read_text())['cc_aggregator'] in ('sum', 'max')
        scores = json.loads((v / 'final_metric_scores.json').read_text())
        assert {'val_micro_f1', 'val_acc', 'val_auroc', 'val_loss'} <= set(scores)
        assert any(f.name.startswith('epoch=') and f.name.endswith('.ckpt') for f in v.iterdir())
    assert 'train_model' in ns
