"""GPU checks of SURVEY 8f rows 1, 2, 4 on a task directory in the reference's on-disk formats whose similarity caches
the UNMODIFIED reference wrote (tests/golden/task_tiny, tests/golden/make_task_golden.py):

  * reference caches load directly (file names, layouts, component order) and training runs from them;
  * every cached product recomputed by the CUDA kernels equals what the reference wrote (border sets as sets,
    min-hop similarities exactly, DTW similarities to fp32 rounding) and is written back under the same names;
  * a train_config.py-style driver (optuna study -> pl.Trainer.fit -> checkpoints -> Trainer.test) runs on the stand-ins
    with the CUDA-backed SubGNN module, through the fused CUDA-graph step and through the hook-by-hook autograd path.
"""
import json
import os
import shutil
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / 'golden'
TASK = GOLDEN / 'task_tiny'
PATHS = ('task_tiny/edge_list.txt', 'task_tiny/subgraphs.pth', 'task_tiny/gin_embeddings.pth', 'task_tiny/similarities/',
         'task_tiny/shortest_path_matrix.npy', 'task_tiny/degree_sequence.txt', 'task_tiny/ego_graphs.txt')


def _hp(**kw):
    hp = json.loads((TASK / 'hyperparams.json').read_text())
    hp.update(kw)
    return hp


def _model(root, **kw):
    from subgnn_b200 import SubGNN as sg
    sg.PROJECT_ROOT = Path(root)
    return sg.SubGNN(_hp(**kw), *PATHS)


def _copy_task(tmp_path, keep=lambda name: True):
    shutil.copytree(TASK, tmp_path / 'task_tiny')
    for p in (tmp_path / 'task_tiny' / 'similarities').iterdir():
        if not keep(p.name):
            p.unlink()
    return tmp_path


def test_reference_caches_load_directly_and_train(tmp_path):
    root = _copy_task(tmp_path)
    before = {p.name: p.stat().st_mtime_ns for p in (root / 'task_tiny' / 'similarities').iterdir()}
    m = _model(root)
    train_val = {n for n in before if '_test_' not in n}
    assert set(m.sim_cache.loaded) == train_val and m.sim_cache.saved == []
    ref_p = np.load(TASK / 'similarities' / m.sim_cache.struc_patches())
    assert np.array_equal(m.prepared['structure_anchors'], ref_p)
    assert np.array_equal(m.prepared['int_rw_all'], np.load(TASK / 'similarities' / m.sim_cache.walks(True)))
    assert np.array_equal(m.prepared['I_S_sim']['val'], np.load(TASK / 'similarities' / m.sim_cache.struc_sim(True, 'val')))
    # the engine's resolved similarities come from the reference's dense slab: same values as a hop-table lookup
    t = m.engine.tables['train']
    from subgnn_b200 import ops
    ids = t.n_ids[1][0]
    want = ops.sp_min_gather(m.graph.hop, t.cc_nodeptr, t.cc_nodes, ids.contiguous())
    assert torch.equal(t.n_sim[1][0], want)
    losses = []
    for epoch in range(3):
        for batch in m.train_dataloader():
            losses.append(float(m.training_step_fused(batch)['loss']))
    assert np.isfinite(losses).all() and np.mean(losses[-4:]) < np.mean(losses[:4])
    outs = [m.validation_step(b, i) for i, b in enumerate(m.val_dataloader())]
    res = m.validation_epoch_end(outs)
    assert {'val_loss', 'val_micro_f1', 'val_macro_f1', 'val_acc', 'val_auroc', 'avg_val_acc', 'avg_macro_f1', 'val_auroc_class_0'} <= set(res['log'])
    outs = [m.test_step(b, i) for i, b in enumerate(m.test_dataloader())]
    assert 'test_micro_f1' in m.test_epoch_end(outs)['log'] and set(m.sim_cache.loaded) == set(before)
    after = {p.name: p.stat().st_mtime_ns for p in (root / 'task_tiny' / 'similarities').iterdir()}
    assert after == before, 'existing caches must not be rewritten'


def _sets(arr):
    return [[frozenset(int(x) for x in row if x) for row in sub] for sub in arr]


def test_recomputed_products_equal_what_the_reference_wrote(tmp_path):
    """keep only the sampled patches + walks (stochastic, reference MT stream); recompute everything derived from them."""
    root = _copy_task(tmp_path, keep=lambda n: n.startswith('struc_patches') or 'random_walks' in n)
    m = _model(root)
    m.prepare_test_data()
    c = m.sim_cache
    sim_dir = root / 'task_tiny' / 'similarities'
    assert {p.name for p in sim_dir.iterdir()} == {p.name for p in (TASK / 'similarities').iterdir()}
    for s in ('train', 'val', 'test'):
        assert _sets(np.load(sim_dir / c.border_set(s))) == _sets(np.load(TASK / 'similarities' / c.border_set(s))), s
        mine, ref = np.load(sim_dir / c.np_sim(s)), np.load(TASK / 'similarities' / c.np_sim(s))
        assert mine.dtype == ref.dtype == np.float32 and np.array_equal(mine, ref), s
        for inside in (True, False):
            mine, ref = np.load(sim_dir / c.struc_sim(inside, s)), np.load(TASK / 'similarities' / c.struc_sim(inside, s))
            assert mine.shape == ref.shape and mine.dtype == ref.dtype
            np.testing.assert_allclose(mine, ref, rtol=1e-6, atol=1e-7)                  # fp64 DTW cast to fp32 on both sides


def test_full_pipeline_from_an_empty_cache_and_multilabel_labels(tmp_path):
    root = _copy_task(tmp_path, keep=lambda n: False)
    os.remove(root / 'task_tiny' / 'shortest_path_matrix.npy')                          # hop table then comes from the GPU BFS
    m = _model(root, compute_similarities=True)
    assert np.array_equal(m.graph.hop.cpu().numpy(), np.load(TASK / 'shortest_path_matrix.npy').astype(np.uint8))
    assert len(m.sim_cache.saved) == 11 and m.sim_cache.loaded == []
    from subgnn_b200 import formats
    formats.write_graph_metrics(root / 'task_tiny', m.graph)
    assert np.array_equal(np.load(root / 'task_tiny' / 'shortest_path_matrix.npy'), np.load(TASK / 'shortest_path_matrix.npy'))
    assert json.loads((root / 'task_tiny' / 'degree_sequence.txt').read_text()) == json.loads((TASK / 'degree_sequence.txt').read_text())
    assert json.loads((root / 'task_tiny' / 'ego_graphs.txt').read_text()) == json.loads((TASK / 'ego_graphs.txt').read_text())
    # multi-label variant of the same subgraphs (HPO-NEURO style file): BCE-with-logits path end to end
    splits, _, _ = formats.read_subgraphs(TASK / 'subgraphs.pth')
    rows = []
    rs = np.random.RandomState(0)
    for sp in ('train', 'val', 'test'):
        for nodes, lab in zip(*splits[sp]):
            labs = sorted({'c%d' % lab[0], 'c%d' % rs.randint(3)})
            rows.append((nodes, labs, sp))
    rows[0] = (rows[0][0], ['c0', 'c1', 'c2'], rows[0][2])
    formats.write_subgraphs(root / 'task_tiny' / 'subgraphs.pth', [r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows])
    m2 = _model(root, compute_similarities=True)
    assert m2.multilabel and m2.num_classes == 3 and m2.engine.prepared['labels']['train'].shape == (25, 3)
    l0 = [float(m2.training_step_fused(b)['loss']) for b in m2.train_dataloader()]
    for _ in range(6):
        l1 = [float(m2.training_step_fused(b)['loss']) for b in m2.train_dataloader()]
    assert np.isfinite(l1).all() and np.mean(l1) < np.mean(l0)
    out = m2.training_step(next(iter(m2.train_dataloader())))
    out['loss'].backward()
    res = m2.validation_epoch_end([m2.validation_step(b, i) for i, b in enumerate(m2.val_dataloader())])
    assert 0.0 <= float(res['log']['val_micro_f1']) <= 1.0 and 'val_auroc_class_2' in res['log']


def _driver(run_config, trial):
    """the call pattern of train_config.py:90-211 (build_model -> build_trainer -> fit -> metric_scores), written against
    the same module names the reference imports."""
    import optuna  # noqa: F401
    import pytorch_lightning as pl
    import SubGNN as md
    from pytorch_lightning.callbacks import ModelCheckpoint
    from pytorch_lightning.loggers import TensorBoardLogger
    from optuna.integration import PyTorchLightningPruningCallback
    hp = dict(run_config['hyperparams_fix'])
    hp.update({k: getattr(trial, v['type'])(k, *v['args'], **v.get('kwargs', {})) for k, v in run_config['hyperparams_optuna'].items()})
    torch.manual_seed(hp['seed'])
    model = md.SubGNN(hp, run_config['graph_path'], run_config['subgraphs_path'], run_config['embedding_path'], run_config['similarities_path'],
                      run_config['shortest_paths_path'], run_config['degree_sequence_path'], run_config['ego_graph_path'])
    logger = TensorBoardLogger(run_config['tb']['dir_full'], name=run_config['tb']['name'], version='version_%d' % trial.number)
    os.makedirs(logger.log_dir, exist_ok=True)
    kwargs = {'max_epochs': hp['max_epochs'], 'gpus': 1, 'num_sanity_val_steps': 0, 'progress_bar_refresh_rate': 0, 'gradient_clip_val': hp['grad_clip'],
              'logger': logger,
              'checkpoint_callback': ModelCheckpoint(filepath=os.path.join(logger.log_dir, '{epoch}-{val_micro_f1:.2f}-{val_acc:.2f}-{val_auroc:.2f}'),
                                                     save_top_k=2, verbose=False, monitor=run_config['optuna']['monitor_metric'], mode='max'),
              'early_stop_callback': PyTorchLightningPruningCallback(trial, monitor=run_config['optuna']['monitor_metric'])}
    kwargs.update(run_config.get('trainer_extra', {}))
    trainer = pl.Trainer(**kwargs)
    trainer.fit(model)
    run_config.setdefault('models', []).append((model, trainer, logger.log_dir))
    return float(np.max([score[run_config['optuna']['monitor_metric']].numpy() for score in model.metric_scores]))


@pytest.mark.parametrize('fused', [True, False])
def test_train_config_style_driver_on_the_stand_ins(tmp_path, fused):
    from subgnn_b200 import run_reference_script
    root = _copy_task(tmp_path)
    replaced = run_reference_script.setup(root)
    assert 'SubGNN' in sys.modules and sys.modules['config'].PROJECT_ROOT == Path(root)
    import optuna
    hp = _hp(max_epochs=4)
    run_config = {'graph_path': PATHS[0], 'subgraphs_path': PATHS[1], 'embedding_path': PATHS[2], 'similarities_path': PATHS[3],
                  'shortest_paths_path': PATHS[4], 'degree_sequence_path': PATHS[5], 'ego_graph_path': PATHS[6],
                  'tb': {'dir_full': str(tmp_path / 'tb'), 'name': 'study'}, 'optuna': {'monitor_metric': 'val_micro_f1'},
                  'hyperparams_fix': {k: v for k, v in hp.items() if k not in ('learning_rate', 'cc_aggregator')},
                  'hyperparams_optuna': {'learning_rate': {'type': 'suggest_float', 'args': [1e-3, 1e-2], 'kwargs': {'log': True}},
                                         'cc_aggregator': {'type': 'suggest_categorical', 'args': [['sum', 'max']]}},
                  'trainer_extra': {'fused': fused}}
    study = optuna.create_study(direction='maximize', sampler=optuna.samplers.RandomSampler(), pruner=optuna.pruners.MedianPruner(),
                                storage='sqlite:///' + str(tmp_path / 'tb' / 'study.db'), study_name='s', load_if_exists=True)
    study.optimize(lambda t: _driver(run_config, t), n_trials=2, n_jobs=1)
    assert len(study.trials) == 2 and 0.0 <= study.best_value <= 1.0 and set(study.best_params) == {'learning_rate', 'cc_aggregator'}
    for model, trainer, log_dir in run_config['models']:
        assert len(model.metric_scores) == 4 and trainer._fused_active == fused
        ckpts = [f for f in os.listdir(log_dir) if f.startswith('epoch=') and f.endswith('.ckpt')]
        assert 1 <= len(ckpts) <= 2
        ck = torch.load(os.path.join(log_dir, ckpts[0]))
        assert set(ck['state_dict']) == set(model.state_dict())
        assert ck['optimizer_states'] and len(ck['optimizer_states'][0]['state']) > 0          # Adam moments travel in both step modes
        # train.py:307-316 restore pattern, then the test pass (train.py:411-417)
        fresh = sys.modules['SubGNN'].SubGNN(dict(model.hparams), *PATHS)
        model_dict = fresh.state_dict()
        fresh.load_state_dict({k: v for k, v in ck['state_dict'].items() if k in model_dict})
        for k, v in ck['state_dict'].items():
            assert torch.equal(fresh.state_dict()[k].cpu(), v)
        trainer.test(fresh)
        assert {'test_micro_f1', 'test_acc', 'test_auroc'} <= set(fresh.test_results)
    rows = [json.loads(l) for l in open(os.path.join(run_config['models'][0][2], 'metrics.jsonl'))]
    assert sum('val_micro_f1' in r for r in rows) == 4


def test_resample_anchor_patches_each_epoch(tmp_path):
    """SubGNN.py:449-457: with resample_anchor_patches the N / P / S anchors are redrawn at validation_epoch_end (patches, walks,
    border sets and similarities are kept); parameters and optimizer state carry over and training continues."""
    root = _copy_task(tmp_path)
    m = _model(root, resample_anchor_patches=True)
    before = {k: np.array(v) for k, v in m.engine.prepared['anchors_neigh_int']['train'].items()}
    pos_before = np.array(m.engine.prepared['anchors_pos_ext'][0])
    sims_before = np.array(m.engine.prepared['I_S_sim']['train'])
    w_before = m.state_dict()['lin.weight'].clone()
    losses = []
    for epoch in range(3):
        losses.append(np.mean([float(m.training_step_fused(b)['loss']) for b in m.train_dataloader()]))
        res = m.validation_epoch_end([m.validation_step(b, i) for i, b in enumerate(m.val_dataloader())])
        assert np.isfinite(float(res['log']['val_loss']))
    after = m.engine.prepared['anchors_neigh_int']['train']
    assert any(not np.array_equal(before[l], after[l]) for l in before) and not np.array_equal(pos_before, m.engine.prepared['anchors_pos_ext'][0])
    assert np.array_equal(sims_before, m.engine.prepared['I_S_sim']['train'])
    assert not torch.equal(w_before, m.state_dict()['lin.weight']) and int(m.engine.step_dev.item()) == 3 * len(m.train_dataloader())
    assert losses[-1] < losses[0] and len(m.metric_scores) == 3


def test_fused_checkpoint_restores_adam_state_and_step_counter(tmp_path):
    """a checkpoint written in fused mode carries the engine's Adam moments and step count (torch.optim.Adam layout);
    restoring them continues training exactly, restoring the weights alone does not; frozen embeddings appear in the state_dict
    like the reference's nn.Embedding.from_pretrained(freeze=True) (SubGNN.py:568)."""
    root = _copy_task(tmp_path)
    m = _model(root, lin_dropout=0.2)
    torch.manual_seed(0)
    batches = list(m.train_dataloader())
    for b in batches:
        m.training_step_fused(b)
    torch.cuda.synchronize()
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    osd = m.engine.optimizer_state_dict()
    assert set(osd['state'][0]) >= {'step', 'exp_avg', 'exp_avg_sq'} and float(osd['state'][0]['step']) == len(batches)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt.load_state_dict({'state': osd['state'], 'param_groups': osd['param_groups']})       # loads into torch's own Adam as is
    for b in batches:
        m.training_step_fused(b)
    conts = {}
    for name, restore in (('warm', True), ('cold', False)):
        f = _model(root, lin_dropout=0.2)
        f.load_state_dict(sd)
        if restore:
            f.engine.load_optimizer_state_dict(osd)
        for b in batches:
            f.training_step_fused(b)
        torch.cuda.synchronize()
        conts[name] = {k: v.detach().cpu().numpy() for k, v in f.state_dict().items()}
    want = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    for k in want:
        np.testing.assert_allclose(conts['warm'][k], want[k], rtol=1e-5, atol=1e-7, err_msg=k)
    assert max(float(np.abs(conts['cold'][k] - want[k]).max()) for k in want) > 1e-4
    mf = _model(root, freeze_node_embeds=True)
    assert 'node_embeddings.weight' in mf.state_dict() and not dict(mf.named_parameters())['node_embeddings.weight'].requires_grad
