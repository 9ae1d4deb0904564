"""SubGNN module with the reference's class / method / state_dict names (SubGNN/SubGNN.py:90-1164), backed by
the CUDA step engine.

Drop-in surface kept from the reference LightningModule (pytorch-lightning 0.7.1 protocol — the module is a
plain nn.Module so it does not need Lightning to be installed):
    SubGNN(hparams, graph_path, subgraph_path, embedding_path, similarities_path, shortest_paths_path,
           degree_dict_path, ego_graph_path)            SubGNN.py:94-96
    prepare_data() / prepare_test_data()                 :1024 / :994     (GPU kernels instead of Python loops)
    forward(dataset_type, ..., subgraph_idx, ...)        :225
    training_step / validation_step / test_step          :317 / :393 / :399
    configure_optimizers / backward                      :1156 / :1163
    train_dataloader / val_dataloader / test_dataloader  :1116-1151   (batches carry indices + labels only)
plus ``training_step_fused`` — the whole optimisation step (forward, loss, backward, clip, Adam) as one captured
CUDA graph, which is what bench.py measures.
"""
import math
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from . import PAD_VALUE, _abi
from ._abi import call, ptr
from .engine import Engine
from .graph import DeviceGraph

PROJECT_ROOT = Path('.')      # config.py:5 equivalent; set subgnn_b200.SubGNN.PROJECT_ROOT like config.PROJECT_ROOT


def read_subgraphs(sub_f):
    """subgraph_utils.py:24-92 — 'n1-n2-..\\tlabel[-label]\\tsplit' lines -> train/val/test node lists and labels."""
    labels, out = {}, {'train': ([], []), 'val': ([], []), 'test': ([], [])}
    multilabel = False
    with open(sub_f) as fin:
        for line in fin:
            parts = line.rstrip('\n').split('\t')
            nodes = [int(n) for n in parts[0].split('-') if n != '']
            if not nodes:
                continue
            labs = parts[1].split('-')
            multilabel |= len(labs) > 1
            for lab in labs:
                labels.setdefault(lab, len(labels))
            split = parts[2].strip()
            if split in out:
                out[split][0].append(nodes)
                out[split][1].append([labels[l] for l in labs])
    if len(out['val'][0]) < len(out['test'][0]):                       # subgraph_utils.py:89-90
        out['val'], out['test'] = out['test'], out['val']
    return out, multilabel, len(labels)


class _LSTMFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        runner, arena = module._runner_for(x.shape[0], x.shape[1])
        module._pack(arena, params)
        xc = x.contiguous().float()
        st = _abi.stream_ptr()
        call('subgnn_inc_step', ptr(module._step), st)
        runner.forward(None, module.training, module._seed, ptr(module._step), st, dense_x=xc.view(-1, x.shape[2]))
        ctx.module, ctx.runner, ctx.arena, ctx.training = module, runner, arena, module.training
        ctx.x_shape = x.shape
        return runner.Y.clone()

    @staticmethod
    def backward(ctx, dy):
        m, runner, arena = ctx.module, ctx.runner, ctx.arena
        st = _abi.stream_ptr()
        call('subgnn_fill_zero', ptr(arena.grads), arena.size, st)
        runner.dEMB.copy_(dy.contiguous())
        dx = torch.empty(ctx.x_shape, dtype=torch.float32, device=dy.device)
        runner.backward(None, None, ctx.training, m._seed, ptr(m._step), st, dense_dx=dx.view(-1, ctx.x_shape[2]))
        return (None, dx) + tuple(arena.view(n, 'grads').clone() for n in m._arena_names)


class LSTM(nn.Module):
    """SubGNN.py:60-88 — bidirectional LSTM walk encoder with a linear head, same constructor and parameter names
    (``lstm.weight_ih_l0`` ... ``linear.weight``); the computation runs on csrc/lstm.cu + csrc/gemm.cu."""

    def __init__(self, n_features, h, dropout=0.0, num_layers=1, batch_first=True, aggregator='last'):
        super().__init__()
        if h != n_features:
            raise NotImplementedError('the walk encoder is built for h == n_features (SubGNN.py:175)')
        if aggregator not in ('last', 'sum'):
            raise NotImplementedError                                                              # SubGNN.py:86-87
        self.num_layers, self.aggregator, self.dropout, self.h = num_layers, aggregator, dropout, h
        self.lstm = nn.LSTM(n_features, h, num_layers=num_layers, batch_first=batch_first, dropout=dropout, bidirectional=True)   # parameters only
        self.linear = nn.Linear(h * 2, n_features)
        self._runners = {}
        self._seed = 0x5eed
        self._step = None
        self._arena_names = []
        for k in range(num_layers):
            for nm in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh'):
                self._arena_names += ['lstm.lstm.%s_l%d' % (nm, k), 'lstm.lstm.%s_l%d_reverse' % (nm, k)]
        self._arena_names += ['lstm.linear.weight', 'lstm.linear.bias']

    def _module_params(self):
        return [self.get_parameter(n[len('lstm.'):]) for n in self._arena_names]

    def _runner_for(self, n_seq, T):
        from .engine import LstmRunner, ParamArena
        # a fresh runner per call: its buffers hold the activations saved for this call's backward (the reference calls the
        # shared LSTM several times per step — internal / border side of every layer — before loss.backward())
        key = (n_seq, T, len(self._runners))
        if key not in self._runners:
            self._runners.clear()
            dev = self.linear.weight.device
            hp = {'node_embed_size': self.h, 'n_layers': 1, 'freeze_node_embeds': True, 'use_neighborhood': False, 'use_position': False,
                  'use_structure': False, 'lstm_n_layers': self.num_layers, 'linear_hidden_dim_1': 1, 'linear_hidden_dim_2': 1, 'trainable_cc': False,
                  'n_triangular_walks': 1, 'lstm_aggregator': self.aggregator, 'lstm_dropout': self.dropout}
            arena = ParamArena(hp, 0, 1, 1, device=dev)
            self._runners[key] = (LstmRunner(arena, hp, None, n_seq, dev, n_seq=n_seq, T=T), arena)
            if self._step is None:
                self._step = torch.zeros(1, dtype=torch.int32, device=dev)
        return self._runners[key]

    def _pack(self, arena, params):
        for name, p in zip(self._arena_names, params):
            arena.view(name).copy_(p.detach())

    def forward(self, input):
        return _LSTMFunction.apply(self, input, *self._module_params())


class _EngineFunction(torch.autograd.Function):
    """logits = engine(split, indices); backward feeds d logits to the CUDA backward and returns the parameter
    gradients (views of the gradient arena are cloned out so torch can accumulate them)."""

    @staticmethod
    def forward(ctx, module, split, indices, training, *params):
        eng = module.engine
        logits, _ = eng.forward(split, indices, training=training)
        ctx.module, ctx.split, ctx.B = module, split, len(indices)
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        m = ctx.module
        eng = m.engine
        eng.backward(ctx.split, ctx.B, training=True, external_dlogits=dlogits.contiguous())
        grads = [eng.arena.view(name, 'grads').clone() for name in m._param_names]
        return (None, None, None, None) + tuple(grads)


class SubGNN(nn.Module):
    def __init__(self, hparams, graph_path=None, subgraph_path=None, embedding_path=None, similarities_path=None,
                 shortest_paths_path=None, degree_dict_path=None, ego_graph_path=None, device=None):
        super().__init__()
        self.device_ = torch.device(device or ('cuda' if torch.cuda.is_available() else 'cpu'))
        if self.device_.type != 'cuda':
            raise RuntimeError('subgnn_b200.SubGNN needs a CUDA device: there is no CPU path')
        self.hparams = hparams
        self.graph_path, self.subgraph_path, self.embedding_path = graph_path, subgraph_path, embedding_path
        self.similarities_path, self.shortest_paths_path = similarities_path, shortest_paths_path
        self.degree_dict_path, self.ego_graph_path = degree_dict_path, ego_graph_path
        if 'structure_similarity_fn' not in self.hparams:
            self.hparams['structure_similarity_fn'] = 'dtw'                                    # SubGNN.py:186-187
        self.engine = None
        self.metric_scores = []
        self._param_names = []
        if graph_path is not None:
            self.read_data()

    # ---- construction helpers ----------------------------------------------------------------------
    @classmethod
    def from_engine(cls, engine):
        self = cls(dict(engine.hp), device=engine.device)
        self._attach(engine)
        return self

    @classmethod
    def from_prepared(cls, hparams, prepared, graph=None, device='cuda', seed=0, init_seed=None):
        self = cls(dict(hparams), device=device)
        eng = Engine(self.hparams, prepared, device=device, graph=graph, seed=seed)
        if init_seed is not None:
            eng.init_parameters(init_seed)
        self._attach(eng)
        return self

    def _attach(self, engine):
        """expose the arena as nn.Parameters under the reference's names (state_dict compatible)."""
        self.engine = engine
        self.num_classes = engine.num_classes
        self.multilabel = engine.tables['train'].multilabel
        self._param_names = list(engine.arena.entries)
        for name in self._param_names:
            p = nn.Parameter(engine.arena.view(name), requires_grad=True)
            self._register_by_path(name, p)

    def _register_by_path(self, path, param):
        mod = self
        parts = path.split('.')
        for part in parts[:-1]:
            if not hasattr(mod, part):
                setattr(mod, part, nn.Module())
            mod = getattr(mod, part)
        mod.register_parameter(parts[-1], param)

    # ---- data (SubGNN.py:519-570) --------------------------------------------------------------------
    def read_data(self):
        root = PROJECT_ROOT
        edges = np.loadtxt(root / self.graph_path, dtype=np.int64, usecols=(0, 1)).reshape(-1, 2)
        n_nodes = int(edges.max()) + 1
        emb = torch.load(root / self.embedding_path, map_location='cpu')
        n_nodes = max(n_nodes, emb.shape[0])
        self.graph = DeviceGraph.from_edges(n_nodes, edges, device=self.device_, one_indexed=False)    # ids become 1-indexed (:555-559)
        splits, self.multilabel, n_labels = read_subgraphs(root / self.subgraph_path)
        if self.multilabel:
            raise NotImplementedError('multi-label files: use from_prepared with labels_multi')
        self.sub_G = {k: [[n + 1 for n in s] for s in v[0]] for k, v in splits.items()}
        self.sub_G_label = {k: np.array([l[0] for l in v[1]], dtype=np.int64) for k, v in splits.items()}
        if self.hparams.get('subset_data', False):                                             # :542-546
            B = self.hparams['batch_size']
            self.sub_G = {k: v[:B] for k, v in self.sub_G.items()}
            self.sub_G_label = {k: v[:B] for k, v in self.sub_G_label.items()}
        self.num_classes = int(max(v.max() for v in self.sub_G_label.values() if len(v))) + 1
        self.hparams['node_embed_size'] = emb.shape[1]
        self.embeddings = np.concatenate([np.zeros((1, emb.shape[1]), dtype=np.float32), emb.numpy().astype(np.float32)])   # :564-565
        sp = root / self.shortest_paths_path if self.shortest_paths_path else None
        if sp is not None and sp.exists() and (self.hparams['use_position'] or self.hparams['use_neighborhood']):
            self.graph.set_hop_table(np.load(sp, allow_pickle=True))

    def prepare_data(self, seed=None):
        """SubGNN.py:1024-1063 on the GPU (components, border sets, walks, DTW / SP similarities, anchors)."""
        from . import prepare as prep
        seed = self.hparams.get('seed', 0) if seed is None else seed
        splits = [s for s in ('train', 'val') if len(self.sub_G[s])]
        prepared = prep.prepare(self.hparams, self.graph, self.sub_G, self.sub_G_label, self.embeddings, seed=seed, splits=splits,
                                num_classes=self.num_classes)
        eng = Engine(self.hparams, prepared, device=self.device_, graph=self.graph, seed=seed)
        eng.init_parameters(seed)
        self.prepared = prepared
        self._attach(eng)

    # ---- forward / steps -----------------------------------------------------------------------------
    def forward(self, dataset_type, *unused, subgraph_idx=None, **kw):
        """SubGNN.py:225.  Only ``dataset_type`` and ``subgraph_idx`` are consumed: component ids, anchors and
        similarities of the split are already resident on the device (the reference's remaining positional
        arguments are accepted and ignored)."""
        if subgraph_idx is None:
            # reference positional order: (.., subgraph_ids, cc_ids, subgraph_idx, NP_sim, I_S_sim, B_S_sim)
            subgraph_idx = unused[8]
        idx = subgraph_idx.reshape(-1).cpu().numpy()
        params = [self.get_parameter(n) for n in self._param_names]
        return _EngineFunction.apply(self, dataset_type, idx, self.training, *params)

    def _loss(self, logits, labels):
        if self.multilabel:
            return nn.functional.binary_cross_entropy_with_logits(logits.squeeze(1), labels.type_as(logits))
        return nn.functional.cross_entropy(logits, labels)

    def training_step(self, train_batch, batch_idx=0):
        """SubGNN.py:317-348 (autograd path: usable with any torch optimizer / trainer loop)."""
        logits = self.forward('train', subgraph_idx=train_batch['subgraph_idx'])
        labels = train_batch['label'].to(logits.device).squeeze(-1)
        loss = self._loss(logits, labels)
        acc = (logits.argmax(dim=1) == labels).float().mean() if not self.multilabel else torch.tensor(0.0)
        return {'loss': loss, 'log': {'train_loss': loss, 'train_acc': acc}}

    def training_step_fused(self, train_batch, use_graph=True):
        """Whole optimisation step on the device (forward, loss, backward, clip_grad_norm_, Adam) — the fast path.
        ``train_batch['subgraph_idx']`` may be a host tensor; it is the only per-step input."""
        idx = train_batch['subgraph_idx'].reshape(-1).numpy() if not train_batch['subgraph_idx'].is_cuda else train_batch['subgraph_idx'].reshape(-1).cpu().numpy()
        loss = self.engine.train_step(idx, use_graph=use_graph)
        return {'loss': loss}

    def val_test_step(self, batch, batch_idx=0, is_test=False):
        split = 'test' if is_test else 'val'
        with torch.no_grad():
            idx = batch['subgraph_idx'].reshape(-1).cpu().numpy()
            logits, loss = self.engine.forward(split, idx, training=False)
        labels = batch['label'].to(logits.device).squeeze(-1)
        acc = (logits.argmax(dim=1) == labels).float().mean()
        k = 'test' if is_test else 'val'
        return {k + '_loss': loss.clone().squeeze(), k + '_acc': acc, k + '_logits': logits.clone(), k + '_labels': labels}

    def validation_step(self, val_batch, batch_idx=0):
        return self.val_test_step(val_batch, batch_idx, is_test=False)

    def test_step(self, test_batch, batch_idx=0):
        return self.val_test_step(test_batch, batch_idx, is_test=True)

    # ---- loaders (SubGNN.py:1116-1151): batches carry indices and labels only ---------------------------
    def _loader(self, split, shuffle):
        labels = torch.as_tensor(np.asarray(self.engine.prepared['labels'][split]))
        n, B = len(labels), self.hparams['batch_size']
        order = torch.randperm(n) if shuffle else torch.arange(n)
        drop_last = shuffle and B <= n
        for i in range(0, n - (B - 1 if drop_last else 0), B):
            idx = order[i:i + B]
            if len(idx):
                yield {'subgraph_idx': idx.view(-1, 1), 'label': labels[idx]}

    def train_dataloader(self):
        return self._loader('train', True)

    def val_dataloader(self):
        return self._loader('val', False)

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=self.hparams['learning_rate'])               # SubGNN.py:1156-1161

    def backward(self, trainer, loss, optimizer, optimizer_idx):
        loss.backward(retain_graph=True)                                                        # SubGNN.py:1163-1164
