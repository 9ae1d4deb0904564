"""Loads the UNMODIFIED reference modules from /root/reference under sys.modules stubs.

TEST INFRASTRUCTURE, BUILD-CONTAINER ONLY: /root/reference does not exist on the GPU box, so
nothing in ``-m gpu`` tests, smoke() or bench.py imports this file.  It is used by
tests/golden/make_golden.py (to generate the committed golden vectors) and by the ``not gpu``
pin tests, which skip when the reference is absent.

Stubs: matplotlib.pyplot, pytorch_lightning.LightningModule (= nn.Module),
torch_geometric.{nn.MessagePassing, nn.GINConv, utils.add_self_loops, utils.convert.to_networkx}
with torch_geometric 1.6.1's propagate plumbing, fastdtw (= oracle.gamma.fastdtw, the restated
third-party algorithm), plus the np.asmatrix shim for subgraph_utils.get_border_nodes (F11).
"""
import importlib
import inspect
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

REF_ROOT = Path('/root/reference')


def available():
    return (REF_ROOT / 'SubGNN' / 'SubGNN.py').exists()


class _Inspector:
    def __init__(self, owner):
        self.owner = owner

    def distribute(self, fn_name, coll):
        params = list(inspect.signature(getattr(self.owner, fn_name)).parameters)
        return {k: coll[k] for k in params[1:] if k in coll} if fn_name != 'message' else {k: coll[k] for k in params if k in coll}


class _MessagePassing(nn.Module):
    """torch_geometric 1.6.1 MessagePassing surface used by subgraph_mpn.py:176-224."""

    def __init__(self, aggr='add', flow='source_to_target', node_dim=0):
        super().__init__()
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self.inspector = _Inspector(self)
        self.__user_args__ = {'x_j', 'similarity', 'x'}

    def __check_input__(self, edge_index, size):
        return [None, None]

    def __collect__(self, args, edge_index, size, kwargs):
        x = kwargs['x']
        out = dict(kwargs)
        out['x_j'] = x.index_select(0, edge_index[0])
        out['index'] = edge_index[1]
        out['dim_size'] = x.size(0)
        out['ptr'] = None
        return out

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        out = torch.zeros((dim_size,) + tuple(inputs.shape[1:]), dtype=inputs.dtype, device=inputs.device)
        return out.index_add_(0, index, inputs)


def _install_stubs():
    import oracle.gamma as og

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if 'matplotlib' not in sys.modules:
        mod('matplotlib', pyplot=mod('matplotlib.pyplot'))
    mod('pytorch_lightning', LightningModule=nn.Module)
    tg = mod('torch_geometric')
    tg.nn = mod('torch_geometric.nn', MessagePassing=_MessagePassing, GINConv=object)
    tg.utils = mod('torch_geometric.utils', add_self_loops=lambda *a, **k: None)
    tg.utils.convert = mod('torch_geometric.utils.convert', to_networkx=lambda *a, **k: None)
    def _fastdtw(x, y, radius=1, dist=None):
        # the reference also calls fastdtw for padded (empty) components and then overwrites those
        # entries with 0 (SubGNN.py:807-815, :831); the value returned here is never used.
        if len(x) == 0 or len(y) == 0:
            return 0.0, []
        return og.fastdtw(x, y, radius=radius, dist=dist)

    mod('fastdtw', fastdtw=_fastdtw)


_loaded = {}


def load(project_root=None):
    """Returns a namespace with the reference modules: anchor_patch_samplers, gamma,
    subgraph_mpn, subgraph_utils, datasets, SubGNN, config."""
    if 'ns' in _loaded:
        if project_root is not None:
            _loaded['ns'].config.PROJECT_ROOT = Path(project_root)
        return _loaded['ns']
    assert available(), '/root/reference is not present (GPU box?)'
    _install_stubs()
    sys.path.insert(0, str(REF_ROOT))
    sys.path.insert(0, str(REF_ROOT / 'SubGNN'))
    os.environ.setdefault('CUDA_VISIBLE_DEVICES', '')
    ns = types.SimpleNamespace()
    for name in ('config', 'subgraph_utils', 'anchor_patch_samplers', 'gamma', 'subgraph_mpn', 'datasets', 'SubGNN'):
        setattr(ns, name, importlib.import_module(name))
    if project_root is not None:
        ns.config.PROJECT_ROOT = Path(project_root)
    # F11: scipy sparse *arrays* return ndarray from .todense(); get_border_nodes relies on np.matrix semantics
    import networkx as nx
    _orig_adj = nx.adjacency_matrix

    class _AsMatrix:
        def __init__(self, a):
            self.a = a

        def todense(self):
            return np.asmatrix(self.a.todense())

    ns.subgraph_utils.nx = types.SimpleNamespace(**{k: getattr(nx, k) for k in dir(nx) if not k.startswith('_')})
    ns.subgraph_utils.nx.adjacency_matrix = lambda g: _AsMatrix(_orig_adj(g))
    _loaded['ns'] = ns
    return ns
