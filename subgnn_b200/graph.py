"""Device graph store: CSR base graph (+ optional uint8 hop table) resident in HBM.

Replaces the networkx adjacency dicts the reference walks (SubGNN.py:525, 555-556) and the fp64
``shortest_path_matrix.npy`` (SubGNN.py:848).  Node ids keep the reference convention at every API
(1-indexed, 0 = PAD); CSR arrays are 0-indexed with sorted neighbour lists (int32).
"""
import weakref

import numpy as np
import torch


class DeviceGraph:
    def __init__(self, n_nodes, rowptr, col, device='cuda'):
        self.n_nodes = int(n_nodes)
        self.device = torch.device(device)
        self.rowptr_host = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.col_host = np.ascontiguousarray(col, dtype=np.int32)
        self.rowptr = torch.from_numpy(self.rowptr_host).to(self.device)
        self.col = torch.from_numpy(self.col_host).to(self.device)
        self.hop = None            # uint8 (N, N) on device, 0 = self / unreachable
        self._all_nodes = None

    # ---- constructors -------------------------------------------------------------------------
    @classmethod
    def from_edges(cls, n_nodes, edges, device='cuda', one_indexed=True):
        """edges: (E, 2) array of undirected edges; ids 1..N if one_indexed else 0..N-1."""
        e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        if one_indexed:
            e = e - 1
        both = np.concatenate([e, e[:, ::-1]], axis=0)
        key = np.unique(both[:, 0] * n_nodes + both[:, 1])          # sorts by (row, col) and de-duplicates
        rows, cols = key // n_nodes, key % n_nodes
        rowptr = np.zeros(n_nodes + 1, dtype=np.int64)
        np.add.at(rowptr, rows + 1, 1)
        rowptr = np.cumsum(rowptr)
        return cls(n_nodes, rowptr, cols, device)

    @classmethod
    def from_networkx(cls, G, device='cuda'):
        """G: networkx graph whose nodes are the 1-indexed ints 1..N (SubGNN.py:555-556)."""
        n = G.number_of_nodes()
        edges = np.array([(int(u), int(v)) for u, v in G.edges()], dtype=np.int64).reshape(-1, 2)
        assert n == 0 or (edges.size == 0 or (edges.min() >= 1 and edges.max() <= n)), 'nodes must be 1..N'
        return cls.from_edges(n, edges, device)

    # ---- accessors ----------------------------------------------------------------------------
    def set_hop_table(self, hop):
        """hop: (N, N) array-like of hop counts (any numeric dtype; stored uint8)."""
        h = np.asarray(hop)
        assert h.shape == (self.n_nodes, self.n_nodes)
        assert h.max() < 256, 'hop counts must fit uint8'
        self.hop = torch.from_numpy(np.ascontiguousarray(h.astype(np.uint8))).to(self.device)
        return self

    def all_nodes(self):
        if self._all_nodes is None:
            self._all_nodes = torch.arange(1, self.n_nodes + 1, dtype=torch.int32, device=self.device)
        return self._all_nodes

    def degrees(self):
        return (self.rowptr[1:] - self.rowptr[:-1])


_cache = weakref.WeakKeyDictionary()


def resolve_graph(g, device='cuda'):
    """Accepts a DeviceGraph or a networkx graph (reference call sites pass ``networkx_graph``) and
    returns the cached device handle."""
    if isinstance(g, DeviceGraph):
        return g
    try:
        return _cache[g]
    except KeyError:
        dg = DeviceGraph.from_networkx(g, device)
        _cache[g] = dg
        return dg


def ragged_from_padded(rows):
    """(n, L) int tensor/array with PAD=0 (left-packed or not) -> (ptr int32 [n+1], items int32) on CPU numpy."""
    r = np.asarray(rows)
    mask = r != 0
    lens = mask.sum(axis=1)
    ptr = np.zeros(r.shape[0] + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    return ptr.astype(np.int32), r[mask].astype(np.int32)
