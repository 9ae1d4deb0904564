"""Oracle: triangular random walks and structure anchor-patch sampling.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  anchor_patch_samplers.py:20-47   is_triangle / get_neighbors
  anchor_patch_samplers.py:49-113  triangular_random_walk
  anchor_patch_samplers.py:118-158 perform_random_walks
  anchor_patch_samplers.py:210-243 sample_structure_anchor_patches
  subgraph_utils.py:126-144        get_border_nodes (sparse restatement, no dense N x N)

The walk *logic* is written once against two small interfaces:
  graph adapter : node order / neighbour order / edge test
  stream        : random draws (oracle.rng.MTStream == the reference's streams,
                  oracle.rng.PhiloxStream == the CUDA kernels' streams)
so that pinning the logic against the reference with MTStream carries over to the
Philox-driven comparison with the GPU.
"""
import numpy as np

from .rng import PhiloxStream, MTStream, u32_to_index, TAG_WALK, TAG_STRUC

PAD = 0  # config.py:8


# ----------------------------------------------------------------------------------------
# graph adapters
class SortedAdj:
    """CSR with sorted neighbour lists, nodes 1..N (the layout the CUDA kernels use)."""

    def __init__(self, n_nodes, edges):
        """edges: iterable of (u, v), 1-indexed, undirected (both directions are inserted)."""
        nb = [set() for _ in range(n_nodes + 1)]
        for u, v in edges:
            nb[u].add(v)
            nb[v].add(u)
        self.n_nodes = n_nodes
        self.adj = [sorted(s) for s in nb]
        self.adjset = nb

    @classmethod
    def from_csr(cls, rowptr, col):
        """rowptr/col 0-indexed CSR (node i <-> id i+1)."""
        obj = cls.__new__(cls)
        n = len(rowptr) - 1
        obj.n_nodes = n
        obj.adj = [[]] + [sorted(int(c) + 1 for c in col[rowptr[i]:rowptr[i + 1]]) for i in range(n)]
        obj.adjset = [set(a) for a in obj.adj]
        return obj

    def all_nodes(self):
        return list(range(1, self.n_nodes + 1))

    def patch(self, nodes):
        """induced-subgraph view on the non-PAD ids of a patch (duplicates allowed)."""
        return _SortedPatch(self, nodes)

    def neighbors(self, n):
        return self.adj[n]

    def has_edge(self, a, b):
        return b in self.adjset[a]

    def degree(self, n):
        # networkx counts a self loop twice in degree(); graphs here have none.
        return len(self.adj[n]) + (1 if n in self.adjset[n] else 0)


class _SortedPatch:
    def __init__(self, g, nodes):
        self.order = sorted(set(int(n) for n in nodes))
        self.pset = set(self.order)
        self.g = g

    def neighbors(self, n):
        return [m for m in self.g.adj[n] if m in self.pset]


class NxAdapter:
    """Wraps a networkx graph and reproduces the *iteration orders* the reference sees
    (node order of G / of G.subgraph(ids), adjacency order).  Used only to pin the walk logic
    against the reference with the reference's own MT streams."""

    def __init__(self, G):
        self.G = G
        self.n_nodes = G.number_of_nodes()

    def all_nodes(self):
        return list(self.G.nodes())

    def patch(self, nodes):
        return _NxPatch(self.G.subgraph(np.asarray(nodes)))       # anchor_patch_samplers.py:138

    def neighbors(self, n):
        return list(self.G.neighbors(n))

    def has_edge(self, a, b):
        return self.G.has_edge(a, b)

    def degree(self, n):
        return self.G.degree(n)


class _NxPatch:
    def __init__(self, view):
        self.view = view
        self.order = list(view.nodes())
        self.pset = set(self.order)

    def neighbors(self, n):
        return list(self.view.neighbors(n))


# ----------------------------------------------------------------------------------------
# stream glue: one "move" needs (maybe) a coin and one uniform choice
def _move(stream, n_tri, n_non, beta):
    """Returns (take_triangular, index) following anchor_patch_samplers.py:96-106."""
    if isinstance(stream, PhiloxStream):
        coin, r = stream.coin_and_choice()          # one block per move, always
        if n_tri == 0:
            return False, u32_to_index(r, n_non)
        if n_non == 0:
            return True, u32_to_index(r, n_tri)
        if coin <= np.float32(beta):
            return True, u32_to_index(r, n_tri)
        return False, u32_to_index(r, n_non)
    # reference order of draws
    if n_tri == 0:
        return False, stream.choice(n_non)
    if n_non == 0:
        return True, stream.choice(n_tri)
    if stream.uniform() <= beta and n_tri != 0:
        return True, stream.choice(n_tri)
    return False, stream.choice(n_non)


def split_neighbors(g, nbrs, prev):
    """anchor_patch_samplers.py:41-47 with is_triangle(:20-24) == has_edge(prev, n):
    n is drawn from N(curr), so n in N(prev) ∩ N(curr)  <=>  n in N(prev)."""
    tri, non = [], []
    for n in nbrs:
        (tri if g.has_edge(prev, n) else non).append(n)
    return tri, non


def triangular_random_walk(g, beta, walk_len, start_nodes, nbr_fn, stream, border):
    """anchor_patch_samplers.py:49-113.

    start_nodes : candidate start nodes (inside: the (sub)graph node list, :70; border: in_border_nodes, :78)
    nbr_fn(n)   : neighbours of n in the view being walked (inside: subgraph neighbours, :72/:35;
                  border: base-graph neighbours filtered to all_valid_nodes, :79/:38)
    """
    if len(start_nodes) == 0:      # reference would raise inside np.random.choice; GPU emits an all-PAD walk
        return []
    pick = stream.choice1 if border else stream.choice
    prev = start_nodes[pick(len(start_nodes))]
    nb = nbr_fn(prev)
    if len(nb) == 0:               # :74/:80/:83-84  PAD sentinel -> length-1 walk
        return [prev]
    curr = nb[pick(len(nb))]
    visited = [prev, curr]
    for _ in range(walk_len - 2):  # :88
        tri, non = split_neighbors(g, nbr_fn(curr), prev)
        if len(tri) + len(non) == 0:
            break                  # :94
        take_tri, idx = _move(stream, len(tri), len(non), beta)
        nxt = tri[idx] if take_tri else non[idx]
        prev, curr = curr, nxt
        visited.append(nxt)
    return visited


def border_nodes(g, patch_nodes):
    """subgraph_utils.py:126-144 without the dense adjacency: nodes of the patch with >=1
    neighbour outside the patch, in patch-node order."""
    pset = set(patch_nodes)
    return [n for n in patch_nodes if any(m not in pset for m in g.neighbors(n))]


def perform_random_walks(g, patches, n_walks, walk_len, beta, inside, stream_factory):
    """anchor_patch_samplers.py:118-158.  patches: (P, Lp) int array with PAD.
    stream_factory(patch_index, walk_index) -> stream.  Returns int64 (P, n_walks, walk_len)."""
    patches = np.asarray(patches)
    out = np.zeros((patches.shape[0], n_walks, walk_len), dtype=np.int64)
    for p, row in enumerate(patches):
        nodes = row[row != PAD]
        if nodes.shape[0] == 0:
            continue                                               # :134-135
        view = g.patch(nodes)                                      # :138 induced subgraph (de-duplicated)
        order, pset = view.order, view.pset
        if inside:
            starts = order
            nbr_fn = view.neighbors
        else:
            starts = border_nodes(g, order)                        # :141
            bset = set(starts)
            # :143 valid = in_border ∪ (V \ patch)  ->  n valid  <=>  n in in_border or n not in patch
            nbr_fn = lambda n, pset=pset, bset=bset: [m for m in g.neighbors(n) if (m in bset) or (m not in pset)]
        for w in range(n_walks):                                   # :149
            walk = triangular_random_walk(g, beta, walk_len, starts, nbr_fn, stream_factory(p, w), border=not inside)
            out[p, w, :len(walk)] = walk
    return out


def sample_structure_anchor_patches(g, n_samples, sample_walk_len, beta, stream_factory):
    """anchor_patch_samplers.py:210-243 (structure_patch_type == 'triangular_random_walk').
    stream_factory(i) -> stream for patch i.  Returns int64 (n_samples, max_len) padded with 0."""
    nodes = g.all_nodes()
    walks = []
    for i in range(n_samples):
        walks.append(triangular_random_walk(g, beta, sample_walk_len, nodes, g.neighbors, stream_factory(i), border=False))
    max_len = max(len(w) for w in walks)
    out = np.zeros((n_samples, max_len), dtype=np.int64)
    for i, w in enumerate(walks):
        out[i, :len(w)] = w
    return out


# Philox stream factories shared with the GPU tests ---------------------------------------
def philox_patch_factory(seed):
    return lambda i: PhiloxStream(seed, i, TAG_STRUC)


def philox_walk_factory(seed, n_walks):
    return lambda p, w: PhiloxStream(seed, p * n_walks + w, TAG_WALK)
