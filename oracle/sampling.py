"""Oracle: k-hop border sets, neighbourhood / position anchor sampling, SP-min similarity.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  subgraph_utils.py:146-176        get_component_border_neighborhood_set (ego_graph_dict is None branch)
  SubGNN.py:673-700                initialize_border_sets
  anchor_patch_samplers.py:163-198 sample_neighborhood_anchor_patch
  anchor_patch_samplers.py:200-208 sample_position_anchor_patches
  anchor_patch_samplers.py:316-328 init_anchors_structure
  SubGNN.py:752-781                compute_shortest_path_similarities
  SubGNN.py:575-607                initialize_cc_ids
"""
import numpy as np
import torch

from .rng import PhiloxStream, u32_to_index, TAG_NEIGH, TAG_POS

PAD = 0


def connected_components(g, nodes):
    """SubGNN.py:590-591 — components of the induced subgraph, as sorted id lists ordered by
    their smallest member (the reference's order is networkx set order; callers that compare
    against the reference canonicalise the same way)."""
    nodes = sorted(set(int(n) for n in nodes))
    pset, seen, comps = set(nodes), set(), []
    for s in nodes:
        if s in seen:
            continue
        comp, stack = [], [s]
        seen.add(s)
        while stack:
            u = stack.pop()
            comp.append(u)
            for v in g.neighbors(u):
                if v in pset and v not in seen:
                    seen.add(v)
                    stack.append(v)
        comps.append(sorted(comp))
    return comps


def initialize_cc_ids(g, subgraphs):
    """SubGNN.py:575-607 -> int64 (n_sub, max_n_cc, max_len_cc), PAD = 0."""
    cc = [connected_components(g, s) for s in subgraphs]
    C = max(len(c) for c in cc)
    L = max(len(x) for c in cc for x in c)
    out = np.zeros((len(subgraphs), C, L), dtype=np.int64)
    for s, comps in enumerate(cc):
        for c, comp in enumerate(comps):
            out[s, c, :len(comp)] = comp
    return out


def khop_border_set(g, component, k):
    """subgraph_utils.py:146-176: union of radius-k ego graphs of the component's nodes minus the component."""
    comp = set(int(n) for n in np.asarray(component).reshape(-1) if int(n) != PAD)
    reach = set()
    for src in comp:
        frontier, seen = {src}, {src}
        for _ in range(k):
            nxt = set()
            for u in frontier:
                for v in g.neighbors(u):
                    if v not in seen:
                        seen.add(v)
                        nxt.add(v)
            frontier = nxt
        reach |= seen
    return reach - comp


def initialize_border_sets(g, cc_ids, k):
    """SubGNN.py:673-700 -> int64 (n_sub, C, max_border); each set stored ascending
    (the reference stores python-set order; consumers only sample from it)."""
    cc_ids = np.asarray(cc_ids)
    n_sub, C, _ = cc_ids.shape
    sets = [[sorted(khop_border_set(g, cc_ids[s, c], k)) for c in range(C)] for s in range(n_sub)]
    Lb = max(max(len(x) for x in row) for row in sets)
    out = np.zeros((n_sub, C, max(Lb, 1)), dtype=np.int64)
    for s in range(n_sub):
        for c in range(C):
            out[s, c, :len(sets[s][c])] = sets[s][c]
    return out


def sample_neighborhood_anchor_patch_torch(rows, n_anchors):
    """anchor_patch_samplers.py:163-198 verbatim semantics on torch's global RNG:
    rand = randn; rand[pad] = 0; argmax.  rows: LongTensor (n_sub, C, L) -> (n_sub, C, n_anchors)."""
    n_sub, C, _ = rows.shape
    flat = rows.reshape(n_sub * C, -1)
    samples = []
    for _ in range(n_anchors):
        rand = torch.randn(flat.shape)
        rand[flat == PAD] = PAD
        samples.append(flat[torch.arange(flat.shape[0]), torch.argmax(rand, dim=1)])
    return torch.stack(samples).transpose(0, 1).reshape(n_sub, C, -1).contiguous()


def neighborhood_law(k_valid, width):
    """The law of one draw of the construction above for a row with k valid entries out of
    ``width`` columns: P(PAD) = 2**-k if the row has >=1 pad column (and k>0),
    1 if k == 0, else 0; the remaining mass is uniform over the k valid entries."""
    if k_valid == 0:
        return 1.0
    return 2.0 ** (-k_valid) if k_valid < width else 0.0


def sample_neighborhood_anchor_patch_philox(rows, n_anchors, seed, layer_tag):
    """Same law, driven by Philox exactly as csrc/sample.cu does: for (row r, anchor a) one block
    from counter (r*n_anchors + a, layer_tag, TAG_NEIGH); word1 -> PAD test on its top k bits
    (all zero <=> prob 2**-k), word0 -> uniform index.  rows: int array (n_sub, C, L)."""
    rows = np.asarray(rows)
    n_sub, C, L = rows.shape
    flat = rows.reshape(n_sub * C, L)
    out = np.zeros((n_sub * C, n_anchors), dtype=np.int64)
    for r in range(flat.shape[0]):
        k = int((flat[r] != PAD).sum())
        if k == 0:
            continue
        for a in range(n_anchors):
            st = PhiloxStream(seed, r * n_anchors + a, TAG_NEIGH)
            st.step = layer_tag
            w = st.draw()
            if k < L and k < 32 and (w[1] >> (32 - k)) == 0:
                continue                                    # PAD with probability 2**-k (k >= 32: never)
            out[r, a] = flat[r][u32_to_index(w[0], k)]      # valid entries are left-packed
    return out.reshape(n_sub, C, n_anchors)


def sample_position_philox(pool_sizes, pools, n_anchors, seed, layer_tag):
    """anchor_patch_samplers.py:200-208 with Philox: anchor a of pool i = pool[i][index(word0, len)]
    from counter (i*n_anchors + a, layer_tag, TAG_POS)."""
    out = np.zeros((len(pools), n_anchors), dtype=np.int64)
    for i, pool in enumerate(pools):
        for a in range(n_anchors):
            st = PhiloxStream(seed, i * n_anchors + a, TAG_POS)
            st.step = layer_tag
            out[i, a] = pool[u32_to_index(st.draw()[0], len(pool))]
    return out


def shortest_path_similarities(hop, cc_ids):
    """SubGNN.py:752-781: sim[s,c,:] = min over the component's rows of the hop table; padded CC -> 0.
    hop: (N, N) numeric (raw hop counts, 0 for self and for unreachable)."""
    cc_ids = np.asarray(cc_ids)
    n_sub, C, _ = cc_ids.shape
    out = np.zeros((n_sub, C, hop.shape[1]), dtype=np.float32)
    for s in range(n_sub):
        for c in range(C):
            comp = cc_ids[s, c][cc_ids[s, c] != PAD]
            if len(comp) > 0:
                out[s, c, :] = np.min(hop[comp - 1, :], axis=0)    # :772
    return out


def all_pairs_hops(g, dtype=np.uint8):
    """prepare_dataset/precompute_graph_metrics.py:20-26 semantics: BFS hop counts, 0 for self and unreachable."""
    N = g.n_nodes
    out = np.zeros((N, N), dtype=dtype)
    for s in range(1, N + 1):
        dist = {s: 0}
        frontier = [s]
        d = 0
        while frontier:
            d += 1
            nxt = []
            for u in frontier:
                for v in g.neighbors(u):
                    if v not in dist:
                        dist[v] = d
                        nxt.append(v)
            frontier = nxt
        for v, dv in dist.items():
            out[s - 1, v - 1] = dv
    return out
