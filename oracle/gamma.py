"""Oracle: degree sequences and DTW similarity (structure channel gamma).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  gamma.py:21-49   get_degree_sequence
  gamma.py:51-52   calc_dist
  gamma.py:54-59   calc_dtw  (fastdtw(x, y, dist=calc_dist) with the package default radius=1)
  SubGNN.py:783-833 compute_structure_patch_similarities (all (cc, patch) pairs, padded CCs -> 0)

fastdtw: third-party ``fastdtw==0.3.4`` (SubGNN.yml:109), source not under /root/reference.
Its published pure-Python algorithm is restated here (coarsen by pairwise means, recurse,
expand the projected path by ``radius``, windowed DP with first-minimum tie-break in the
order (i-1,j), (i,j-1), (i-1,j-1)).  PARITY UNPINNED for this function: the reference holds
no golden vectors for it and the package cannot be installed offline; known-answer vectors in
tests/ are hand-derived.
"""
import numpy as np

PAD = 0


def get_degree_sequence(g, nodes, internal=True):
    """gamma.py:21-49.  ``nodes``: 1-D int array with PAD; duplicates are KEPT (one entry per
    listed node) while the induced subgraph is built on the unique set."""
    nodes = [int(n) for n in np.asarray(nodes).reshape(-1) if int(n) != PAD]   # :27
    pset = set(nodes)                                                          # :29 graph.subgraph(nodes)
    internal_seq = [sum(1 for m in g.neighbors(n) if m in pset) + (1 if g.has_edge(n, n) else 0) for n in nodes]  # :30
    if internal:
        return sorted(internal_seq)                                            # :33-36
    full = [g.degree(n) for n in nodes]                                        # :42-45
    return sorted(f - i for f, i in zip(full, internal_seq))                   # :47-49


def calc_dist(a, b):
    """gamma.py:51-52"""
    return ((max(a, b) + 1) / (min(a, b) + 1)) - 1


# ---------------------------------------------------------------------------------------
# fastdtw 0.3.4 pure-Python algorithm, restated
def _dtw_window(x, y, window, dist):
    """Windowed DP.  window: list of (i, j) 0-based cells in row-major scan order, or None for all.
    D[i,j] = dist + min(D[i-1,j], D[i,j-1], D[i-1,j-1]) with first-minimum tie break in that order."""
    n, m = len(x), len(y)
    if window is None:
        window = [(i, j) for i in range(n) for j in range(m)]
    INF = float('inf')
    D = {(0, 0): (0.0, 0, 0)}
    get = lambda i, j: D.get((i, j), (INF, 0, 0))[0]
    for (i0, j0) in window:
        i, j = i0 + 1, j0 + 1
        dt = dist(x[i - 1], y[j - 1])
        best = (get(i - 1, j) + dt, i - 1, j)
        c = get(i, j - 1) + dt
        if c < best[0]:
            best = (c, i, j - 1)
        c = get(i - 1, j - 1) + dt
        if c < best[0]:
            best = (c, i - 1, j - 1)
        D[(i, j)] = best
    path = []
    i, j = n, m
    while not (i == 0 and j == 0):
        path.append((i - 1, j - 1))
        _, i, j = D[(i, j)]
    path.reverse()
    return D[(n, m)][0], path


def _reduce_by_half(x):
    return [(x[i] + x[i + 1]) / 2 for i in range(0, len(x) - len(x) % 2, 2)]


def _expand_window(path, len_x, len_y, radius):
    cells = set(path)
    for (i, j) in path:
        for a in range(-radius, radius + 1):
            for b in range(-radius, radius + 1):
                cells.add((i + a, j + b))
    fine = set()
    for (i, j) in cells:
        fine.update(((2 * i, 2 * j), (2 * i, 2 * j + 1), (2 * i + 1, 2 * j), (2 * i + 1, 2 * j + 1)))
    window = []
    start_j = 0
    for i in range(len_x):
        new_start = None
        for j in range(start_j, len_y):
            if (i, j) in fine:
                window.append((i, j))
                if new_start is None:
                    new_start = j
            elif new_start is not None:
                break
        start_j = new_start
    return window


def fastdtw(x, y, radius=1, dist=calc_dist):
    x = [float(v) for v in x]
    y = [float(v) for v in y]
    return _fastdtw(x, y, radius, dist)


def _fastdtw(x, y, radius, dist):
    min_time_size = radius + 2
    if len(x) < min_time_size or len(y) < min_time_size:
        return _dtw_window(x, y, None, dist)
    d, path = _fastdtw(_reduce_by_half(x), _reduce_by_half(y), radius, dist)
    window = _expand_window(path, len(x), len(y), radius)
    return _dtw_window(x, y, window, dist)


def dtw_exact(x, y, dist=calc_dist):
    x = [float(v) for v in x]
    y = [float(v) for v in y]
    return _dtw_window(x, y, None, dist)


def calc_dtw(component_degree, patch_degree, mode='fastdtw_r1'):
    """gamma.py:54-59.  Empty sequences (padded CCs / all-PAD patches) short-circuit to 0:
    the reference overwrites those entries at SubGNN.py:831."""
    if len(component_degree) == 0 or len(patch_degree) == 0:
        return 0.0
    if mode == 'fastdtw_r1':
        d, _ = fastdtw(component_degree, patch_degree, radius=1)
    elif mode == 'exact':
        d, _ = dtw_exact(component_degree, patch_degree)
    else:
        raise NotImplementedError(mode)
    return 1.0 / (d + 1.0)


def structure_patch_similarities(g, cc_ids, patches, internal, mode='fastdtw_r1'):
    """SubGNN.py:783-833.  cc_ids (n_sub, C, Lcc), patches (P, Lp) -> float32 (n_sub, C, P)."""
    cc_ids = np.asarray(cc_ids)
    patches = np.asarray(patches)
    n_sub, C, _ = cc_ids.shape
    pseq = [get_degree_sequence(g, p, internal) for p in patches]
    out = np.zeros((n_sub, C, len(patches)), dtype=np.float32)
    for s in range(n_sub):
        for c in range(C):
            if cc_ids[s, c, 0] == PAD:
                continue                                   # :831
            cseq = get_degree_sequence(g, cc_ids[s, c], internal)
            for a, ps in enumerate(pseq):
                out[s, c, a] = np.float32(calc_dtw(cseq, ps, mode))
    return out
