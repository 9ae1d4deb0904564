"""On-disk formats of a SubGNN task directory, read and written with the reference's names and layouts
so that reference datasets / similarity caches load directly and caches written here load in the reference.

    <task>/edge_list.txt                 'u v' per line, 0-indexed                       SubGNN.py:525 (nx.read_edgelist)
    <task>/subgraphs.pth                 'n1-n2-..\\tlabel[-label..]\\tsplit' per line     subgraph_utils.py:24-92
    <task>/{gin,graphsaint_gcn}_embeddings.pth   torch.save(FloatTensor (N, D))          SubGNN.py:562
    <task>/shortest_path_matrix.npy      float64 (N, N), 0 = self / unreachable          precompute_graph_metrics.py:20-70
    <task>/degree_sequence.txt           JSON {0-indexed node: degree}                   precompute_graph_metrics.py:47-59
    <task>/ego_graphs.txt                JSON {0-indexed node: [1-hop nodes]}            precompute_graph_metrics.py:34-45
    <task>/similarities/*.npy            caches named as in SubGNN.py:726-728, 852-854, 893-931

Host-side plumbing only (numpy / json / torch.load); the arithmetic that FILLS these files runs in the CUDA kernels.
"""
import json
import os
from pathlib import Path

import numpy as np
import torch

from . import PAD_VALUE


# ---- readers ---------------------------------------------------------------------------------------------------
def read_edge_list(path):
    """-> (E, 2) int64 array of 0-indexed undirected edges (extra columns ignored, like nx.read_edgelist's data)."""
    rows = []
    with open(path) as f:
        for line in f:
            p = line.split('#', 1)[0].split()
            if len(p) >= 2:
                rows.append((int(p[0]), int(p[1])))
    return np.asarray(rows, dtype=np.int64).reshape(-1, 2)


def read_subgraphs(sub_f):
    """subgraph_utils.py:24-92.  Labels are numbered in order of first appearance; a line with several '-'-joined
    labels makes the task multi-label.  If val has fewer subgraphs than test the two are swapped (:89-90).
    -> ({'train'|'val'|'test': (node lists, label-id lists)}, multilabel, n_labels)"""
    labels, out = {}, {'train': ([], []), 'val': ([], []), 'test': ([], [])}
    multilabel = False
    with open(sub_f) as fin:
        for line in fin:
            parts = line.split('\t')
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
                out[split][1].append([labels[lab] for lab in labs])
    if len(out['val'][0]) < len(out['test'][0]):
        out['val'], out['test'] = out['test'], out['val']
    return out, multilabel, len(labels)


def load_embeddings(path):
    emb = torch.load(path, map_location='cpu')
    if isinstance(emb, torch.nn.Parameter):
        emb = emb.data
    return np.ascontiguousarray(emb.detach().to(torch.float32).numpy())


def load_hop_table(path, n_nodes=None, chunk_rows=2048):
    """shortest_path_matrix.npy (float64 N x N; 26 GB at the EM-USER shape) -> uint8 (N, N), converted in row chunks
    through a memory map so that the fp64 matrix is never resident in host memory."""
    m = np.load(path, mmap_mode='r', allow_pickle=False)
    n = m.shape[0]
    assert m.ndim == 2 and m.shape[1] == n, 'shortest_path_matrix.npy must be square'
    if n_nodes is not None and n != n_nodes:
        raise ValueError('shortest_path_matrix.npy is %d x %d but the graph has %d nodes' % (n, n, n_nodes))
    out = np.empty((n, n), dtype=np.uint8)
    for r in range(0, n, chunk_rows):
        blk = np.asarray(m[r:r + chunk_rows])
        if blk.max(initial=0) > 255:
            raise ValueError('hop counts above 255 do not fit the uint8 hop table')
        out[r:r + chunk_rows] = blk.astype(np.uint8)
    return out


def load_degree_dict(path):
    """degree_sequence.txt -> int64 (N,) degrees indexed by 0-indexed node id, or None if the file is absent."""
    if not Path(path).exists():
        return None
    with open(path) as f:
        d = json.load(f)
    n = max(int(k) for k in d) + 1
    out = np.zeros(n, dtype=np.int64)
    for k, v in d.items():
        out[int(k)] = int(v)
    return out


# ---- writers (the task-directory producers: prepare_dataset / precompute_graph_metrics outputs) -------------------
def write_edge_list(path, edges):
    with open(path, 'w') as f:
        for u, v in np.asarray(edges).reshape(-1, 2):
            f.write('%d %d\n' % (u, v))


def write_subgraphs(path, subgraphs, labels, splits):
    """subgraphs: 0-indexed node lists; labels: a label or list of labels per subgraph; splits: 'train'|'val'|'test'."""
    with open(path, 'w') as f:
        for nodes, lab, sp in zip(subgraphs, labels, splits):
            labs = lab if isinstance(lab, (list, tuple)) else [lab]
            f.write('%s\t%s\t%s\n' % ('-'.join(str(int(n)) for n in nodes), '-'.join(str(x) for x in labs), sp))


def write_embeddings(path, emb):
    torch.save(torch.as_tensor(np.asarray(emb), dtype=torch.float32), path)


def write_graph_metrics(task_dir, g, hop=True, degrees=True, ego=True):
    """precompute_graph_metrics.py:29-70 from the device graph: hop table -> shortest_path_matrix.npy (float64, the
    reference's dtype), CSR degrees -> degree_sequence.txt, 1-hop neighbour lists -> ego_graphs.txt."""
    task_dir = Path(task_dir)
    (task_dir / 'similarities').mkdir(parents=True, exist_ok=True)
    if hop:
        from . import ops
        h = g.hop if g.hop is not None else ops.hop_table(g)
        np.save(task_dir / 'shortest_path_matrix.npy', h.cpu().numpy().astype(np.float64))
    rp, col = g.rowptr_host, g.col_host
    if degrees:
        with open(task_dir / 'degree_sequence.txt', 'w') as f:
            json.dump({str(i): int(rp[i + 1] - rp[i]) for i in range(g.n_nodes)}, f)
    if ego:
        with open(task_dir / 'ego_graphs.txt', 'w') as f:
            json.dump({str(i): [int(x) for x in col[rp[i]:rp[i + 1]]] for i in range(g.n_nodes)}, f)


# ---- similarity caches --------------------------------------------------------------------------------------------
class SimilarityCache:
    """File names of <task>/similarities/ exactly as the reference builds them; load() honours
    ``compute_similarities`` (SubGNN.py:731 ff.: an existing file is ignored when it is set)."""

    def __init__(self, sim_path, hp, write=True, dense_limit_bytes=1 << 30):
        self.dir = Path(sim_path)
        self.hp = hp
        self.write = write
        self.dense_limit_bytes = dense_limit_bytes
        self.recompute = bool(hp.get('compute_similarities', False))
        self.loaded, self.saved = [], []

    # names ---------------------------------------------------------------------------------------------------
    def _struc_tag(self):
        hp = self.hp
        return '%d_%s_%d' % (hp['sample_walk_len'], hp['structure_patch_type'], hp['max_sim_epochs'])

    def border_set(self, split):                                                     # SubGNN.py:726-728
        return '%d_%d_%s_border_set.npy' % (self.hp['neigh_sample_border_size'], PAD_VALUE, split)

    def np_sim(self, split):                                                         # SubGNN.py:852-854
        return '%d_%s_similarities.npy' % (PAD_VALUE, split)

    def struc_patches(self):                                                         # SubGNN.py:893
        return 'struc_patches_%s.npy' % self._struc_tag()

    def walks(self, inside):                                                         # SubGNN.py:904, 913
        hp = self.hp
        return '%s_struc_patch_random_walks_%d_%d_%s.npy' % ('int' if inside else 'bor', hp['n_triangular_walks'], hp['random_walk_len'],
                                                             self._struc_tag())

    def struc_sim(self, inside, split):                                              # SubGNN.py:924-931
        fn = self.hp.get('structure_similarity_fn', 'dtw')
        return '%s_struc_%s_%d%s_%s_similarities.npy' % ('int' if inside else 'bor', self._struc_tag(), PAD_VALUE,
                                                         '' if fn == 'dtw' else '_' + fn, split)

    # io ------------------------------------------------------------------------------------------------------
    def load(self, name):
        p = self.dir / name
        if self.recompute or not p.exists():
            return None
        self.loaded.append(name)
        return np.load(p, allow_pickle=True)

    def save(self, name, arr, dense=False):
        if not self.write:
            return False
        arr = np.asarray(arr)
        if dense and arr.nbytes > self.dense_limit_bytes:
            return False
        os.makedirs(self.dir, exist_ok=True)
        np.save(self.dir / name, arr)
        self.saved.append(name)
        return True


class MemoryCache(SimilarityCache):
    """The same lookups served from an already prepared dict (anchor resampling, SubGNN.py:449-457: the patches, walks,
    border sets and similarities are kept, only the anchor draws change)."""

    def __init__(self, hp, prepared):
        super().__init__('.', hp, write=False)
        self.recompute = False
        self.p = prepared

    def load(self, name):
        p = self.p
        if p.get('structure_anchors') is not None:
            for n, key in ((self.struc_patches(), 'structure_anchors'), (self.walks(True), 'int_rw_all'), (self.walks(False), 'bor_rw_all')):
                if name == n:
                    return np.asarray(p[key])
        for s in p['cc_ids']:
            if p.get('I_S_sim') and name == self.struc_sim(True, s):
                return np.asarray(p['I_S_sim'][s])
            if p.get('B_S_sim') and name == self.struc_sim(False, s):
                return np.asarray(p['B_S_sim'][s])
            if name == self.border_set(s) and s in p.get('N_border', {}):
                ptr, items = p['N_border'][s]
                n_sub, C = np.asarray(p['cc_ids'][s]).shape[:2]
                return pad_ragged(np.asarray(ptr), np.asarray(items), (n_sub, C))
            if name == self.np_sim(s) and p.get('NP_sim') and s in p['NP_sim']:
                return np.asarray(p['NP_sim'][s])
        return None


def pad_ragged(ptr, items, n_rows_shape, dtype=np.int64):
    """ragged rows (ptr [n+1], items) -> padded (.., L) array with PAD right-fill; n_rows_shape = leading dims."""
    ptr = np.asarray(ptr, dtype=np.int64)
    items = np.asarray(items)
    lens = ptr[1:] - ptr[:-1]
    n = lens.shape[0]
    L = int(lens.max()) if n else 0
    out = np.full((n, max(L, 1)), PAD_VALUE, dtype=dtype)
    if items.size:
        rows = np.repeat(np.arange(n), lens)
        cols = np.arange(items.shape[0]) - np.repeat(ptr[:-1], lens)
        out[rows, cols] = items
    return out.reshape(tuple(n_rows_shape) + (out.shape[1],))
