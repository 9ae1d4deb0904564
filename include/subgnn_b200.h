/* libsubgnn_b200 — C ABI of the B200-native SubGNN subgraph message-passing hot path.
 *
 * The reference (mims-harvard/SubGNN) is pure Python: it has no FFI layer.  Its seam for this
 * path is a set of Python module functions; the Python package `subgnn_b200`
 * keeps those names/signatures and calls the entry points below through ctypes.  Every entry
 * point cites the reference function (file:line under /root/reference) whose work it does.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (the library never allocates);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream;
 *   - node ids follow the reference: 1-indexed, 0 = PAD (config.py:8); CSR arrays are 0-indexed
 *     (row i <-> node id i+1) with neighbour lists sorted ascending;
 *   - ids/indices are int32, values fp32 unless stated; the hop table is uint8, row-major;
 *   - return value: 0 on success, negative on error (SUBGNN_ERR_*); subgnn_last_error() returns a
 *     thread-local message;
 *   - stochastic ops take a 64-bit Philox4x32-10 key (`seed`) and derive the counter from the
 *     item index, so results do not depend on launch geometry (oracle/rng.py restates the scheme).
 */
#ifndef SUBGNN_B200_H
#define SUBGNN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SUBGNN_ABI_VERSION 1
#define SUBGNN_DTW_EXACT 0
#define SUBGNN_DTW_FASTDTW_R1 1

const char* subgnn_last_error(void);
int subgnn_abi_version(void);
int subgnn_device_sm_count(void);

/* ---- (1) anchor_patch_samplers.py ------------------------------------------------------------- */

/* anchor_patch_samplers.py:210-243 sample_structure_anchor_patches (structure_patch_type ==
 * 'triangular_random_walk'): n_walks triangular walks (:49-113) of length walk_len over the full graph.
 * out_walks[n_walks][walk_len], 1-indexed ids, 0-padded after an early stop. */
int subgnn_walk_full(const int* rowptr, const int* col, int n_nodes, int n_walks, int walk_len, float beta,
                     unsigned long long seed, int* out_walks, void* stream);

/* anchor_patch_samplers.py:118-158 perform_random_walks: for each patch, n_walks triangular walks of
 * length walk_len inside the induced subgraph (border = 0) or over its border (border = 1: start on a
 * patch node with an edge leaving the patch, move over in-border ∪ outside nodes; replaces
 * subgraph_utils.py:126-144 get_border_nodes).  patches[n_patches][patch_stride]: SORTED UNIQUE 1-indexed
 * ids, patch_len[p] valid entries.  out_walks[n_patches][n_walks][walk_len]. */
int subgnn_walk_patch(const int* rowptr, const int* col, int n_nodes, const int* patches, const int* patch_len,
                      int patch_stride, int n_patches, int n_walks, int walk_len, float beta, int border,
                      unsigned long long seed, int* out_walks, void* stream);

/* anchor_patch_samplers.py:163-198 sample_neighborhood_anchor_patch (pad_rule = 1, position_stream = 0) and
 * :200-208 sample_position_anchor_patches (pad_rule = 0, position_stream = 1): n_anchors draws with
 * replacement from each ragged row items[ptr[r] .. ptr[r+1]).  `width` is the padded width of the
 * reference tensor: with pad_rule a row with k < width valid entries yields PAD with probability 2^-k.
 * `step` selects the layer / call.  out[n_rows][n_anchors]. */
int subgnn_sample_rows(const int* ptr, const int* items, int n_rows, int width, int n_anchors, int pad_rule,
                       unsigned long long seed, unsigned step, int position_stream, int* out, void* stream);

/* subgraph_utils.py:146-176 get_component_border_neighborhood_set (ego_graph_dict is None branch) for every
 * component, as used by SubGNN.py:673-700 initialize_border_sets.  Pass 1 writes one bitmap
 * (ceil(n_nodes/32) words) and the set size per component; pass 2 expands to ascending id lists at
 * out[out_ptr[c] ..) (out_ptr = exclusive prefix sum of counts, computed by the caller). */
int subgnn_border_khop_bitmap(const int* rowptr, const int* col, int n_nodes, const int* cc_ptr, const int* cc_nodes, int n_cc, int k,
                              unsigned* bitmaps, int* counts, void* stream);
int subgnn_border_khop_expand(const unsigned* bitmaps, int n_nodes, int n_cc, const long long* out_ptr, int* out, void* stream);

/* ---- (2) gamma.py / similarity precomputation --------------------------------------------------- */

/* SubGNN.py:752-781 compute_shortest_path_similarities: out[row][n_nodes] = min over the component's rows
 * of the hop table (0 for an empty = padded component).  Streaming form (the reference's dense slab). */
int subgnn_sp_min_dense(const unsigned char* hop, int n_nodes, long long hop_stride, const int* cc_ptr, const int* cc_nodes,
                        int n_rows, float* out, void* stream);

/* Same quantity resolved only at the sampled anchors (what subgraph_mpn.py:92-94 later looks up):
 * out[row][a] = min_{u in cc(row)} hop[u][anchors[anchor_row[row]][a] - 1]; anchor_row may be NULL (identity). */
int subgnn_sp_min_gather(const unsigned char* hop, long long hop_stride, const int* cc_ptr, const int* cc_nodes, int n_rows,
                         const int* anchors, const int* anchor_row, int n_anchors, float* out, void* stream);

/* gamma.py:21-49 get_degree_sequence for a batch of id rows (PAD anywhere, duplicates kept):
 * internal != 0 -> sorted degrees inside the induced subgraph, else sorted (deg_G - internal degree).
 * out_seq[n_rows][stride] ascending, zero padded; out_len[n_rows]. */
int subgnn_degree_seq(const int* rowptr, const int* col, const int* rows, int n_rows, int stride, int internal, int* out_seq,
                      int* out_len, void* stream);

/* gamma.py:54-59 calc_dtw over all pairs (SubGNN.py:811-822): out[a][b] = 1/(1+DTW(seqA[a], seqB[b])) with
 * dist = gamma.py:51 calc_dist; 0 when either sequence is empty (SubGNN.py:831).  mode = SUBGNN_DTW_FASTDTW_R1
 * reproduces fastdtw(radius=1) (reference default), SUBGNN_DTW_EXACT the unconstrained DTW. fp64 inside. */
int subgnn_dtw_batch(const int* seqA, const int* lenA, int nA, int strideA, const int* seqB, const int* lenB, int nB, int strideB,
                     int max_len_a, int max_len_b, int mode, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
