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
#define SUBGNN_DTW_EXACT_THREAD 2   /* exact DTW on the thread-per-pair mapping (same values as SUBGNN_DTW_EXACT) */

const char* subgnn_last_error(void);
int subgnn_abi_version(void);
int subgnn_device_sm_count(void);
/* number of kernels this library has launched (or recorded into a CUDA graph under capture) so far */
unsigned long long subgnn_launch_count(void);
/* test aid: the distinct kernel template instantiations launched so far, ';'-separated ("row_fwd<2>;lstm_fwd_tile<1,4,64>;..."),
 * written into buf (cap bytes, NUL terminated); returns their number.  _reset clears the log. */
int subgnn_variant_log(char* buf, int cap);
void subgnn_variant_log_reset(void);

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
/* the same over a subset of the A rows: pairs (rowsA[i], b) for i < n_rows (rowsA == NULL: rows 0 .. n_rows-1); out keeps its
 * [nA][nB] layout and is written at the ORIGINAL row.  max_len_a bounds the lengths of the listed rows only: the host buckets
 * the components by length so that the wavefront of SUBGNN_DTW_EXACT runs with as many lanes per pair as the rows it holds. */
int subgnn_dtw_batch_rows(const int* seqA, const int* lenA, const int* rowsA, int n_rows, int strideA, const int* seqB, const int* lenB,
                          int nB, int strideB, int max_len_a, int max_len_b, int mode, float* out, void* stream);


/* prepare_dataset/precompute_graph_metrics.py:20-26 get_shortest_path for sources [src_begin, src_end): BFS hop counts
 * into the uint8 table (0 = self / unreachable), i.e. the content of shortest_path_matrix.npy (SubGNN.py:848). */
int subgnn_hop_table(const int* rowptr, const int* col, int n_nodes, int src_begin, int src_end, unsigned char* hop,
                     long long hop_stride, void* stream);

/* ---- dense building blocks (gemm.cu) ------------------------------------------------------------- */

/* y[m][n] = act(sum_k X[row(m)][k] W[n][k] + bias[n]); row(m) = gather_ids ? gather_ids[m] : m.
 * Replaces nn.Linear forward at subgraph_mpn.py:239, SubGNN.py:74,306-310 and the LSTM input projections
 * (SubGNN.py:73); with gather_ids it also replaces the embedding lookup anchor_patch_samplers.py:409. */
int subgnn_linear_fwd(const float* x, int ldx, const int* gather_ids, const float* w, int ldw, const float* bias, float* y, int ldy,
                      int M, int N, int K, int relu, void* stream);
/* dx[row(m)][k] (+)= sum_n dy[m][n] W[n][k]; with scatter_ids the rows are atomically added into a table
 * gradient (embedding_dense_backward), id 0 (PAD) skipped. */
int subgnn_linear_bwd_input(const float* dy, int ldy, const float* w, int ldw, float* dx, int lddx, const int* scatter_ids, int M, int N,
                            int K, int accumulate, void* stream);
/* dW[n][k] += sum_m dy[m][n] X[row(m)][k];  db[n] += sum_m dy[m][n] (db may be NULL).  m_dev (may be NULL): device
 * int holding the number of valid rows (<= M) when it is only known on the device. */
int subgnn_linear_bwd_weight(const float* dy, int ldy, const float* x, int ldx, const int* gather_ids, float* dw, int lddw, float* db,
                             int M, int N, int K, const int* m_dev, void* stream);

/* tcgen05 (kind::tf32, 3xTF32 error-compensated => fp32-accurate) versions of the three dense entry points above, same contracts
 * (tcgemm.cu).  Rows must be 16-byte aligned (ld % 4 == 0). */
int subgnn_tc_linear_fwd(const float* x, int ldx, const int* gather_ids, const float* w, int ldw, const float* bias, float* y, int ldy,
                         int M, int N, int K, int relu, void* stream);
int subgnn_tc_linear_bwd_input(const float* dy, int ldy, const float* w, int ldw, float* dx, int lddx, const int* scatter_ids, int M, int N,
                               int K, int accumulate, void* stream);
int subgnn_tc_linear_bwd_weight(const float* dy, int ldy, const float* x, int ldx, const int* gather_ids, float* dw, int lddw, float* db,
                                int M, int N, int K, void* stream);

/* Grouped form (tcgemm_ws.cu): up to 6 dense products in ONE persistent, warp-specialised launch (TMA-fed 3-stage operand ring,
 * converter warps for the 3xTF32 low parts, double-buffered TMEM accumulators).  The three entry points above route their dense
 * (gather_ids == NULL) calls here; the step engine groups the gate-gradient consumers of an LSTM layer (input gradient + three
 * weight gradients) into one launch.  op selects the product:
 *   SUBGNN_GEMM_FWD               out[M][N]      = a[M][K] . b[N][K]^T + bias [, relu]             (a = x, b = W)
 *   SUBGNN_GEMM_BWD_INPUT         out[row(m)][K] (+)= a[M][N] . b[N][K]     row(m) = scatter_ids ? scatter_ids[m] : m  (a = dy, b = W)
 *   SUBGNN_GEMM_BWD_WEIGHT        out[N][K]     += a[M][N]^T . b[M][K]                             (a = dy, b = x)
 *   SUBGNN_GEMM_BWD_WEIGHT_SHIFT  out[N][K]     += sum_m a[m][N]^T . b[m + shift][K] over the rows with (m % period) != (shift < 0 ? 0 :
 *                                 period - 1): the recurrent-weight gradient dW_hh = sum_t dG_t^T h_(t-1) of nn.LSTM (SubGNN.py:73)
 *                                 for sequences of `period` steps stored row after row (shift = -1 forward, +1 reverse direction).
 * accumulate (BWD_INPUT without scatter_ids): 0 = store, 1 = add onto the destination (read-add-store; one writer per element),
 * 2 = add with atomics (another problem of the same group adds into the same rows; the destination must hold the initial value).
 * max_ctas > 0 caps the grid (a companion launch leaves the other SMs to the kernel on the critical chain). */
#define SUBGNN_GEMM_FWD 0
#define SUBGNN_GEMM_BWD_INPUT 1
#define SUBGNN_GEMM_BWD_WEIGHT 2
#define SUBGNN_GEMM_BWD_WEIGHT_SHIFT 3
typedef struct subgnn_gemm_desc {
  const float* a;
  const float* b;
  float* out;
  const float* bias;
  const int* scatter_ids;
  int op, lda, ldb, ldo, M, N, K, relu, accumulate, shift, period;
} subgnn_gemm_desc;
int subgnn_gemm_desc_size(void);
int subgnn_tc_ws_available(void);   /* 1 when the driver exports cuTensorMapEncodeTiled (TMA descriptors) */
int subgnn_tc_gemm_group(const subgnn_gemm_desc* problems, int n_problems, int max_ctas, void* stream);
/* out[m][:] = table[ids[m]][:] (anchor_patch_samplers.py:409: embedding lookup of the walk nodes, once per step) */
int subgnn_gather_rows(const float* table, const int* ids, float* out, int M, int D, void* stream);

/* ---- walk-encoder LSTM (lstm.cu): SubGNN.py:60-88, anchor_patch_samplers.py:413-433 ------------------ */
int subgnn_lstm_prep(const float* whh, const float* b_ih, const float* b_hh, float* whh_t, float* bsum, int H, void* stream);
/* the same for every layer of a stack in ONE launch: HOST arrays of n_layers device pointers each */
int subgnn_lstm_prep_layers(const float* const* whh, const float* const* b_ih, const float* const* b_hh, float* const* whh_t,
                            float* const* bsum, int n_layers, int H, void* stream);
int subgnn_lstm_recur_fwd(float* G, const float* whh_t, float* OUT, float* CS, int n_seq, int T, int H, int steps_fwd, int steps_rev,
                          void* stream);
/* db_ih / db_hh (optional, [2][4H]): += column sums of d(pre-activation), the gradient of both bias vectors.
   zero_untaken is a flag word: SUBGNN_LSTM_ZERO_UNTAKEN = zero the gate slots of time steps that were never taken;
   SUBGNN_LSTM_DOUT_LAST_ONLY = dOUT carries gradient (and has been written) only in its rows t = T-1 — the 'last' aggregator's
   top layer, SubGNN.py:83 — every other row is taken as zero without being read (register-tiled recurrence only:
   subgnn_lstm_fused_dropout_supported(H)). */
#define SUBGNN_LSTM_ZERO_UNTAKEN 1
#define SUBGNN_LSTM_DOUT_LAST_ONLY 2
int subgnn_lstm_recur_bwd(float* G, const float* whh, const float* OUT, const float* CS, const float* dOUT, int n_seq, int T, int H,
                          int steps_fwd, int steps_rev, int zero_untaken, float* db_ih, float* db_hh, void* stream);
/* the same recurrences with nn.LSTM's inter-layer dropout fused in (training, p > 0): forward also writes xdrop = dropout_p(OUT), the
   next layer's input; backward applies the same mask to dOUT (then the gradient w.r.t. xdrop) on load.  Mask stream = the one of
   subgnn_dropout(seed, salt, step_dev).  Available when subgnn_lstm_fused_dropout_supported(H). */
int subgnn_lstm_fused_dropout_supported(int H);
/* the same forward recurrence on the tensor cores (lstm_tc.cu: tcgen05 3xTF32, 4-CTA clusters split by hidden unit, h exchanged through
 * distributed shared memory); H = 64 only; whh is the NATIVE nn.LSTM weight_hh [2][4H][H] (not the packed copy); xdrop may be NULL */
int subgnn_lstm_recur_fwd_tc_supported(int H);
int subgnn_lstm_recur_fwd_tc(float* G, const float* whh, float* OUT, float* CS, int n_seq, int T, int H, int steps_fwd, int steps_rev,
                             float* xdrop, float p, unsigned long long seed, unsigned salt, const int* step_dev, void* stream);
int subgnn_lstm_recur_fwd_drop(float* G, const float* whh_t, float* OUT, float* CS, int n_seq, int T, int H, int steps_fwd, int steps_rev,
                               float* xdrop, float p, unsigned long long seed, unsigned salt, const int* step_dev, void* stream);
int subgnn_lstm_recur_bwd_drop(float* G, const float* whh, const float* OUT, const float* CS, const float* dOUT, int n_seq, int T, int H,
                               int steps_fwd, int steps_rev, int zero_untaken, float* db_ih, float* db_hh, float p, unsigned long long seed,
                               unsigned salt, const int* step_dev, void* stream);
/* BPTT with a second gradient source: dOUT_add has dOUT's layout, only its rows t = T-1 are written (and read) — the input gradient of
   the reverse direction of the layer ABOVE, which consumed this layer's output at the last step only ('last' aggregator, SubGNN.py:83).
   Lets that product run in the same grouped launch as the all-rows input gradient instead of behind it.  p = 0: no dropout mask. */
int subgnn_lstm_recur_bwd_add(float* G, const float* whh, const float* OUT, const float* CS, const float* dOUT, const float* dOUT_add, int n_seq,
                              int T, int H, int steps_fwd, int steps_rev, int zero_untaken, float* db_ih, float* db_hh, float p,
                              unsigned long long seed, unsigned salt, const int* step_dev, void* stream);
/* walk-group head (anchor_patch_samplers.py:429-433: patch embedding = sum over its walks of Linear(agg(lstm_out)), SubGNN.py:83-88).
   The head is linear, so the walks are summed first: fwd  AGG[g] = sum_w agg(OUT[g*group+w]),  EMB[g] = W AGG[g] + group * bias;
   bwd  dOUT rows <- W^T dEMB[g] (t = T-1 only for 'last'),  db += group * sum_g dEMB[g]   (dW = dEMB^T AGG: subgnn_linear_bwd_weight).
   sum_mode: 0 = 'last', 1 = 'sum', SUBGNN_HEAD_LAST_NO_FILL (bwd only) = 'last' writing the rows t = T-1 alone: the other rows of
   dOUT are left untouched, for a BPTT launched with SUBGNN_LSTM_DOUT_LAST_ONLY */
#define SUBGNN_HEAD_LAST_NO_FILL 2
int subgnn_lstm_head_fwd(const float* OUT, float* AGG, float* EMB, const float* W, const float* bias, int n_groups, int group, int T, int H2,
                         int D, int sum_mode, void* stream);
int subgnn_lstm_head_bwd(const float* dEMB, const float* W, float* dOUT, float* db, int n_groups, int group, int T, int H2, int D,
                         int sum_mode, void* stream);
int subgnn_dropout(const float* x, float* y, long long n, float p, unsigned long long seed, unsigned salt, const int* step_dev,
                   void* stream);
int subgnn_colsum(const float* dy, int ldy, float* db, int M, int N, const int* m_dev, void* stream);
int subgnn_model_desc_size(void);

/* ---- fused per-step model path (model.cu) ------------------------------------------------------------
 * One descriptor names every device buffer of a training / inference step.  Rows = the valid connected
 * components of the batch's subgraphs (ragged; padded components never influence logits or gradients
 * because SubGNN.py:303 masks them in the readout).  Z column layout == the reference's concat order
 * (SubGNN.py:271-295): [init D | per layer: N_in D, N_out D | P_in A_pi, P_out A_pb | S_in A_s, S_out A_s].
 * MPN parameter block of (channel, layer l, side s) at mpn_params[ch] + (2 l + s) * (2 D D + 2 D + 1) floats:
 * linear.weight (D x 2D) | linear.bias (D) | linear_position.weight (D) | linear_position.bias (1). */
typedef struct subgnn_model_desc {
  /* split tables (built once) */
  const int* sub_ccptr;      /* [n_sub+1] first component row of each subgraph */
  const int* sub_maxlen;     /* [n_sub]   largest component of each subgraph */
  const int* cc_nodeptr;     /* [n_cc+1] */
  const int* cc_nodes;       /* ids of all components, concatenated */
  const int* n_ids[2];       /* [L][n_cc][A_n*]  neighbourhood anchors (internal, border), 0 = PAD */
  const float* n_sim[2];     /* [L][n_cc][A_n*]  resolved similarity of (component, anchor) */
  const int* p_int_ids;      /* [L][n_sub][A_pi] */
  const int* p_bor_ids;      /* [L][A_pb] */
  const float* p_sim[2];     /* [L][n_cc][A_pi | A_pb] */
  const float* s_sim[2];     /* [L][n_cc][A_s] */
  const int* labels;         /* [n_sub] class id (multiclass) */
  const float* labels_multi; /* [n_sub][n_classes] 0/1 (multilabel) */
  /* parameters / gradients */
  const float* E;            /* [(N+1)][D] node embeddings, row 0 = 0 */
  float* dE;
  const float* mpn_params[3];   /* N, P, S */
  float* mpn_grads[3];
  const float* cc_tab[2];    /* trainable_cc: N_I, N_B tables [n_sub][C_pad][D]; else NULL */
  float* cc_tab_grad[2];
  const float* lin_w[3];     /* lin (h1 x hid), lin2 (h2 x h1), lin3 (K x h2) */
  const float* lin_b[3];
  float* lin_gw[3];          /* gradients of lin / lin2 / lin3 (may be NULL for inference) */
  float* lin_gb[3];
  const float* emb_s;        /* [2][L][A_s][D] structure anchor embeddings (LSTM output) */
  float* d_emb_s;
  /* per-step scratch */
  const int* batch_idx;      /* [B] subgraph indices of this step */
  const int* step_dev;       /* [1] optimizer step counter on the device (dropout salt; graph-replay safe) */
  int* b_rowptr;             /* [B+1] */
  int* meta;                 /* [0] = R (rows in the batch), [1] = max component length in the batch */
  int* row_b; int* row_g;    /* [R_cap] batch slot and global component id of every row */
  float* n_wt;               /* [L][2][2D][D] transposed MPN weights (N channel) */
  float* lin_wt[3];          /* transposed MLP weights */
  float* q_pi; float* q_pb; float* q_s;          /* [L][B][A_pi], [L][A_pb], [L][2][A_s] */
  float* dq_pi; float* dq_pb; float* dq_s;
  float* X0;                 /* [R_cap][D] pooled component embeddings */
  float* Nh;                 /* [L+1][2][R_cap][D] N-channel component embeddings per layer */
  float* Nagg;               /* [L][2][R_cap][D] aggregated messages */
  float* Ndpre;              /* [L][2][R_cap][D] d(pre-activation) */
  float* Z;                  /* [B][hid] subgraph embeddings */
  float* H1; float* H2;      /* [B][h1], [B][h2] post-relu post-dropout */
  float* logits;             /* [B][K] */
  float* loss_b;             /* [B] per-sample loss terms (already divided by B [and K]) */
  float* loss_sum;           /* [1] zeroed by the caller; subgnn_model_readout adds the per-sample terms (the step's loss) */
  float* dlogits; float* dH2; float* dH1; float* dZ;
  unsigned long long seed;   /* dropout key */
  int n_sub, n_cc, n_nodes;
  int D, L, hid, h1, h2, n_classes;
  int use_n, use_p, use_s;
  int A_ni, A_nb, A_pi, A_pb, A_s;
  int pool_max, trainable_cc, C_pad, multilabel, training, use_proj;
  int B, R_cap;
  unsigned step;             /* dropout salt (optimizer step counter) */
  float lin_dropout;
  int mlp_fused;             /* subgnn_model_readout also accumulates the MLP weight / bias gradients (from its own d logits): the fused
                                training step; subgnn_model_wgrad then skips them */
  int wgrad_rows;            /* rows of the batch per CTA of the N-channel weight-gradient launch (0: one chunk of up to 512 rows: few CTAs,
                                out of the way of a long BPTT chain; 64: many CTAs, when that launch itself ends the backward pass) */
} subgnn_model_desc;

/* batch bookkeeping + per-step weight transposes + q = w_p . x_anchor for every shared anchor list */
int subgnn_model_prep_batch(const subgnn_model_desc* d, void* stream);
int subgnn_model_prep_weights(const subgnn_model_desc* d, void* stream);
#define SUBGNN_Q_POS 1     /* position-channel anchors (node embeddings) */
#define SUBGNN_Q_STRUC 2   /* structure-channel anchors (LSTM output emb_s) */
int subgnn_model_q_fwd_part(const subgnn_model_desc* d, int which, void* stream);
/* The row pass in phases, so that the neighbourhood chains (independent of the walk-encoder LSTM) can run concurrently
 * with it on another stream: SUBGNN_PHASE_N = pooling + N channel, SUBGNN_PHASE_PS = P / S property-aware outputs. */
#define SUBGNN_PHASE_N 1
#define SUBGNN_PHASE_P 2    /* position-channel property-aware outputs (depend on the batch only) */
#define SUBGNN_PHASE_S 4    /* structure-channel property-aware outputs (depend on the walk-encoder LSTM) */
#define SUBGNN_PHASE_PS 6
/* rows_fwd + mlp_fwd together are SubGNN.py:225-312 forward for the batch (cc pooling :609-622, all SG_MPN layers
 * subgraph_mpn.py:133-241, masked_sum readout subgraph_utils.py:213-237, MLP :306-310) + loss :338-342; when d->training
 * mlp_fwd also runs the MLP backward down to dZ. */
int subgnn_model_rows_fwd(const subgnn_model_desc* d, int phases, void* stream);
int subgnn_model_mlp_fwd(const subgnn_model_desc* d, void* stream);       /* readout MLP + loss (+ MLP backward when training); H1 must be zero on entry (split-K target) */
/* The same in stages, restricted to column sets of Z: only the structure-channel columns of Z depend on the LSTM, so the first-layer
 * product over every other column (SUBGNN_COLS_NOT_S) runs beside the LSTM chain and only the SUBGNN_COLS_S slices, the rest of the
 * MLP and the S columns of dZ stay on the step's critical chain.  stages: bit mask, launched in this order. */
#define SUBGNN_MLP_LIN1 1   /* H1pre += Z[:, cols] W1[:, cols]^T */
#define SUBGNN_MLP_REST 2   /* bias / relu / dropout, lin2, lin3, loss, d logits, dH2, dH1 */
#define SUBGNN_MLP_DZ 4     /* dZ[:, cols] = dH1 W1[:, cols] (training only) */
#define SUBGNN_COLS_ALL 0
#define SUBGNN_COLS_NOT_S 1
#define SUBGNN_COLS_S 2
int subgnn_model_mlp_stage(const subgnn_model_desc* d, int stages, int cols, void* stream);
/* backward of all rows (autograd of the above): N-channel chains, property-aware outputs, pooling; scatters into dE / dq / Ndpre */
int subgnn_model_rows_bwd(const subgnn_model_desc* d, int phases, void* stream);
/* The readout section as ONE launch of 8-CTA clusters (SubGNN.py:303-310 masked-sum output -> lin -> lin2 -> lin3, loss :338-342
 * and, when training, the backward down to dZ and — with d->mlp_fused — the MLP weight / bias gradients): every CTA of a cluster
 * owns a slice of the hid columns of Z / W1 (kept in shared memory for forward, dZ and dW1), partial first-layer sums are
 * exchanged through distributed shared memory; a cluster serves 8 samples (one per CTA behind the first layer).  Same outputs and dropout masks as
 * subgnn_model_mlp_fwd (+ the MLP part of subgnn_model_wgrad); needs d->loss_sum zeroed.  _supported: shapes the kernel takes. */
int subgnn_model_readout_supported(const subgnn_model_desc* d);
int subgnn_model_readout(const subgnn_model_desc* d, void* stream);
/* per-sample MLP backward from externally supplied dlogits (autograd entry) */
int subgnn_model_mlp_bwd(const subgnn_model_desc* d, void* stream);
int subgnn_model_q_bwd(const subgnn_model_desc* d, void* stream);          /* == q_bwd_part(POS | STRUC) */
/* backward of q = w_p . x per anchor list: SUBGNN_Q_STRUC -> d emb_s (feeds the LSTM head gradient), SUBGNN_Q_POS -> rows of dE */
int subgnn_model_q_bwd_part(const subgnn_model_desc* d, int which, void* stream);
/* weight gradients of the N-channel MPN projections and (unless d->mlp_fused says they are produced elsewhere) the MLP */
int subgnn_model_wgrad(const subgnn_model_desc* d, void* stream);
/* the MLP's weight / bias gradients alone (dW1 = dH1^T Z, dW2 = dH2^T H1, dW3 = dlogits^T H2: SubGNN.py:303-310 backward): ready as soon
 * as the readout section has run, so a caller can launch them beside the backward chain and set mlp_fused for subgnn_model_wgrad */
int subgnn_model_mlp_wgrad(const subgnn_model_desc* d, void* stream);

/* ---- stand-alone SG_MPN.forward (mpn.cu): subgraph_mpn.py:36-131,176-231 on the reference's materialised inputs ----
 * cc (R x D), anchor_embeds (R x A x D), sims (R x n_opt), mask (R x A, uint8).  Exactly one of anchor_ids (R x A: first node id
 * of every anchor patch; similarity column = id - 1, N/P channels :92-94) and sim_index (A: similarity column per anchor, S channel
 * :96-99) is non-NULL.  Outputs: cat (R x 2D) = [cc | agg], pos_lin (R x A) = w_p . msg + b_p, s_eff (R x A) = mask * sim. */
int subgnn_mpn_fwd(const float* cc, const float* anchor_embeds, const float* sims, int n_opt, const int* anchor_ids, const int* sim_index,
                   const unsigned char* mask, const float* wp, const float* bp, float* cat, float* pos_lin, float* s_eff, int R, int A,
                   int D, void* stream);
int subgnn_mpn_bwd(const float* anchor_embeds, const float* s_eff, const float* dcat, const float* dpos, const float* wp, float* dx, float* dwp,
                   float* dbp, int R, int A, int D, void* stream);

/* ---- optimizer (optim.cu): SubGNN.py:1156-1164 Adam + Lightning's clip_grad_norm_ ---------------------- */
int subgnn_fill_zero(float* p, long long n, void* stream);
int subgnn_grad_sumsq(const float* g, long long n, float* out_sumsq /* [1], accumulated */, void* stream);
int subgnn_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                     const int* step_dev /* [1] step count t >= 1 */, const float* sumsq_dev, float clip_norm, float grad_scale,
                     void* stream);
/* both of the above in one launch (one CTA per SM, grid barrier between the norm and the update): sumsq_dev (may be NULL) is
 * ACCUMULATED like subgnn_grad_sumsq's output; clip_norm <= 0 disables clipping */
int subgnn_clip_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                          const int* step_dev, float* sumsq_dev, float clip_norm, float grad_scale, void* stream);
int subgnn_sum_to_scalar(const float* x, int n, float* out, void* stream);

/* ---- data-parallel exchange fused with the optimizer (dp.cu) ------------------------------------------------------------------
 * The gradient / parameter arenas of all ranks are symmetric allocations mapped into every process (NVLink peer memory); the
 * pointer tables are HOST arrays of device addresses, one per rank.  Rank r owns shard r = elements [r * shard, (r + 1) * shard):
 *   reduce_scatter: gsum[i] = sum over ranks of grads_q[r * shard + i] (fixed order); slot r of EVERY rank's `slots` array receives
 *                   the shard's sum of squares.  Waits (inside the kernel) until every rank has signalled that its gradients are complete.
 *   adam_allgather: waits for every rank's reduce_scatter; global norm from the local slots (fixed order) -> clip_grad_norm_
 *                   coefficient; Adam on the shard with the rank's own m / v; the new parameters are stored into shard r of every
 *                   rank's parameter arena; a closing 1-thread kernel waits until every shard of the own arena has been written and
 *                   advances *epoch_dev.
 * Cross-GPU synchronisation: peer_flags[q] -> rank q's symmetric flag array of subgnn_dp_flag_words() 32-bit words (zero-initialised),
 * *epoch_dev (device counter, starts at 1, advanced once per exchange identically on every rank) is the value signalled.
 * epoch_dev == NULL: no synchronisation inside the kernels — the caller brackets them with its own cross-GPU barriers (before
 * reduce_scatter, between the two, after adam_allgather).
 * Together: the averaged-gradient Adam step of SubGNN.py:1156-1164 under data parallelism, with sharded optimizer state.
 * mc_grads / mc_params (may be NULL): NVLS multicast mappings of the two arenas — the shard sum becomes one multimem.ld_reduce per
 * 16 bytes (reduced inside the NVSwitch), the parameter broadcast one multimem.st. */
int subgnn_dp_flag_words(void);
int subgnn_dp_reduce_scatter(const unsigned long long* peer_grads, const unsigned long long* peer_slots, const unsigned long long* peer_flags,
                             const float* mc_grads, const unsigned* epoch_dev, int world, int rank, long long n, long long shard, float* gsum,
                             void* stream);
int subgnn_dp_adam_allgather(const unsigned long long* peer_params, const unsigned long long* peer_flags, float* mc_params, unsigned* epoch_dev, int world,
                             int rank, long long n, long long shard, const float* gsum, float* m, float* v, float lr, float beta1, float beta2,
                             float eps, const int* step_dev, const float* slots_local, float clip_norm, float grad_scale, void* stream);
int subgnn_inc_step(int* step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
