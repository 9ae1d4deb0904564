// Fused per-step SubGNN path: one CTA per subgraph of the batch, one warp per connected component.
//
// Replaces, per training / inference step (reference):
//   SubGNN.py:609-622                   initialize_cc_embeddings (sum | max pooling of node embeddings)
//   anchor_patch_samplers.py:333-411    get_anchor_patches / embed_anchor_patch (N, P channels: row gathers)
//   subgraph_mpn.py:36-241              SG_MPN.forward: edge list, similarity lookup, msg = sim * x_anchor,
//                                       scatter-add, relu(W [x ; agg] + b), relu(w_p . msg + b_p)
//   SubGNN.py:225-312                   forward: layer / channel loops, concat, masked_sum readout, MLP
//   SubGNN.py:338-342                   CrossEntropy | BCEWithLogits loss
//   and the autograd backward of all of the above.
//
// Closed form used (verified against the reference): for a valid component row r with
// anchors a, mask m = (anchor id != PAD), similarity s:
//     agg[r]   = sum_a m s x_a            cc'[r] = relu(W [cc[r] ; agg[r]] + b)
//     pos[r,a] = relu(m s (w_p . x_a) + b_p)
// Only the outputs that reach the logits are computed: for the P and S channels the updated component
// embeddings are carried by the reference but never concatenated (SubGNN.py:281,291), and for the N channel
// the property-aware output is discarded (SubGNN.py:265-266); their parameters receive no gradient in the
// reference either.  Padded components are skipped: SubGNN.py:303 masks them out of the readout.
// q[a] = w_p . x_a is computed once per shared anchor list (model_q_fwd) instead of once per (component, anchor).
#include "common.cuh"
#include "../../include/subgnn_b200.h"

#define SUB_WARPS 8
#define SUB_THREADS (SUB_WARPS * 32)
#define MAXDPL 8   // D <= 256

typedef subgnn_model_desc Desc;

__device__ __forceinline__ int mpn_blk(const Desc& d) { return 2 * d.D * d.D + 2 * d.D + 1; }
__device__ __forceinline__ int layer_width(const Desc& d) {
  return (d.use_n ? 2 * d.D : 0) + (d.use_p ? d.A_pi + d.A_pb : 0) + (d.use_s ? 2 * d.A_s : 0);
}
__device__ __forceinline__ int col_n(const Desc& d, int l, int side) { return d.D + l * layer_width(d) + side * d.D; }
__device__ __forceinline__ int col_p(const Desc& d, int l, int side) {
  return d.D + l * layer_width(d) + (d.use_n ? 2 * d.D : 0) + (side ? d.A_pi : 0);
}
__device__ __forceinline__ int col_s(const Desc& d, int l, int side) {
  return d.D + l * layer_width(d) + (d.use_n ? 2 * d.D : 0) + (d.use_p ? d.A_pi + d.A_pb : 0) + side * d.A_s;
}

// ------------------------------------------------------------------------------------------------
// prep: batch row offsets, R, max component length; transposed weights
__global__ void __launch_bounds__(256) prep_batch_kernel(Desc d) {
  sg_pdl_sync();
  __shared__ int part[256];
  __shared__ int pmax[256];
  const int t = threadIdx.x;
  const int chunk = (d.B + 255) / 256;
  const int beg = min(d.B, t * chunk), end = min(d.B, beg + chunk);
  int s = 0, mx = 0;
  for (int b = beg; b < end; ++b) {
    const int sub = d.batch_idx[b];
    s += d.sub_ccptr[sub + 1] - d.sub_ccptr[sub];
    mx = max(mx, d.sub_maxlen[sub]);
  }
  part[t] = s;
  pmax[t] = mx;
  __syncthreads();
  if (t == 0) {
    int run = 0, m = 0;
    for (int i = 0; i < 256; ++i) { const int v = part[i]; part[i] = run; run += v; m = max(m, pmax[i]); }
    d.meta[0] = run;
    d.meta[1] = m;
    d.b_rowptr[d.B] = run;
  }
  __syncthreads();
  int run = part[t];
  for (int b = beg; b < end; ++b) {
    const int sub = d.batch_idx[b];
    d.b_rowptr[b] = run;
    const int g0 = d.sub_ccptr[sub], n = d.sub_ccptr[sub + 1] - g0;
    for (int i = 0; i < n; ++i) { d.row_b[run + i] = b; d.row_g[run + i] = g0 + i; }
    run += n;
  }
}

// out[c][r] = in[r][c] for a list of matrices: job j -> (src, dst, rows, cols)
__global__ void transpose_weights_kernel(Desc d) {
  sg_pdl_wait_only();        // the MLP kernels stage the transposed weights before their dependency wait
  // jobs: N-channel projections (L*2 of D x 2D), then lin (h1 x hid), lin2 (h2 x h1), lin3 (K x h2)
  const int n_jobs_n = d.use_n ? d.L * 2 : 0;
  for (int job = blockIdx.y; job < n_jobs_n + 3; job += gridDim.y) {
    const float* src;
    float* dst;
    int rows, cols;
    if (job < n_jobs_n) {
      src = d.mpn_params[0] + (size_t)job * (2 * d.D * d.D + 2 * d.D + 1);
      dst = d.n_wt + (size_t)job * 2 * d.D * d.D;
      rows = d.D; cols = 2 * d.D;
    } else {
      const int k = job - n_jobs_n;
      src = d.lin_w[k];
      dst = d.lin_wt[k];
      rows = k == 0 ? d.h1 : (k == 1 ? d.h2 : d.n_classes);
      cols = k == 0 ? d.hid : (k == 1 ? d.h1 : d.h2);
    }
    const int total = rows * cols;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
      const int c = e / rows, r = e % rows;     // consecutive threads -> consecutive dst addresses (dst[c][r])
      dst[(size_t)c * rows + r] = src[(size_t)r * cols + c];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// q = w_p . x for every shared anchor list.  Entry space per layer l:
//   [0, B*A_pi)            P-internal anchors of the batch's subgraphs
//   [.., + A_pb)           P-border anchors
//   [.., + 2*A_s)          S anchors (internal side, border side): x = LSTM embedding
__device__ __forceinline__ int q_entries_per_layer(const Desc& d) {
  return (d.use_p ? d.B * d.A_pi + d.A_pb : 0) + (d.use_s ? 2 * d.A_s : 0);
}

struct QEntry {
  const float* x;      // anchor row (D floats) or nullptr (PAD)
  const float* wp;     // linear_position.weight of the owning MPN
  float* q;            // destination of q
  float* dq;           // gradient source
  float* dwp;          // gradient of w_p
  float* dx;           // gradient row destination (atomic if !unique)
  bool unique;
};

__device__ __forceinline__ QEntry q_entry(const Desc& d, int l, int e) {
  QEntry r;
  const int blk = 2 * d.D * d.D + 2 * d.D + 1;
  const int wp_off = 2 * d.D * d.D + d.D;
  int id = -1;
  if (d.use_p) {
    if (e < d.B * d.A_pi) {
      const int b = e / d.A_pi, a = e % d.A_pi;
      id = d.p_int_ids[((size_t)l * d.n_sub + d.batch_idx[b]) * d.A_pi + a];
      r.wp = d.mpn_params[1] + (size_t)(2 * l + 0) * blk + wp_off;
      r.dwp = d.mpn_grads[1] ? d.mpn_grads[1] + (size_t)(2 * l + 0) * blk + wp_off : nullptr;
      r.q = d.q_pi + (size_t)l * d.B * d.A_pi + e;
      r.dq = d.dq_pi + (size_t)l * d.B * d.A_pi + e;
    } else if (e < d.B * d.A_pi + d.A_pb) {
      const int a = e - d.B * d.A_pi;
      id = d.p_bor_ids[(size_t)l * d.A_pb + a];
      r.wp = d.mpn_params[1] + (size_t)(2 * l + 1) * blk + wp_off;
      r.dwp = d.mpn_grads[1] ? d.mpn_grads[1] + (size_t)(2 * l + 1) * blk + wp_off : nullptr;
      r.q = d.q_pb + (size_t)l * d.A_pb + a;
      r.dq = d.dq_pb + (size_t)l * d.A_pb + a;
    }
    e -= d.B * d.A_pi + d.A_pb;
  }
  if (id >= 0) {
    r.x = id != 0 ? d.E + (size_t)id * d.D : nullptr;
    r.dx = (id != 0 && d.dE) ? d.dE + (size_t)id * d.D : nullptr;
    r.unique = false;
    return r;
  }
  // structure
  const int side = e / d.A_s, a = e % d.A_s;
  const size_t gi = ((size_t)side * d.L + l) * d.A_s + a;
  r.x = d.emb_s + gi * d.D;
  r.dx = d.d_emb_s ? d.d_emb_s + gi * d.D : nullptr;
  r.unique = true;
  r.wp = d.mpn_params[2] + (size_t)(2 * l + side) * blk + wp_off;
  r.dwp = d.mpn_grads[2] ? d.mpn_grads[2] + (size_t)(2 * l + side) * blk + wp_off : nullptr;
  r.q = d.q_s + ((size_t)l * 2 + side) * d.A_s + a;
  r.dq = d.dq_s + ((size_t)l * 2 + side) * d.A_s + a;
  return r;
}

// which: bit 0 = position-channel entries, bit 1 = structure-channel entries (these need the LSTM output emb_s)
__global__ void __launch_bounds__(256) q_fwd_kernel(Desc d, int which) {
  sg_pdl_sync();
  const int lane = threadIdx.x & 31;
  const int per_layer = q_entries_per_layer(d);
  const int total = per_layer * d.L;
  const int n_pos = d.use_p ? d.B * d.A_pi + d.A_pb : 0;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  for (int w = warp_global; w < total; w += warps_total) {
    const bool is_pos = (w % per_layer) < n_pos;
    if (!((which & 1) && is_pos) && !((which & 2) && !is_pos)) continue;
    const QEntry en = q_entry(d, w / per_layer, w % per_layer);
    float acc = 0.f;
    if (en.x)
      for (int k = lane; k < d.D; k += 32) acc = fmaf(en.wp[k], en.x[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) *en.q = acc;
  }
}

// grid (blocks_per_group, L * groups); group = one (channel side) parameter block, so that d w_p is reduced
// inside the CTA and added once per CTA.
// which: SUBGNN_Q_POS / SUBGNN_Q_STRUC select the groups a launch covers (the structure groups feed the LSTM head gradient: they
// are on the step's critical chain; the position groups only scatter into dE and run beside the BPTT chain)
__global__ void __launch_bounds__(256) q_bwd_kernel(Desc d, int which) {
  sg_pdl_sync();
  __shared__ float s_dwp[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = (d.use_p ? 2 : 0) + (d.use_s ? 2 : 0);
  const int l = blockIdx.y / groups;
  int grp = blockIdx.y % groups;
  {
    const bool is_pos = d.use_p && grp < 2;
    if (is_pos ? !(which & SUBGNN_Q_POS) : !(which & SUBGNN_Q_STRUC)) return;
  }
  int e_beg, e_cnt;
  const int pi = d.B * d.A_pi;
  if (d.use_p && grp == 0) { e_beg = 0; e_cnt = pi; }
  else if (d.use_p && grp == 1) { e_beg = pi; e_cnt = d.A_pb; }
  else {
    if (d.use_p) grp -= 2;
    e_beg = (d.use_p ? pi + d.A_pb : 0) + grp * d.A_s;
    e_cnt = d.A_s;
  }
  float dwp[MAXDPL];
#pragma unroll
  for (int q = 0; q < MAXDPL; ++q) dwp[q] = 0.f;
  float* dwp_dst = nullptr;
  for (int i = blockIdx.x * 8 + warp; i < e_cnt; i += gridDim.x * 8) {
    const QEntry en = q_entry(d, l, e_beg + i);
    dwp_dst = en.dwp;
    const float g = *en.dq;
    if (!en.x) continue;
#pragma unroll
    for (int q = 0; q < MAXDPL; ++q) {
      const int k = lane + 32 * q;
      if (k < d.D) {
        dwp[q] = fmaf(g, en.x[k], dwp[q]);
        if (en.dx) {
          const float v = g * en.wp[k];
          if (en.unique) en.dx[k] = v; else if (v != 0.f) atomicAdd(en.dx + k, v);
        }
      }
    }
  }
  if (!dwp_dst) {   // this warp had no entry: still need the destination for the block reduction
    if (e_cnt > 0) dwp_dst = q_entry(d, l, e_beg).dwp;
  }
#pragma unroll
  for (int q = 0; q < MAXDPL; ++q) if (lane + 32 * q < 256) s_dwp[warp][lane + 32 * q] = dwp[q];
  __syncthreads();
  if (dwp_dst)
    for (int k = threadIdx.x; k < d.D; k += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += s_dwp[w][k];
      if (s != 0.f) atomicAdd(dwp_dst + k, s);
    }
}

// column sets of Z (SUBGNN_COLS_*): the structure-channel property-aware outputs are the only columns that depend on the LSTM
__device__ __forceinline__ bool z_col_selected(const Desc& d, int col, int cols) {
  if (cols == SUBGNN_COLS_ALL) return true;
  bool is_s = false;
  if (d.use_s && col >= d.D) {
    const int lw = layer_width(d);
    is_s = (col - d.D) % lw >= lw - 2 * d.A_s;
  }
  return cols == SUBGNN_COLS_S ? is_s : !is_s;
}

__device__ __forceinline__ unsigned mlp_salt(const Desc& d) { return (d.step_dev ? (unsigned)*d.step_dev : d.step) * 4u; }

// ------------------------------------------------------------------------------------------------
// per-sample MLP backward from external d logits (autograd path); work: shared scratch of at least K + h2 + h1 floats.
// d logits (global) -> dH2, dH1 (pre-activation grads), dZ; dzs (shared, hid floats) receives dZ[b]
__device__ void mlp_backward(const Desc& d, int b, float* dzs, float* work) {
  float* dl = work;                 // K
  float* g2 = dl + d.n_classes;     // h2
  float* g1 = g2 + d.h2;            // h1
  const int tid = threadIdx.x;
  const int K = d.n_classes;
  for (int c = tid; c < K; c += blockDim.x) dl[c] = d.dlogits[(size_t)b * K + c];
  __syncthreads();
  for (int j = tid; j < d.h2; j += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < K; ++c) acc = fmaf(dl[c], d.lin_w[2][(size_t)c * d.h2 + j], acc);
    const float hv = d.H2[(size_t)b * d.h2 + j];
    float sc = d.training ? sg_dropout_scale(d.seed, mlp_salt(d) + 1, (uint64_t)b * d.h2 + j, d.lin_dropout) : 1.f;
    acc = hv > 0.f ? acc * sc : 0.f;
    g2[j] = acc;
    d.dH2[(size_t)b * d.h2 + j] = acc;
  }
  __syncthreads();
  for (int i = tid; i < d.h1; i += blockDim.x) {
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < d.h2; ++j) acc = fmaf(g2[j], __ldg(d.lin_w[1] + (size_t)j * d.h1 + i), acc);
    const float hv = d.H1[(size_t)b * d.h1 + i];
    float sc = d.training ? sg_dropout_scale(d.seed, mlp_salt(d) + 0, (uint64_t)b * d.h1 + i, d.lin_dropout) : 1.f;
    acc = hv > 0.f ? acc * sc : 0.f;
    g1[i] = acc;
    d.dH1[(size_t)b * d.h1 + i] = acc;
  }
  __syncthreads();
  for (int i = tid; i < d.hid; i += blockDim.x) {
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < d.h1; ++j) acc = fmaf(g1[j], __ldg(d.lin_w[0] + (size_t)j * d.hid + i), acc);
    dzs[i] = acc;
    d.dZ[(size_t)b * d.hid + i] = acc;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// ROW kernels: one CTA per connected component (row) of the batch, ROW_THREADS / 32 warps cooperate on it:
// the first half of the warps owns the internal side, the second half the border side of the N channel; inside a side the warps
// split the anchor gathers (8 independent row gathers in flight per warp) and the 2D x D projection.
// The subgraph readout is a global atomicAdd into Z[b] (zeroed every step); the per-sample MLP is its own
// small kernel (grid = B) after all rows are done.
// [r2] two widths, chosen per launch from the row capacity (row_width): 16 warps per row (8 per side) when all rows of the batch fit the
// GPU in one wave — the anchor gathers of a layer are ONE batch of 8 rows in flight per warp instead of two, the 2D-long projection
// reductions are split four ways: the row's dependent chain is shorter — and 8 warps per row when rows outnumber the CTA slots
// (throughput: more rows resident per SM).  Measured on B200, same box, 256 against 512 threads: PPI-BP shape 0.3163 -> 0.3077
// ms/step (100 steps), density 0.166 -> 0.153, HPO-METAB 0.654 -> 0.645; the whole train split as one batch (R ~ 8 k rows):
// 2.03 -> 2.29 ms with 512 — hence the choice by capacity.
#define ROW_WIDE 512
#define ROW_NARROW 256

struct RowSmem {
  float* x0;      // [D]
  float* h;       // [2][D]
  float* in;      // [2][2D]   [h ; agg]
  float* part;    // [2][SIDE_WARPS][D] partial matvec sums
  float* aggl;    // [L][2][D]              similarity-weighted aggregates of every layer
  float* partl;   // [L][2][SIDE_WARPS][D]  their per-warp partials
};

__device__ __forceinline__ RowSmem row_smem(float* sm, int D, int L, int SIDE_WARPS) {
  RowSmem r;
  r.x0 = sm;
  r.h = r.x0 + D;
  r.in = r.h + 2 * D;
  r.part = r.in + 4 * D;
  r.aggl = r.part + 2 * SIDE_WARPS * D;
  r.partl = r.aggl + (size_t)L * 2 * D;
  return r;
}
static size_t row_smem_bytes(int D, int L, int row_threads) {
  const int SIDE_WARPS = row_threads / 64;
  return (size_t)(D + 2 * D + 4 * D + 2 * SIDE_WARPS * D + L * 2 * D + L * 2 * SIDE_WARPS * D + 16) * sizeof(float);
}

// phases: bit 0 = pooling + neighbourhood channel, bit 1 = position / structure property-aware outputs (needs q)
template <int DPL, int RT>
__global__ void __launch_bounds__(RT, RT > 256 ? 2 : (DPL <= 2 ? 4 : 2)) row_fwd_kernel(Desc d, int phases) {   // <= 64 registers (<= 128: narrow, D > 64)
  constexpr int ROW_THREADS = RT, SIDE_THREADS = RT / 2, SIDE_WARPS = RT / 64;
  sg_pdl_sync();
  extern __shared__ float sm[];
  const int D = d.D;
  const RowSmem S = row_smem(sm, D, d.L, SIDE_WARPS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int side = warp / SIDE_WARPS, sw = warp % SIDE_WARPS, st = tid % SIDE_THREADS;
  const int R = d.meta[0], maxlen = d.meta[1];
  const int blk = 2 * D * D + 2 * D + 1;
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const int b = d.row_b[r], g = d.row_g[r];
    const int sub = d.batch_idx[b];
    float* zrow = d.Z + (size_t)b * d.hid;
    __syncthreads();
    // ---- pooling (SubGNN.py:609-622): threads (k, part) split the component's nodes ----
    if (phases & 1) {
      const int nb = d.cc_nodeptr[g], ne = d.cc_nodeptr[g + 1];
      const int parts = D <= ROW_THREADS ? ROW_THREADS / D : 1;
      float* ps = S.part;                                   // reuse: [parts][D] <= 8 * D floats when parts <= 8 ... guard below
      const int pmax = 2 * SIDE_WARPS;                      // capacity of S.part in units of D
      const int np = parts < pmax ? parts : pmax;
      if (tid < np * D) {
        const int k = tid % D, part = tid / D;
        float acc = d.pool_max ? -INFINITY : 0.f;
        for (int i = nb + part; i < ne; i += np) {
          const float v = d.E[(size_t)d.cc_nodes[i] * D + k];
          acc = d.pool_max ? fmaxf(acc, v) : acc + v;
        }
        ps[part * D + k] = acc;
      }
      __syncthreads();
      for (int k = tid; k < D; k += ROW_THREADS) {
        float acc = ps[k];
        for (int p_ = 1; p_ < np; ++p_) acc = d.pool_max ? fmaxf(acc, ps[p_ * D + k]) : acc + ps[p_ * D + k];
        if (d.pool_max && (ne - nb) < maxlen) acc = fmaxf(acc, 0.f);     // PAD rows (zero vectors) take part in the max
        S.x0[k] = acc;
        d.X0[(size_t)r * D + k] = acc;
        atomicAdd(zrow + k, acc);
      }
      __syncthreads();
    }
    // ---- neighbourhood channel ----
    if (d.use_n && (phases & 1)) {
      const int A = side ? d.A_nb : d.A_ni;
      for (int k = st; k < D; k += SIDE_THREADS) {
        const float v = d.trainable_cc ? d.cc_tab[side][((size_t)sub * d.C_pad + (g - d.sub_ccptr[sub])) * D + k] : S.x0[k];
        S.h[side * D + k] = v;
        d.Nh[((size_t)(0 * 2 + side) * d.R_cap + r) * D + k] = v;
      }
      // ---- gather-scale-reduce of EVERY layer first: agg_l = sum_a s_a E[id_a] depends on the anchors, the similarities and the
      // embedding table only, not on the previous layer's output, so the L x 2 gathers run back to back (8 rows in flight per
      // warp, ids / similarities of a warp's anchors fetched with one coalesced load and handed out by shuffles) and only the
      // L small projections remain a dependent chain.  Summation order per warp is unchanged (anchors in increasing order).
      for (int l = 0; l < d.L; ++l) {
        const int* ids = d.n_ids[side] + ((size_t)l * d.n_cc + g) * A;
        const float* sims = d.n_sim[side] + ((size_t)l * d.n_cc + g) * A;
        float agg[DPL];
#pragma unroll
        for (int q = 0; q < DPL; ++q) agg[q] = 0.f;
        const int n_w = A > sw ? (A - sw + SIDE_WARPS - 1) / SIDE_WARPS : 0;       // anchors sw, sw + SIDE_WARPS, ... of this warp
        for (int base = 0; base < n_w; base += 32) {
          const int mine = base + lane;
          const int id_l = mine < n_w ? ids[sw + mine * SIDE_WARPS] : 0;
          const float s_l = mine < n_w ? sims[sw + mine * SIDE_WARPS] : 0.f;
          const int cnt = min(32, n_w - base);
          for (int i = 0; i < cnt; i += 8) {
            int id8[8];
            float s8[8], v8[8][DPL];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              id8[u] = __shfl_sync(0xffffffffu, id_l, (i + u) & 31);
              s8[u] = __shfl_sync(0xffffffffu, s_l, (i + u) & 31);
              if (i + u >= cnt) id8[u] = 0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
              for (int q = 0; q < DPL; ++q) {
                const int k = lane + 32 * q;
                v8[u][q] = (id8[u] && k < D) ? d.E[(size_t)id8[u] * D + k] : 0.f;
              }
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (i + u < cnt)
#pragma unroll
                for (int q = 0; q < DPL; ++q) agg[q] = fmaf(s8[u], v8[u][q], agg[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < DPL; ++q) {
          const int k = lane + 32 * q;
          if (k < D) S.partl[((size_t)(l * 2 + side) * SIDE_WARPS + sw) * D + k] = agg[q];
        }
      }
      __syncthreads();
      for (int idx = st; idx < d.L * D; idx += SIDE_THREADS) {
        const int l = idx / D, k = idx % D;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < SIDE_WARPS; ++w) v += S.partl[((size_t)(l * 2 + side) * SIDE_WARPS + w) * D + k];
        S.aggl[(size_t)(l * 2 + side) * D + k] = v;
        d.Nagg[((size_t)(l * 2 + side) * d.R_cap + r) * D + k] = v;
      }
      __syncthreads();
      for (int l = 0; l < d.L; ++l) {
        for (int k = st; k < D; k += SIDE_THREADS) {
          S.in[side * 2 * D + k] = S.h[side * D + k];
          S.in[side * 2 * D + D + k] = S.aggl[(size_t)(l * 2 + side) * D + k];
        }
        __syncthreads();
        if (d.use_proj) {
          // out[j] = b[j] + sum_kk in[kk] * WT[kk][j]; (j, part) split of the 2D-long reduction over the side's threads
          const float* wt = d.n_wt + (size_t)(2 * l + side) * 2 * D * D;
          const int parts = (D <= SIDE_THREADS && SIDE_THREADS % D == 0) ? min(SIDE_THREADS / D, SIDE_WARPS) : 1;
          const float* inp = S.in + side * 2 * D;
          const float* bias = d.mpn_params[0] + (size_t)(2 * l + side) * blk + 2 * D * D;
          const float bias_j = st < D ? __ldg(bias + st) : 0.f;          // issued with the weight loads, used after the barrier
          for (int idx = st; idx < parts * D; idx += SIDE_THREADS) {
            const int j = idx % D, part = idx / D;
            const int k0 = 2 * D * part / parts, k1 = 2 * D * (part + 1) / parts;
            float acc = 0.f;
#pragma unroll 32
            for (int kk = k0; kk < k1; ++kk) acc = fmaf(inp[kk], __ldg(wt + (size_t)kk * D + j), acc);   // 32 weight loads in flight
            S.part[(side * SIDE_WARPS + part) * D + j] = acc;
          }
          __syncthreads();
          for (int j = st; j < D; j += SIDE_THREADS) {
            float acc = j == st ? bias_j : bias[j];
            for (int p_ = 0; p_ < parts; ++p_) acc += S.part[(side * SIDE_WARPS + p_) * D + j];
            S.h[side * D + j] = fmaxf(acc, 0.f);
          }
        } else {
          for (int j = st; j < D; j += SIDE_THREADS) S.h[side * D + j] = S.in[side * 2 * D + D + j];   // subgraph_mpn.py:240-241
        }
        __syncthreads();
        const int zc = col_n(d, l, side);
        for (int j = st; j < D; j += SIDE_THREADS) {
          const float v = S.h[side * D + j];
          d.Nh[((size_t)((l + 1) * 2 + side) * d.R_cap + r) * D + j] = v;
          atomicAdd(zrow + zc + j, v);
        }
      }
    }
    // ---- position / structure channels: property-aware outputs relu(s q + b_p) ----
    if (phases & SUBGNN_PHASE_PS) {
      const int wp_ = (d.use_p ? d.A_pi + d.A_pb : 0), ws_ = (d.use_s ? 2 * d.A_s : 0);
      const int per_layer = wp_ + ws_;
      for (int e = tid; e < d.L * per_layer; e += ROW_THREADS) {
        const int l = e / per_layer;
        int o = e % per_layer;
        if (!(phases & (o < wp_ ? SUBGNN_PHASE_P : SUBGNN_PHASE_S))) continue;
        float s, q, bp;
        int zc;
        if (o < wp_) {
          const int sd = o >= d.A_pi ? 1 : 0, a = sd ? o - d.A_pi : o, A = sd ? d.A_pb : d.A_pi;
          s = d.p_sim[sd][((size_t)l * d.n_cc + g) * A + a];
          q = sd ? d.q_pb[(size_t)l * A + a] : d.q_pi[((size_t)l * d.B + b) * A + a];
          bp = d.mpn_params[1][(size_t)(2 * l + sd) * blk + 2 * D * D + 2 * D];
          zc = col_p(d, l, sd) + a;
        } else {
          o -= wp_;
          const int sd = o / d.A_s, a = o % d.A_s;
          s = d.s_sim[sd][((size_t)l * d.n_cc + g) * d.A_s + a];
          q = d.q_s[((size_t)l * 2 + sd) * d.A_s + a];
          bp = d.mpn_params[2][(size_t)(2 * l + sd) * blk + 2 * D * D + 2 * D];
          zc = col_s(d, l, sd) + a;
        }
        const float v = fmaxf(fmaf(s, q, bp), 0.f);
        if (v != 0.f) atomicAdd(zrow + zc, v);
      }
    }
  }
}

// per-sample MLP + loss (+ MLP backward when training): grid = B
// 1024 threads per sample: the hid-long reductions of lin (and the hid-wide d Z) are split over 1024 / h1 thread groups so that
// every thread's dependent load -> FMA chain is short (the 256-thread version was L2-latency bound: 87 us for 32 samples).
#define MLP_THREADS 1024
__global__ void __launch_bounds__(MLP_THREADS) mlp_bwd_kernel(Desc d) {
  extern __shared__ float sm[];
  mlp_backward(d, blockIdx.x, sm, sm + d.hid);
}

// ------------------------------------------------------------------------------------------------
// Readout MLP as three batch-level kernels (SubGNN.py:306-310 lin -> lin2 -> lin3, subgraph_utils / loss :338-343).  The
// per-sample kernel above reads all of W1 (hid x h1, ~0.45 MB) once PER SAMPLE in ~14 dependent L2 round trips; here W1 is
// read once per step, K-sliced over CTAs, and every phase issues its loads together:
//   mlp_lin1_kernel  CTA per 64-row slice of hid: H1pre[b][j] += sum_{k in slice} Z[b][k] W1t[k][j]   (H1 zeroed per step)
//   mlp_rest_kernel  CTA per sample: bias / relu / dropout, lin2, lin3, loss, d logits, dH2, dH1       (W2 staged in smem)
//   mlp_dz_kernel    CTA per 64-column slice of hid: dZ[b][i] = sum_j dH1[b][j] W1[j][i]
#define MLP_SLICE 64
#define MLP_BT 128      // samples per CTA of the two sliced kernels (grid.y tiles larger batches)
__global__ void __launch_bounds__(256) mlp_lin1_kernel(Desc d, int cols) {
  extern __shared__ float sm[];
  const int b0 = blockIdx.y * MLP_BT, bn = min(MLP_BT, d.B - b0);
  float* zs = sm;                               // [bn][MLP_SLICE]
  float* ws = sm + (size_t)min(d.B, MLP_BT) * MLP_SLICE;     // [MLP_SLICE][h1]
  const int k0 = blockIdx.x * MLP_SLICE, kn = min(MLP_SLICE, d.hid - k0);
  if (cols != SUBGNN_COLS_ALL) {                // slices without a column of the requested set retire at once
    int any = 0;
    for (int kk = threadIdx.x; kk < kn; kk += blockDim.x) any |= z_col_selected(d, k0 + kk, cols) ? 1 : 0;
    if (!__syncthreads_or(any)) return;
  }
  // the weight slice was transposed several launches ago (complete by transitivity of the dependency waits): staged before
  // this kernel's own wait, while the row kernel that finishes Z is still running
  const float* wt = d.lin_wt[0] + (size_t)k0 * d.h1;
  sg_stage<16>(ws, kn * d.h1, [&](int e) { return __ldg(wt + e); });
  sg_pdl_sync();
  sg_stage<8>(zs, bn * MLP_SLICE, [&](int e) {
    const int b = e / MLP_SLICE, kk = e % MLP_SLICE;
    return (kk < kn && z_col_selected(d, k0 + kk, cols)) ? d.Z[(size_t)(b0 + b) * d.hid + k0 + kk] : 0.f;
  });
  __syncthreads();
  for (int o = threadIdx.x; o < bn * d.h1; o += blockDim.x) {
    const int b = o / d.h1, j = o % d.h1;
    const float* zr = zs + (size_t)b * MLP_SLICE;
    float acc = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < kn; ++kk) acc = fmaf(zr[kk], ws[kk * d.h1 + j], acc);
    atomicAdd(d.H1 + (size_t)b0 * d.h1 + o, acc);
  }
}

__global__ void __launch_bounds__(256) mlp_rest_kernel(Desc d) {
  extern __shared__ float sm[];
  float* w1t = sm;                               // [h1][h2]  lin2 weights, input-major (forward)
  float* w1n = w1t + (size_t)d.h1 * d.h2;        // [h2][h1]  lin2 weights, native (backward)
  float* h1s = w1n + (size_t)d.h1 * d.h2;        // [h1]
  float* h2s = h1s + d.h1;                       // [h2]
  float* lg = h2s + d.h2;                        // [K]
  float* dl = lg + d.n_classes;                  // [K]
  float* g2 = dl + d.n_classes;                  // [h2]
  float* bs = g2 + d.h2;                         // [h1 + h2 + K]  the three bias vectors
  float* w3s = bs + d.h1 + d.h2 + d.n_classes;   // [K][h2]        lin3 weights
  const int b = blockIdx.x, tid = threadIdx.x, K = d.n_classes;
  const bool bwd = d.training && d.dZ;
  // both copies of the lin2 weights land in one staging pass (w1n directly follows w1t): 16 loads in flight per thread
  const float* src_t = d.lin_wt[1];
  const float* src_n = d.lin_w[1];
  const int nw1 = d.h1 * d.h2;
  sg_stage<16>(w1t, bwd ? 2 * nw1 : nw1, [&](int e) { return e < nw1 ? __ldg(src_t + e) : __ldg(src_n + e - nw1); });
  // every other global operand of the dependent phases below (biases, lin3 weights, the label) is fetched before the wait too:
  // each L2 round trip inside the chain is ~1 us of the step's critical path
  sg_stage<8>(bs, d.h1 + d.h2 + K + K * d.h2, [&](int e) {
    if (e < d.h1) return __ldg(d.lin_b[0] + e);
    if (e < d.h1 + d.h2) return __ldg(d.lin_b[1] + e - d.h1);
    if (e < d.h1 + d.h2 + K) return __ldg(d.lin_b[2] + e - d.h1 - d.h2);
    return __ldg(d.lin_w[2] + e - d.h1 - d.h2 - K);
  });
  int my_label = 0;                              // (batch_idx is written by the step's H2D copy before any kernel; labels are static)
  if (tid == 0 && !d.multilabel && d.labels) my_label = d.labels[d.batch_idx[b]];
  sg_pdl_sync();                                 // (weights staged before the wait: not written by the preceding launch)
  for (int j = tid; j < d.h1; j += blockDim.x) {
    float acc = fmaxf(d.H1[(size_t)b * d.h1 + j] + bs[j], 0.f);
    if (d.training) acc *= sg_dropout_scale(d.seed, mlp_salt(d) + 0, (uint64_t)b * d.h1 + j, d.lin_dropout);
    h1s[j] = acc;
    d.H1[(size_t)b * d.h1 + j] = acc;
  }
  __syncthreads();
  for (int j = tid; j < d.h2; j += blockDim.x) {
    float acc = bs[d.h1 + j];
#pragma unroll 8
    for (int i = 0; i < d.h1; ++i) acc = fmaf(h1s[i], w1t[i * d.h2 + j], acc);
    acc = fmaxf(acc, 0.f);
    if (d.training) acc *= sg_dropout_scale(d.seed, mlp_salt(d) + 1, (uint64_t)b * d.h2 + j, d.lin_dropout);
    h2s[j] = acc;
    d.H2[(size_t)b * d.h2 + j] = acc;
  }
  __syncthreads();
  {                                              // lin3: warp per class, lanes along h2
    const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    for (int c = warp; c < K; c += nw) {
      float acc = 0.f;
      for (int i = lane; i < d.h2; i += 32) acc = fmaf(h2s[i], w3s[c * d.h2 + i], acc);
      acc = warp_sum(acc);
      if (lane == 0) {
        acc += bs[d.h1 + d.h2 + c];
        lg[c] = acc;
        d.logits[(size_t)b * K + c] = acc;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {                                // loss term and d logits (K is tiny)
    float loss = 0.f;
    if (!d.multilabel) {
      float mx = lg[0];
      for (int c = 1; c < K; ++c) mx = fmaxf(mx, lg[c]);
      float se = 0.f;
      for (int c = 0; c < K; ++c) se += expf(lg[c] - mx);
      const float lse = mx + logf(se);
      const int y = my_label;
      loss = (lse - lg[y]) / (float)d.B;
      for (int c = 0; c < K; ++c) {
        dl[c] = (expf(lg[c] - lse) - (c == y ? 1.f : 0.f)) / (float)d.B;
        if (d.dlogits) d.dlogits[(size_t)b * K + c] = dl[c];
      }
    } else {
      const int sub = d.batch_idx[b];
      const float inv = 1.f / ((float)d.B * (float)K);
      for (int c = 0; c < K; ++c) {
        const float x = lg[c], y = d.labels_multi[(size_t)sub * K + c];
        loss += (fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)))) * inv;
        dl[c] = (1.f / (1.f + expf(-x)) - y) * inv;
        if (d.dlogits) d.dlogits[(size_t)b * K + c] = dl[c];
      }
    }
    d.loss_b[b] = loss;
  }
  if (!bwd) return;
  __syncthreads();
  for (int j = tid; j < d.h2; j += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < K; ++c) acc = fmaf(dl[c], w3s[c * d.h2 + j], acc);
    const float sc = sg_dropout_scale(d.seed, mlp_salt(d) + 1, (uint64_t)b * d.h2 + j, d.lin_dropout);
    acc = h2s[j] > 0.f ? acc * sc : 0.f;
    g2[j] = acc;
    d.dH2[(size_t)b * d.h2 + j] = acc;
  }
  __syncthreads();
  for (int i = tid; i < d.h1; i += blockDim.x) {
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < d.h2; ++j) acc = fmaf(g2[j], w1n[j * d.h1 + i], acc);
    const float sc = sg_dropout_scale(d.seed, mlp_salt(d) + 0, (uint64_t)b * d.h1 + i, d.lin_dropout);
    d.dH1[(size_t)b * d.h1 + i] = h1s[i] > 0.f ? acc * sc : 0.f;
  }
}

__global__ void __launch_bounds__(256) mlp_dz_kernel(Desc d, int cols) {
  extern __shared__ float sm[];
  const int b0 = blockIdx.y * MLP_BT, bn = min(MLP_BT, d.B - b0);
  float* g1s = sm;                               // [bn][h1]
  float* ws = sm + (size_t)min(d.B, MLP_BT) * d.h1;           // [h1][MLP_SLICE]
  const int i0 = blockIdx.x * MLP_SLICE, in = min(MLP_SLICE, d.hid - i0);
  if (cols != SUBGNN_COLS_ALL) {
    int any = 0;
    for (int ii = threadIdx.x; ii < in; ii += blockDim.x) any |= z_col_selected(d, i0 + ii, cols) ? 1 : 0;
    if (!__syncthreads_or(any)) return;
  }
  // the W1 slice is a PARAMETER (last written by the previous step's Adam, a full stream dependency ago): staged before the
  // programmatic-dependency wait, i.e. while mlp_rest_kernel is still running
  const float* w0 = d.lin_w[0];
  sg_stage<16>(ws, d.h1 * MLP_SLICE, [&](int e) {
    const int j = e / MLP_SLICE, ii = e % MLP_SLICE;
    return ii < in ? __ldg(w0 + (size_t)j * d.hid + i0 + ii) : 0.f;
  });
  sg_pdl_sync();
  sg_stage<8>(g1s, bn * d.h1, [&](int e) { return d.dH1[(size_t)b0 * d.h1 + e]; });
  __syncthreads();
  for (int o = threadIdx.x; o < bn * MLP_SLICE; o += blockDim.x) {
    const int b = o / MLP_SLICE, ii = o % MLP_SLICE;
    if (ii >= in || !z_col_selected(d, i0 + ii, cols)) continue;
    const float* gr = g1s + (size_t)b * d.h1;
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < d.h1; ++j) acc = fmaf(gr[j], ws[j * MLP_SLICE + ii], acc);
    d.dZ[(size_t)(b0 + b) * d.hid + i0 + ii] = acc;
  }
}


// ------------------------------------------------------------------------------------------------
// Readout section as one cluster kernel (SubGNN.py:303-310, :338-342 and their backward).  Cluster = RO_CL CTAs = RO_BT samples
// (one sample per CTA behind the first layer); a batch of B samples runs on B / RO_BT clusters = B CTAs:
//   every CTA owns a slice of S = ceil(hid / RO_CL) columns of Z and of W1 (staged ONCE, used by the forward product, by dZ and
//   by dW1).
//   1  partial[b][j] = sum_{i in slice} Z[b][i] W1[j][i]                 -> sent to the owner CTA of sample b (DSMEM)
//   2  owner: H1 = drop(relu(sum of the RO_CL partials + b1)), H2, logits, loss, d logits, dH2, dH1
//      dH1 row broadcast to every CTA of the cluster (DSMEM); mlp_fused: dW2, dW3, db1-3 from the owner's sample
//   3  dZ[b][slice] = dH1[b] . W1[:, slice];   mlp_fused: dW1[:, slice] += dH1^T Z[:, slice]
// (first version: 32 samples per cluster, i.e. 8 CTAs at the reference batch of 32 — 45 us against 30 us for the three
// batch-level kernels: too few SMs for the 0.45 MB of W1; one sample per CTA puts B CTAs on the section)
#ifdef WS_TRACE
__device__ long long ro_trace_buf[64];
#define RO_TR(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) ro_trace_buf[i] = clock64(); } while (0)
extern "C" int subgnn_ro_trace_read(long long* host_buf) { cudaDeviceSynchronize(); return (int)cudaMemcpyFromSymbol(host_buf, ro_trace_buf, sizeof(long long) * 64); }
#else
#define RO_TR(i) do { } while (0)
#endif
#define RO_CL 8
#define RO_BT 8
#define RO_THREADS 256
#define RO_IQ 4            // phase 1: the slice is cut in RO_IQ quarters over thread groups

// 1-D TMA bulk copy global -> shared completing on an mbarrier (one issuing thread); chunks of <= 32 KB
__device__ __forceinline__ void ro_bulk_copy(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(reinterpret_cast<const char*>(dst) + off)), "l"(reinterpret_cast<const char*>(src) + off), "r"(n),
                    "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
  }
}

struct RoSmem {
  float *w1t, *zs, *w2t, *w2n, *recv, *gall, *part, *h1s, *h2s, *g2s, *g1s, *lgs, *dls, *bs, *w3s;
  int S, ld1;
};
__host__ __device__ inline int ro_slice(int hid) { return ((hid + RO_CL - 1) / RO_CL + 3) / 4 * 4; }
__host__ __device__ inline size_t ro_smem_floats(int hid, int h1, int h2, int K) {
  const int S = ro_slice(hid);
  const size_t w2 = ((size_t)h1 * h2 + 3) / 4 * 4;
  return (size_t)S * h1 + 2 * w2 + (size_t)RO_BT * S + (size_t)h1 * RO_BT + (size_t)RO_CL * h1 + (size_t)RO_IQ * RO_BT * h1 +
         (size_t)(2 * h1 + 2 * h2 + 2 * K) + (size_t)(h1 + h2 + K + K * h2) + 32;
}
__device__ __forceinline__ RoSmem ro_carve(float* sm, const Desc& d) {
  RoSmem r;
  r.S = ro_slice(d.hid);
  r.ld1 = d.h1;
  const size_t w2 = ((size_t)d.h1 * d.h2 + 3) / 4 * 4;
  r.w1t = sm;                                             // [S][h1]         W1^T slice (bulk-copy destination: 16-byte aligned)
  r.w2t = r.w1t + (size_t)r.S * d.h1;                     // [h1][h2]        (S % 4 == 0 keeps the alignment)
  r.w2n = r.w2t + w2;                                     // [h2][h1]
  r.zs = r.w2n + w2;                                      // [RO_BT][S]      (16-byte aligned rows: S % 4 == 0)
  r.gall = r.zs + (size_t)RO_BT * r.S;                    // [h1][RO_BT]     dH1 of the cluster's samples, sample index contiguous
  r.recv = r.gall + (size_t)d.h1 * RO_BT;                 // [RO_CL][h1]     partial first-layer sums of MY sample from every CTA
  r.part = r.recv + (size_t)RO_CL * d.h1;                 // [RO_IQ][RO_BT][h1]
  r.h1s = r.part + (size_t)RO_IQ * RO_BT * d.h1;          // [h1]
  r.g1s = r.h1s + d.h1;                                   // [h1]
  r.h2s = r.g1s + d.h1;                                   // [h2]
  r.g2s = r.h2s + d.h2;                                   // [h2]
  r.lgs = r.g2s + d.h2;                                   // [K]
  r.dls = r.lgs + d.n_classes;                            // [K]
  r.bs = r.dls + d.n_classes;                             // [h1 + h2 + K]   the three bias vectors
  r.w3s = r.bs + d.h1 + d.h2 + d.n_classes;               // [K][h2]         lin3 weights
  return r;
}

__global__ void __launch_bounds__(RO_THREADS) readout_cluster_kernel(Desc d) {
  extern __shared__ __align__(16) float sm[];
  const RoSmem R = ro_carve(sm, d);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned rank = sg_cluster_rank();
  const int cl = blockIdx.x / RO_CL;
  const int b0 = cl * RO_BT, bn = min(RO_BT, d.B - b0);
  const int S = R.S, ld1 = R.ld1, h1 = d.h1, h2 = d.h2, K = d.n_classes;
  const int i0 = (int)rank * S, Sr = max(0, min(S, d.hid - i0));      // my columns [i0, i0 + Sr)
  const bool bwd = d.training && d.dZ;
  RO_TR(0);
  // ---- weights: staged before the dependency wait (parameters / derived copies whose producers never trigger early): the W1^T
  // slice (contiguous rows i0 .. of the transposed copy, ~56 KB) and both orientations of W2 arrive by 1-D TMA bulk copies ----
  __shared__ uint64_t wbar;
  {
    const float* wt = d.lin_wt[0] + (size_t)i0 * h1;
    const int nw = h1 * h2;
    const bool bulk = (h1 % 4) == 0 && (nw % 4) == 0 && ((((size_t)wt) | ((size_t)d.lin_wt[1]) | ((size_t)d.lin_w[1])) & 15) == 0;
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&wbar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const uint32_t bytes = bulk ? (uint32_t)(((size_t)Sr * h1 + (bwd ? 2 : 1) * (size_t)nw) * sizeof(float)) : 0u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&wbar)), "r"(bytes) : "memory");
      if (bulk) {
        if (Sr > 0) ro_bulk_copy(R.w1t, wt, (uint32_t)((size_t)Sr * h1 * sizeof(float)), &wbar);
        ro_bulk_copy(R.w2t, d.lin_wt[1], (uint32_t)(nw * sizeof(float)), &wbar);
        if (bwd) ro_bulk_copy(R.w2n, d.lin_w[1], (uint32_t)(nw * sizeof(float)), &wbar);
      }
    }
    for (int e = Sr * h1 + tid; e < S * h1; e += RO_THREADS) R.w1t[e] = 0.f;          // rows beyond the matrix (last slice)
    if (!bulk) {
      for (int e = tid; e < Sr * h1; e += RO_THREADS) R.w1t[e] = __ldg(wt + e);
      for (int e = tid; e < nw; e += RO_THREADS) { R.w2t[e] = __ldg(d.lin_wt[1] + e); if (bwd) R.w2n[e] = __ldg(d.lin_w[1] + e); }
    }
    // everything the chain behind the first layer reads from global memory — biases, lin3 weights, the label — is fetched here,
    // before the wait: that chain is ~8 dependent phases and each L2 round trip inside it costs ~1 us of the step's critical path
    sg_stage<8>(R.bs, h1 + h2 + K + K * h2, [&](int e) {
      if (e < h1) return __ldg(d.lin_b[0] + e);
      if (e < h1 + h2) return __ldg(d.lin_b[1] + e - h1);
      if (e < h1 + h2 + K) return __ldg(d.lin_b[2] + e - h1 - h2);
      return __ldg(d.lin_w[2] + e - h1 - h2 - K);
    });
  }
  // (batch_idx is written by the step's H2D copy before any kernel of the step; labels are static)
  const int my_b = blockIdx.x / RO_CL * RO_BT + (int)rank;
  int my_label = 0;
  if (tid == 0 && my_b < d.B && !d.multilabel && d.labels) my_label = d.labels[d.batch_idx[my_b]];
  RO_TR(1);
  sg_pdl_sync();
  RO_TR(2);
  // ---- Z slice ----
  {
    const bool v4 = (d.hid % 4) == 0 && ((((size_t)d.Z) & 15) == 0);    // i0 and S are multiples of 4
    if (v4) {
      const int S4 = S / 4;
      for (int e = tid; e < RO_BT * S4; e += RO_THREADS) {
        const int b = e / S4, i = (e % S4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < bn && i < Sr) {
          const float* src = d.Z + (size_t)(b0 + b) * d.hid + i0 + i;
          if (i + 3 < Sr) v = *reinterpret_cast<const float4*>(src);
          else { v.x = src[0]; if (i + 1 < Sr) v.y = src[1]; if (i + 2 < Sr) v.z = src[2]; }
        }
        *reinterpret_cast<float4*>(R.zs + (size_t)b * S + i) = v;
      }
    } else {
      for (int e = tid; e < RO_BT * S; e += RO_THREADS) {
        const int b = e / S, i = e % S;
        R.zs[e] = (b < bn && i < Sr) ? d.Z[(size_t)(b0 + b) * d.hid + i0 + i] : 0.f;
      }
    }
  }
  __syncthreads();
  {                                                                    // bulk copies landed (phase 0 of the barrier)
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&wbar);
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar) : "memory");
  }
  RO_TR(3);
  // ---- 1: partial first-layer sums of my slice: work item = (unit j, quarter q of the slice), all RO_BT samples ----
  {
    const int S4q = ((S / 4 + RO_IQ - 1) / RO_IQ) * 4;                 // columns per quarter (multiple of 4)
    for (int it = tid; it < h1 * RO_IQ; it += RO_THREADS) {
      const int j = it % h1, q = it / h1;
      const int ia = q * S4q, ib = min(S, ia + S4q);
      float acc[RO_BT];
#pragma unroll
      for (int u = 0; u < RO_BT; ++u) acc[u] = 0.f;
#pragma unroll 2
      for (int i = ia; i < ib; i += 4) {
        const float w0 = R.w1t[(i + 0) * ld1 + j], w1 = R.w1t[(i + 1) * ld1 + j], w2 = R.w1t[(i + 2) * ld1 + j], w3 = R.w1t[(i + 3) * ld1 + j];
#pragma unroll
        for (int u = 0; u < RO_BT; ++u) {
          const float4 z = *reinterpret_cast<const float4*>(R.zs + (size_t)u * S + i);
          acc[u] = fmaf(z.x, w0, fmaf(z.y, w1, fmaf(z.z, w2, fmaf(z.w, w3, acc[u]))));
        }
      }
#pragma unroll
      for (int u = 0; u < RO_BT; ++u) R.part[((size_t)q * RO_BT + u) * h1 + j] = acc[u];
    }
    __syncthreads();
    for (int o = tid; o < RO_BT * h1; o += RO_THREADS) {               // quarter sums -> the owner CTA of sample b = o / h1
      const int b = o / h1, j = o % h1;
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < RO_IQ; ++q) v += R.part[((size_t)q * RO_BT + b) * h1 + j];
      sg_st_cluster(R.recv + (size_t)rank * h1 + j, (unsigned)b, v);
    }
  }
  RO_TR(4);
  sg_cluster_sync();
  RO_TR(5);
  // ---- 2: my sample behind the first layer ----
  const int b = b0 + (int)rank;                                        // the sample I own
  const bool live = b < d.B;
  for (int j = tid; j < h1; j += RO_THREADS) {
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < RO_CL; ++r) acc += R.recv[(size_t)r * h1 + j];
    acc = fmaxf(acc + R.bs[j], 0.f);
    if (d.training) acc *= sg_dropout_scale(d.seed, mlp_salt(d) + 0, (uint64_t)b * h1 + j, d.lin_dropout);
    R.h1s[j] = acc;
    if (live) d.H1[(size_t)b * h1 + j] = acc;
  }
  __syncthreads();
  for (int j = tid; j < h2; j += RO_THREADS) {
    float acc = R.bs[h1 + j];
#pragma unroll 8
    for (int i = 0; i < h1; ++i) acc = fmaf(R.h1s[i], R.w2t[i * h2 + j], acc);
    acc = fmaxf(acc, 0.f);
    if (d.training) acc *= sg_dropout_scale(d.seed, mlp_salt(d) + 1, (uint64_t)b * h2 + j, d.lin_dropout);
    R.h2s[j] = acc;
    if (live) d.H2[(size_t)b * h2 + j] = acc;
  }
  __syncthreads();
  for (int c = warp; c < K; c += RO_THREADS / 32) {                     // lin3: warp per class
    float acc = 0.f;
    for (int i = lane; i < h2; i += 32) acc = fmaf(R.h2s[i], R.w3s[c * h2 + i], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += R.bs[h1 + h2 + c];
      R.lgs[c] = acc;
      if (live) d.logits[(size_t)b * K + c] = acc;
    }
  }
  __syncthreads();
  if (tid == 0) {
    float loss = 0.f;
    if (live) {
      const float* lg = R.lgs;
      const int sub = d.multilabel ? d.batch_idx[b] : 0;
      if (!d.multilabel) {
        float mx = lg[0];
        for (int c = 1; c < K; ++c) mx = fmaxf(mx, lg[c]);
        float se = 0.f;
        for (int c = 0; c < K; ++c) se += expf(lg[c] - mx);
        const float lse = mx + logf(se);
        const int y = my_label;
        loss = (lse - lg[y]) / (float)d.B;
        for (int c = 0; c < K; ++c) R.dls[c] = (expf(lg[c] - lse) - (c == y ? 1.f : 0.f)) / (float)d.B;
      } else {
        const float inv = 1.f / ((float)d.B * (float)K);
        for (int c = 0; c < K; ++c) {
          const float x = lg[c], y = d.labels_multi[(size_t)sub * K + c];
          loss += (fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)))) * inv;
          R.dls[c] = (1.f / (1.f + expf(-x)) - y) * inv;
        }
      }
      if (d.dlogits) for (int c = 0; c < K; ++c) d.dlogits[(size_t)b * K + c] = R.dls[c];
      d.loss_b[b] = loss;
      if (d.loss_sum) atomicAdd(d.loss_sum, loss);
    } else {
      for (int c = 0; c < K; ++c) R.dls[c] = 0.f;
    }
  }
  if (!bwd) return;                                                     // (uniform over the cluster: no CTA waits at barrier 2)
  __syncthreads();
  RO_TR(6);
  for (int j = tid; j < h2; j += RO_THREADS) {
    float acc = 0.f;
    for (int c = 0; c < K; ++c) acc = fmaf(R.dls[c], R.w3s[c * h2 + j], acc);
    const float sc = sg_dropout_scale(d.seed, mlp_salt(d) + 1, (uint64_t)b * h2 + j, d.lin_dropout);
    acc = (live && R.h2s[j] > 0.f) ? acc * sc : 0.f;
    if (live) d.dH2[(size_t)b * h2 + j] = acc;
    R.g2s[j] = acc;
  }
  __syncthreads();
  for (int i = tid; i < h1; i += RO_THREADS) {
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < h2; ++j) acc = fmaf(R.g2s[j], R.w2n[j * h1 + i], acc);
    const float sc = sg_dropout_scale(d.seed, mlp_salt(d) + 0, (uint64_t)b * h1 + i, d.lin_dropout);
    acc = (live && R.h1s[i] > 0.f) ? acc * sc : 0.f;
    if (live) d.dH1[(size_t)b * h1 + i] = acc;
    R.g1s[i] = acc;
#pragma unroll
    for (unsigned r = 0; r < RO_CL; ++r) sg_st_cluster(R.gall + (size_t)i * RO_BT + rank, r, acc);
  }
  __syncthreads();
  if (d.mlp_fused && d.lin_gw[0] && live) {                             // small gradients from my sample (atomics over the B CTAs)
    for (int o = tid; o < h2 * h1; o += RO_THREADS) {                   // dW2[j][i] += g2[j] h1[i]
      const float v = R.g2s[o / h1] * R.h1s[o % h1];
      if (v != 0.f) atomicAdd(d.lin_gw[1] + o, v);
    }
    for (int o = tid; o < K * h2; o += RO_THREADS) {                    // dW3[c][j] += dl[c] h2[j]
      const float v = R.dls[o / h2] * R.h2s[o % h2];
      if (v != 0.f) atomicAdd(d.lin_gw[2] + o, v);
    }
    for (int o = tid; o < h1 + h2 + K; o += RO_THREADS) {               // bias gradients
      if (o < h1) { if (R.g1s[o] != 0.f) atomicAdd(d.lin_gb[0] + o, R.g1s[o]); }
      else if (o < h1 + h2) { if (R.g2s[o - h1] != 0.f) atomicAdd(d.lin_gb[1] + o - h1, R.g2s[o - h1]); }
      else if (R.dls[o - h1 - h2] != 0.f) atomicAdd(d.lin_gb[2] + o - h1 - h2, R.dls[o - h1 - h2]);
    }
  }
  RO_TR(7);
  sg_cluster_sync();                                                    // gall complete in every CTA
  RO_TR(8);
  // ---- 3: dZ and dW1 of my slice: warp = column i of the slice, lanes over the h1 units (rows of the unpadded W1^T slice are read
  // contiguously: conflict free), the RO_BT per-sample sums reduced with shuffles ----
  const bool wgrad = d.mlp_fused && d.lin_gw[0];
  for (int i = warp; i < Sr; i += RO_THREADS / 32) {
    float acc[RO_BT];
#pragma unroll
    for (int u = 0; u < RO_BT; ++u) acc[u] = 0.f;
    for (int j = lane; j < h1; j += 32) {
      const float w = R.w1t[i * ld1 + j];
      const float4* g4 = reinterpret_cast<const float4*>(R.gall + (size_t)j * RO_BT);
      float dw = 0.f;
#pragma unroll
      for (int q = 0; q < RO_BT / 4; ++q) {
        const float4 g = g4[q];
        acc[4 * q + 0] = fmaf(g.x, w, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(g.y, w, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(g.z, w, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(g.w, w, acc[4 * q + 3]);
        if (wgrad) {
          const float4 z = make_float4(R.zs[(size_t)(4 * q + 0) * S + i], R.zs[(size_t)(4 * q + 1) * S + i], R.zs[(size_t)(4 * q + 2) * S + i],
                                       R.zs[(size_t)(4 * q + 3) * S + i]);
          dw = fmaf(g.x, z.x, fmaf(g.y, z.y, fmaf(g.z, z.z, fmaf(g.w, z.w, dw))));
        }
      }
      if (wgrad && dw != 0.f) atomicAdd(d.lin_gw[0] + (size_t)j * d.hid + i0 + i, dw);
    }
#pragma unroll
    for (int u = 0; u < RO_BT; ++u) acc[u] = warp_sum(acc[u]);
    if (lane < bn) {
      float v = acc[0];
#pragma unroll
      for (int u = 1; u < RO_BT; ++u) v = lane == u ? acc[u] : v;
      d.dZ[(size_t)(b0 + lane) * d.hid + i0 + i] = v;
    }
  }
  RO_TR(9);
}

template <int DPL, int RT>
__global__ void __launch_bounds__(RT) row_bwd_kernel(Desc d, int phases) {
  constexpr int ROW_THREADS = RT, SIDE_THREADS = RT / 2, SIDE_WARPS = RT / 64;
  sg_pdl_sync();
  extern __shared__ float sm[];
  const int D = d.D;
  float* dx0 = sm;                  // [D]
  float* dpre = dx0 + D;            // [2][D]
  float* din = dpre + 2 * D;        // [2][2D]  [dh ; dagg]
  float* acc_bp = din + 4 * D;      // [L][4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int side = warp / SIDE_WARPS, sw = warp % SIDE_WARPS, st = tid % SIDE_THREADS;
  const int R = d.meta[0];
  const int blk = 2 * D * D + 2 * D + 1;
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const int b = d.row_b[r], g = d.row_g[r];
    const int sub = d.batch_idx[b];
    const float* dz = d.dZ + (size_t)b * d.hid;
    __syncthreads();
    for (int k = tid; k < D; k += ROW_THREADS) dx0[k] = dz[k];
    for (int k = tid; k < d.L * 4; k += ROW_THREADS) acc_bp[k] = 0.f;
    for (int k = st; k < D; k += SIDE_THREADS) din[side * 2 * D + k] = 0.f;          // dh carried between layers
    __syncthreads();
    if (d.use_n && (phases & 1)) {
      const int A = side ? d.A_nb : d.A_ni;
      for (int l = d.L - 1; l >= 0; --l) {
        const int zc = col_n(d, l, side);
        if (d.use_proj) {
          for (int k = st; k < D; k += SIDE_THREADS) {
            const float outv = d.Nh[((size_t)((l + 1) * 2 + side) * d.R_cap + r) * D + k];
            const float g_ = outv > 0.f ? dz[zc + k] + din[side * 2 * D + k] : 0.f;
            dpre[side * D + k] = g_;
            d.Ndpre[((size_t)(l * 2 + side) * d.R_cap + r) * D + k] = g_;
          }
          __syncthreads();
          const float* w = d.mpn_params[0] + (size_t)(2 * l + side) * blk;            // [D][2D]
          const float* gp = dpre + side * D;
          for (int kk = st; kk < 2 * D; kk += SIDE_THREADS) {
            float acc = 0.f;
#pragma unroll 16                            // 16 weight loads in flight per thread (ncu source page: at 4, 18 % of the kernel's stall samples sat on this FMA)
            for (int j = 0; j < D; ++j) acc = fmaf(gp[j], __ldg(w + (size_t)j * 2 * D + kk), acc);
            din[side * 2 * D + kk] = acc;
          }
        } else {
          __syncthreads();
          for (int k = st; k < D; k += SIDE_THREADS) {
            din[side * 2 * D + D + k] = dz[zc + k] + din[side * 2 * D + k];
          }
          __syncthreads();
          for (int k = st; k < D; k += SIDE_THREADS) din[side * 2 * D + k] = 0.f;      // update() ignores x without projection
        }
        __syncthreads();
        if (d.dE) {
          const int* ids = d.n_ids[side] + ((size_t)l * d.n_cc + g) * A;
          const float* sims = d.n_sim[side] + ((size_t)l * d.n_cc + g) * A;
          const float* dagg = din + side * 2 * D + D;
          for (int a = sw; a < A; a += SIDE_WARPS) {
            const int id = ids[a];
            const float s = sims[a];
            if (id != 0 && s != 0.f)
#pragma unroll
              for (int q = 0; q < DPL; ++q) {
                const int k = lane + 32 * q;
                if (k < D) atomicAdd(d.dE + (size_t)id * D + k, s * dagg[k]);
              }
          }
        }
      }
      __syncthreads();
      // gradient of the layer-0 input of both sides
      if (d.trainable_cc) {
        for (int k = st; k < D; k += SIDE_THREADS)
          if (d.cc_tab_grad[side]) atomicAdd(d.cc_tab_grad[side] + ((size_t)sub * d.C_pad + (g - d.sub_ccptr[sub])) * D + k, din[side * 2 * D + k]);
      } else {
        for (int k = tid; k < D; k += ROW_THREADS) dx0[k] += din[k] + din[2 * D + k];
      }
      __syncthreads();
    }
    // ---- pooling backward ----
    if (d.dE && (phases & 1)) {
      const int nb = d.cc_nodeptr[g], ne = d.cc_nodeptr[g + 1];
      if (!d.pool_max) {
        for (int e = tid; e < (ne - nb) * D; e += ROW_THREADS) {
          const int i = nb + e / D, k = e % D;
          atomicAdd(d.dE + (size_t)d.cc_nodes[i] * D + k, dx0[k]);
        }
      } else {
        for (int k = tid; k < D; k += ROW_THREADS) {
          const float mx = d.X0[(size_t)r * D + k];
          for (int i = nb; i < ne; ++i) {
            const int id = d.cc_nodes[i];
            if (d.E[(size_t)id * D + k] == mx) { atomicAdd(d.dE + (size_t)id * D + k, dx0[k]); break; }
          }
        }
      }
    }
    // ---- property-aware outputs backward: d q (global atomics) and d b_p (CTA reduction) ----
    if (phases & SUBGNN_PHASE_PS) {
      const int wp_ = (d.use_p ? d.A_pi + d.A_pb : 0), ws_ = (d.use_s ? 2 * d.A_s : 0);
      const int per_layer = wp_ + ws_;
      // every warp runs the same number of iterations (out-of-range lanes carry key -1): the bias-gradient terms of a warp are
      // reduced with shuffles per (layer, slot) key before ONE shared-memory atomic — the keys of consecutive entries coincide,
      // and 32-way same-address shared atomics were the bulk of this phase
      const int n_ent = d.L * per_layer;
      for (int e0 = 0; e0 < n_ent; e0 += ROW_THREADS) {
        const int e = e0 + tid;
        int key = -1;
        float g_ = 0.f;
        if (e < n_ent && (phases & ((e % per_layer) < wp_ ? SUBGNN_PHASE_P : SUBGNN_PHASE_S))) {
        const int l = e / per_layer;
        int o = e % per_layer;
        float s, q, bp;
        float* dq;
        int slot;
        if (o < wp_) {
          const int sd = o >= d.A_pi ? 1 : 0, a = sd ? o - d.A_pi : o, A = sd ? d.A_pb : d.A_pi;
          s = d.p_sim[sd][((size_t)l * d.n_cc + g) * A + a];
          q = sd ? d.q_pb[(size_t)l * A + a] : d.q_pi[((size_t)l * d.B + b) * A + a];
          bp = d.mpn_params[1][(size_t)(2 * l + sd) * blk + 2 * D * D + 2 * D];
          g_ = dz[col_p(d, l, sd) + a];
          dq = sd ? d.dq_pb + (size_t)l * A + a : d.dq_pi + ((size_t)l * d.B + b) * A + a;
          slot = sd;
        } else {
          o -= wp_;
          const int sd = o / d.A_s, a = o % d.A_s;
          s = d.s_sim[sd][((size_t)l * d.n_cc + g) * d.A_s + a];
          q = d.q_s[((size_t)l * 2 + sd) * d.A_s + a];
          bp = d.mpn_params[2][(size_t)(2 * l + sd) * blk + 2 * D * D + 2 * D];
          g_ = dz[col_s(d, l, sd) + a];
          dq = d.dq_s + ((size_t)l * 2 + sd) * d.A_s + a;
          slot = 2 + sd;
        }
        if (fmaf(s, q, bp) > 0.f && g_ != 0.f) {
          if (s != 0.f) atomicAdd(dq, s * g_);
          key = l * 4 + slot;
        }
        }
        int kmax = __reduce_max_sync(0xffffffffu, key);
        while (kmax >= 0) {
          const float c = warp_sum(key == kmax ? g_ : 0.f);
          if (lane == 0) atomicAdd(acc_bp + kmax, c);
          if (key == kmax) key = -1;
          kmax = __reduce_max_sync(0xffffffffu, key);
        }
      }
      __syncthreads();
      for (int e = tid; e < d.L * 4; e += ROW_THREADS) {
        const int l = e / 4, slot = e % 4;
        const int ch = slot < 2 ? 1 : 2, sd = slot & 1;
        const bool on = ch == 1 ? d.use_p : d.use_s;
        const float v = acc_bp[e];
        if (on && d.mpn_grads[ch] && v != 0.f) atomicAdd(d.mpn_grads[ch] + (size_t)(2 * l + sd) * blk + 2 * D * D + 2 * D, v);
      }
    }
  }
}

// weight gradients of all N-channel projections in ONE launch: grid (2D/64, D/64, L*2*splits);
// dW[z] (D x 2D) += dpre[z]^T [Nh[z] | Nagg[z]] over the valid rows; db[z] += column sums of dpre[z]
#include "gemm_tile.cuh"
__global__ void __launch_bounds__(256) n_wgrad_kernel(Desc d, int splits, int m_chunk) {
  sg_pdl_sync();
  const int D = d.D;
  const int z = blockIdx.z / splits, sp = blockIdx.z % splits;
  const int R = d.meta[0];
  const int m_beg = sp * m_chunk, m_end = min(R, m_beg + m_chunk);
  if (m_beg >= m_end) return;
  const int blk = 2 * D * D + 2 * D + 1;
  const float* dpre = d.Ndpre + (size_t)z * d.R_cap * D;
  const float* hin = d.Nh + (size_t)z * d.R_cap * D;            // input of layer l == Nh[l] (z = 2 l + side)
  const float* agg = d.Nagg + (size_t)z * d.R_cap * D;
  float* gw = d.mpn_grads[0] + (size_t)z * blk;
  gemm_tile<false, false>(
      D, 2 * D, m_beg, m_end,
      [&](int n, int m) { return dpre[(size_t)m * D + n]; },
      [&](int m, int k) { return k < D ? hin[(size_t)m * D + k] : agg[(size_t)m * D + (k - D)]; },
      [&](int n, int k, float v) { atomicAdd(gw + (size_t)n * 2 * D + k, v); });
  if (blockIdx.x == 0) {                                         // bias gradient for this tile's rows
    const int n = blockIdx.y * 64 + (threadIdx.x % 64), part = threadIdx.x / 64;
    if (n < D) {
      float s = 0.f;
      for (int m = m_beg + part; m < m_end; m += 4) s += dpre[(size_t)m * D + n];
      if (s != 0.f) atomicAdd(gw + 2 * D * D + n, s);
    }
  }
}

// weight / bias gradients of the three MLP layers in ONE launch [r2]: dW_i[n][k] += sum_b dY_i[b][n] X_i[b][k], db_i[n] += sum_b dY_i[b][n]
// with (dY, X) = (dH1, Z), (dH2, H1), (d logits, H2) — SubGNN.py:303-310 backward.  One CTA per 64 x 64 tile of one dW (a single
// writer per element: no atomics); thread (tn, tk) owns a 4 x 4 block, the batch is walked in chunks of 32 samples staged in shared
// memory.  Three generic split-reduction launches took 28 us for these ~4 MFLOP at the reference batch sizes (7 - 11 us of fixed
// latency each); on a one-layer walk encoder they end the backward pass.
#define MW_BT 32
__global__ void __launch_bounds__(256) mlp_wgrad_kernel(Desc d, int tiles0, int tiles1) {
  sg_pdl_sync();
  __shared__ float ys[MW_BT][64 + 4];
  __shared__ float xs[MW_BT][64 + 4];
  int t = blockIdx.x, li = 0;
  if (t >= tiles0) { t -= tiles0; li = 1; }
  if (li == 1 && t >= tiles1) { t -= tiles1; li = 2; }
  const float* dy = li == 0 ? d.dH1 : li == 1 ? d.dH2 : d.dlogits;
  const float* x = li == 0 ? d.Z : li == 1 ? d.H1 : d.H2;
  const int N = li == 0 ? d.h1 : li == 1 ? d.h2 : d.n_classes;
  const int K = li == 0 ? d.hid : li == 1 ? d.h1 : d.h2;
  float* gw = d.lin_gw[li];
  float* gb = d.lin_gb[li];
  const int tiles_k = (K + 63) / 64;
  const int n0 = (t / tiles_k) * 64, k0 = (t % tiles_k) * 64;
  const int tn = threadIdx.x / 16, tk = threadIdx.x % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;                                       // threads 0 .. 63 of the k0 == 0 tiles: column sums of dY
  for (int b0 = 0; b0 < d.B; b0 += MW_BT) {
    const int bn = min(MW_BT, d.B - b0);
    __syncthreads();
    for (int e = threadIdx.x; e < MW_BT * 64; e += 256) {
      const int b = e / 64, c = e % 64;
      ys[b][c] = (b < bn && n0 + c < N) ? dy[(size_t)(b0 + b) * N + n0 + c] : 0.f;
      xs[b][c] = (b < bn && k0 + c < K) ? x[(size_t)(b0 + b) * K + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int b = 0; b < MW_BT; ++b) {
      const float4 yv = *reinterpret_cast<const float4*>(&ys[b][4 * tn]);
      const float4 xv = *reinterpret_cast<const float4*>(&xs[b][4 * tk]);
      const float ya[4] = {yv.x, yv.y, yv.z, yv.w}, xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ya[i], xa[j], acc[i][j]);
    }
    if (k0 == 0 && threadIdx.x < 64)
      for (int b = 0; b < bn; ++b) bsum += ys[b][threadIdx.x];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + 4 * tn + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + 4 * tk + j;
      if (k < K) gw[(size_t)n * K + k] += acc[i][j];
    }
  }
  if (k0 == 0 && threadIdx.x < 64 && n0 + threadIdx.x < N && gb) gb[n0 + threadIdx.x] += bsum;
}

// ------------------------------------------------------------------------------------------------
static size_t bwd_smem(const Desc& d) { return (size_t)(d.D + 2 * d.D + 4 * d.D + d.L * 4 + 16) * sizeof(float); }
static int check_desc(const Desc* d) {
  if (!d) { subgnn_set_error("null descriptor"); return SUBGNN_ERR_ARG; }
  if (d->D < 1 || d->D > 32 * MAXDPL) { subgnn_set_error("node_embed_size must be in [1, %d]", 32 * MAXDPL); return SUBGNN_ERR_ARG; }
  if (d->B < 1 || d->L < 1) { subgnn_set_error("bad batch / layer count"); return SUBGNN_ERR_ARG; }
  if ((size_t)(d->hid + d->h1 + d->h2 + d->n_classes + 8) * sizeof(float) > 200 * 1024) {
    subgnn_set_error("hidden dimension too large for shared memory");
    return SUBGNN_ERR_ARG;
  }
  return SUBGNN_OK;
}

static int mlp_wgrad_split(const subgnn_model_desc* d, void* stream);
static int row_grid(const Desc* d);
// 16 warps per row while the row capacity (B x most components of a subgraph: an upper bound of the rows) stays within the grid cap, else 8 (see ROW_WIDE)
static int row_width(const Desc* d) {
  if (const char* e = getenv("SUBGNN_ROW_THREADS")) { const int v = atoi(e); if (v == ROW_WIDE || v == ROW_NARROW) return v; }
  return d->R_cap <= 8 * subgnn_sm_count() ? ROW_WIDE : ROW_NARROW;
}

template <int DPL, int RT>
static void launch_row_fwd(const Desc* d, int phases, cudaStream_t st) {
  const size_t smem = row_smem_bytes(d->D, d->L, RT);
  if (smem > 48 * 1024)                          // deep / wide configurations: opt in to the large dynamic shared-memory window
    cudaFuncSetAttribute(row_fwd_kernel<DPL, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  subgnn_note_variant("row_fwd_kernel<%d>[%d threads]", DPL, RT);
  sg_launch_pdl<SG_PDL_ROW>(row_fwd_kernel<DPL, RT>, dim3(row_grid(d)), dim3(RT), smem, st, *d, phases);
}
template <int DPL, int RT>
static void launch_row_bwd(const Desc* d, int phases, size_t smem, cudaStream_t st) {
  subgnn_note_variant("row_bwd_kernel<%d>[%d threads]", DPL, RT);
  sg_launch_pdl<SG_PDL_ROW>(row_bwd_kernel<DPL, RT>, dim3(row_grid(d)), dim3(RT), smem, st, *d, phases);
}
#define DISPATCH_ROW(D, RT, FN, ...)                                                                 \
  do {                                                                                               \
    const int dpl_ = ((D) + 31) / 32;                                                                \
    if ((RT) == ROW_WIDE) {                                                                          \
      if (dpl_ <= 1) FN<1, ROW_WIDE>(__VA_ARGS__); else if (dpl_ <= 2) FN<2, ROW_WIDE>(__VA_ARGS__); \
      else if (dpl_ <= 4) FN<4, ROW_WIDE>(__VA_ARGS__); else FN<8, ROW_WIDE>(__VA_ARGS__);           \
    } else {                                                                                         \
      if (dpl_ <= 1) FN<1, ROW_NARROW>(__VA_ARGS__); else if (dpl_ <= 2) FN<2, ROW_NARROW>(__VA_ARGS__); \
      else if (dpl_ <= 4) FN<4, ROW_NARROW>(__VA_ARGS__); else FN<8, ROW_NARROW>(__VA_ARGS__);       \
    }                                                                                                \
  } while (0)

static int row_grid(const Desc* d) {
  const int cap = subgnn_sm_count() * 8;
  return d->R_cap < cap ? d->R_cap : cap;
}

extern "C" {

int subgnn_model_desc_size(void) { return (int)sizeof(subgnn_model_desc); }

int subgnn_model_prep_batch(const subgnn_model_desc* d, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  sg_launch_pdl(prep_batch_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, *d);
  return subgnn_check_launch("prep_batch_kernel");
}

int subgnn_model_prep_weights(const subgnn_model_desc* d, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  dim3 grid(32, (d->use_n ? 2 * d->L : 0) + 3);
  sg_launch_pdl(transpose_weights_kernel, grid, dim3(256), 0, (cudaStream_t)stream, *d);
  return subgnn_check_launch("transpose_weights_kernel");
}

int subgnn_model_q_fwd_part(const subgnn_model_desc* d, int which, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  const long long total = (long long)d->L * ((d->use_p ? d->B * d->A_pi + d->A_pb : 0) + (d->use_s ? 2 * d->A_s : 0));
  if (total == 0) return SUBGNN_OK;
  if (!d->use_p) which &= ~SUBGNN_Q_POS;
  if (!d->use_s) which &= ~SUBGNN_Q_STRUC;
  if (!which) return SUBGNN_OK;
  if (which == SUBGNN_Q_STRUC)            // structure anchors: between the LSTM head and the readout, on the critical chain
    sg_launch_pdl<SG_PDL_CHAIN>(q_fwd_kernel, dim3(sg_grid_for(total, 8, 8)), dim3(256), 0, (cudaStream_t)stream, *d, which);
  else
    sg_launch_pdl(q_fwd_kernel, dim3(sg_grid_for(total, 8, 8)), dim3(256), 0, (cudaStream_t)stream, *d, which);
  return subgnn_check_launch("q_fwd_kernel");
}

int subgnn_model_rows_fwd(const subgnn_model_desc* d, int phases, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!d->use_p) phases &= ~SUBGNN_PHASE_P;
  if (!d->use_s) phases &= ~SUBGNN_PHASE_S;
  if (!phases) return SUBGNN_OK;
  const int rt = row_width(d);
  SG_REQUIRE(row_smem_bytes(d->D, d->L, rt) <= 200 * 1024, "n_layers x node_embed_size too large for the row kernel's shared-memory aggregates");
  DISPATCH_ROW(d->D, rt, launch_row_fwd, d, phases, (cudaStream_t)stream);
  return subgnn_check_launch("row_fwd_kernel");
}

// requires d->H1 zeroed on entry (it is the split-K accumulation target of the first layer)
int subgnn_model_mlp_stage(const subgnn_model_desc* d, int stages, int cols, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  SG_REQUIRE(cols >= SUBGNN_COLS_ALL && cols <= SUBGNN_COLS_S, "unknown column set");
  cudaStream_t st = (cudaStream_t)stream;
  const int slices = sg_div_up(d->hid, MLP_SLICE);
  const int bt = d->B < MLP_BT ? d->B : MLP_BT, b_tiles = sg_div_up(d->B, MLP_BT);
  const size_t s1 = (size_t)(bt * MLP_SLICE + MLP_SLICE * d->h1) * sizeof(float);
  const size_t s2 = (size_t)(2 * d->h1 * d->h2 + d->h1 + 2 * d->h2 + 2 * d->n_classes + (d->h1 + d->h2 + d->n_classes + d->n_classes * d->h2) + 8) * sizeof(float);
  const size_t s3 = (size_t)(bt * d->h1 + d->h1 * MLP_SLICE) * sizeof(float);
  if (s1 > 200 * 1024 || s2 > 200 * 1024 || s3 > 200 * 1024) { subgnn_set_error("MLP dimensions too large for shared memory"); return SUBGNN_ERR_ARG; }
  if (s1 > 48 * 1024) cudaFuncSetAttribute(mlp_lin1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1);
  if (s2 > 48 * 1024) cudaFuncSetAttribute(mlp_rest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2);
  if (s3 > 48 * 1024) cudaFuncSetAttribute(mlp_dz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3);
  if (stages & SUBGNN_MLP_LIN1) {
    // the LSTM-independent columns run beside the LSTM chain: launched normally (an early-scheduled grid would take SM slots from it)
    if (cols == SUBGNN_COLS_NOT_S) sg_launch_pdl(mlp_lin1_kernel, dim3(slices, b_tiles), dim3(256), s1, st, *d, cols);
    else sg_launch_pdl<SG_PDL_CHAIN>(mlp_lin1_kernel, dim3(slices, b_tiles), dim3(256), s1, st, *d, cols);
    rc = subgnn_check_launch("mlp_lin1_kernel");
    if (rc) return rc;
  }
  if (stages & SUBGNN_MLP_REST) {
    sg_launch_pdl<SG_PDL_CHAIN>(mlp_rest_kernel, dim3(d->B), dim3(256), s2, st, *d);
    rc = subgnn_check_launch("mlp_rest_kernel");
    if (rc) return rc;
  }
  if ((stages & SUBGNN_MLP_DZ) && d->training && d->dZ) {
    if (cols == SUBGNN_COLS_NOT_S) sg_launch_pdl(mlp_dz_kernel, dim3(slices, b_tiles), dim3(256), s3, st, *d, cols);
    else sg_launch_pdl<SG_PDL_CHAIN>(mlp_dz_kernel, dim3(slices, b_tiles), dim3(256), s3, st, *d, cols);
    rc = subgnn_check_launch("mlp_dz_kernel");
  }
  return rc;
}

int subgnn_model_mlp_fwd(const subgnn_model_desc* d, void* stream) {
  return subgnn_model_mlp_stage(d, SUBGNN_MLP_LIN1 | SUBGNN_MLP_REST | SUBGNN_MLP_DZ, SUBGNN_COLS_ALL, stream);
}

int subgnn_model_readout_supported(const subgnn_model_desc* d) {
  if (!d || d->h1 < 1 || d->h2 < 1 || d->h1 > 256 || d->h2 > 256 || d->n_classes < 1 || d->n_classes > 64 || d->hid < RO_CL) return 0;
  if (ro_slice(d->hid) > 4 * RO_THREADS) return 0;
  return ro_smem_floats(d->hid, d->h1, d->h2, d->n_classes) * sizeof(float) <= 200 * 1024 ? 1 : 0;
}

int subgnn_model_readout(const subgnn_model_desc* d, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  SG_REQUIRE(subgnn_model_readout_supported(d), "shape outside the cluster readout kernel (use subgnn_model_mlp_fwd)");
  const size_t smem = ro_smem_floats(d->hid, d->h1, d->h2, d->n_classes) * sizeof(float);
  cudaFuncSetAttribute(readout_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(RO_CL * sg_div_up(d->B, RO_BT), 1, 1);      // B CTAs (rounded up to whole clusters)
  cfg.blockDim = dim3(RO_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = RO_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = subgnn_pdl_enabled(SG_PDL_CHAIN);
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  subgnn_note_variant("readout_cluster_kernel");
  cudaLaunchKernelEx(&cfg, readout_cluster_kernel, *d);
  return subgnn_check_launch("readout_cluster_kernel");
}

int subgnn_model_mlp_bwd(const subgnn_model_desc* d, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  const size_t smem = (size_t)(d->hid + d->h1 + d->h2 + d->n_classes + 8) * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mlp_bwd_kernel<<<d->B, MLP_THREADS, smem, (cudaStream_t)stream>>>(*d);
  return subgnn_check_launch("mlp_bwd_kernel");
}

int subgnn_model_rows_bwd(const subgnn_model_desc* d, int phases, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!d->use_p) phases &= ~SUBGNN_PHASE_P;
  if (!d->use_s) phases &= ~SUBGNN_PHASE_S;
  if (!phases) return SUBGNN_OK;
  DISPATCH_ROW(d->D, row_width(d), launch_row_bwd, d, phases, bwd_smem(*d), (cudaStream_t)stream);
  return subgnn_check_launch("row_bwd_kernel");
}

int subgnn_model_q_bwd_part(const subgnn_model_desc* d, int which, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!d->use_p) which &= ~SUBGNN_Q_POS;
  if (!d->use_s) which &= ~SUBGNN_Q_STRUC;
  const int groups = (d->use_p ? 2 : 0) + (d->use_s ? 2 : 0);
  if (groups == 0 || !which) return SUBGNN_OK;
  // grid.x sized for the largest group of the launch (the B * A_pi position-internal entries): ~2-3 entries per warp keeps the
  // dependent id -> row -> atomics chains short; CTAs of the small / unselected groups retire immediately
  int max_cnt = 0;
  if (which & SUBGNN_Q_POS) max_cnt = d->B * d->A_pi > d->A_pb ? d->B * d->A_pi : d->A_pb;
  if ((which & SUBGNN_Q_STRUC) && d->A_s > max_cnt) max_cnt = d->A_s;
  int gx = sg_div_up(max_cnt, 24);
  gx = gx < 8 ? 8 : (gx > 96 ? 96 : gx);
  dim3 grid(gx, d->L * groups);
  if (which & SUBGNN_Q_STRUC) sg_launch_pdl<SG_PDL_CHAIN>(q_bwd_kernel, grid, dim3(256), 0, (cudaStream_t)stream, *d, which);
  else sg_launch_pdl(q_bwd_kernel, grid, dim3(256), 0, (cudaStream_t)stream, *d, which);
  return subgnn_check_launch("q_bwd_kernel");
}

int subgnn_model_q_bwd(const subgnn_model_desc* d, void* stream) { return subgnn_model_q_bwd_part(d, SUBGNN_Q_POS | SUBGNN_Q_STRUC, stream); }

int subgnn_model_wgrad(const subgnn_model_desc* d, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  const int D = d->D;
  if (d->use_n && d->use_proj && d->mpn_grads[0]) {
    // rows per CTA.  [r2] measured (B200, same box, 100 steps): chunks of 512 rows leave the launch at 16 CTAs for 32 - 42 us; where it
    // hides behind a two-layer BPTT chain that is the better shape (PPI-BP 0.3052 ms/step against 0.3065 with 64-row chunks,
    // HPO-METAB 0.5854 / 0.5893), where it ends the backward pass (one-layer walk encoder) many CTAs win (EM-USER 0.5611 -> 0.5345,
    // density 0.147 -> 0.142): the caller says which (desc.wgrad_rows)
    int m_rows = 512;
    if (d->wgrad_rows >= 16) {
      m_rows = sg_div_up(sg_div_up(d->R_cap, 16), 16) * 16;                  // at most 16 chunks
      if (m_rows < d->wgrad_rows) m_rows = d->wgrad_rows;
    }
    if (const char* e = getenv("SUBGNN_NWGRAD_ROWS")) { const int v = atoi(e); if (v >= 16) m_rows = v; }
    int splits = sg_div_up(d->R_cap, m_rows);
    const int m_chunk = sg_div_up(sg_div_up(d->R_cap, splits), 16) * 16;
    splits = sg_div_up(d->R_cap, m_chunk);
    dim3 grid(sg_div_up(2 * D, 64), sg_div_up(D, 64), d->L * 2 * splits);
    sg_launch_pdl(n_wgrad_kernel, grid, dim3(256), 0, (cudaStream_t)stream, *d, splits, m_chunk);
    rc = subgnn_check_launch("n_wgrad_kernel");
    if (rc) return rc;
  }
  // at the END of the main stream (behind a two-layer BPTT chain) the three split-reduction launches stay: measured PPI-BP 0.3071
  // ms/step against 0.3117 with the one-launch kernel below (which wins where these gradients end the step: density 0.1427 -> 0.130)
  if (d->lin_gw[0] && !d->mlp_fused) return mlp_wgrad_split(d, stream);
  return SUBGNN_OK;
}

// weight / bias gradients of the readout MLP: they need dH1, dH2, d logits and the forward activations only, i.e. everything the
// readout section has produced — the fused step launches them right behind it, on a branch beside the backward chain, instead of
// at the end of the main stream (where they queued behind the last BPTT recurrence and shared the SMs with the GEMM tail)
int subgnn_model_mlp_wgrad(const subgnn_model_desc* d, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!d->lin_gw[0]) return SUBGNN_OK;
  static const bool one_launch = !(getenv("SUBGNN_MLP_WGRAD_SPLIT") && atoi(getenv("SUBGNN_MLP_WGRAD_SPLIT")) == 1);
  if (one_launch && d->lin_gw[1] && d->lin_gw[2]) {
    const int t0 = sg_div_up(d->h1, 64) * sg_div_up(d->hid, 64), t1 = sg_div_up(d->h2, 64) * sg_div_up(d->h1, 64);
    const int t2 = sg_div_up(d->n_classes, 64) * sg_div_up(d->h2, 64);
    sg_launch_pdl(mlp_wgrad_kernel, dim3(t0 + t1 + t2), dim3(256), 0, (cudaStream_t)stream, *d, t0, t1);
    return subgnn_check_launch("mlp_wgrad_kernel");
  }
  return mlp_wgrad_split(d, stream);
}

}  // extern "C"

// dW1 = dH1^T Z, dW2 = dH2^T H1, dW3 = dlogits^T H2 (reduction over the B samples) as three split-reduction launches
static int mlp_wgrad_split(const subgnn_model_desc* d, void* stream) {
  int rc;
  rc = subgnn_linear_bwd_weight(d->dH1, d->h1, d->Z, d->hid, nullptr, d->lin_gw[0], d->hid, d->lin_gb[0], d->B, d->h1, d->hid, nullptr, stream);
  if (rc) return rc;
  rc = subgnn_linear_bwd_weight(d->dH2, d->h2, d->H1, d->h1, nullptr, d->lin_gw[1], d->h1, d->lin_gb[1], d->B, d->h2, d->h1, nullptr, stream);
  if (rc) return rc;
  return subgnn_linear_bwd_weight(d->dlogits, d->n_classes, d->H2, d->h2, nullptr, d->lin_gw[2], d->h2, d->lin_gb[2], d->B, d->n_classes, d->h2,
                                  nullptr, stream);
}
