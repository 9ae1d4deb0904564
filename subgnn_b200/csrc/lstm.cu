// Walk-encoder LSTM (bidirectional, 1-2 layers) — recurrent forward / BPTT kernels and the aggregation head.
//
// Replaces the cuDNN/ATen RNN behind (reference):
//   SubGNN.py:60-88                     class LSTM (nn.LSTM bidirectional + 'last' | 'sum' aggregator + Linear(2h -> D))
//   anchor_patch_samplers.py:413-433    aggregate_structure_anchor_patch (embed walks -> LSTM -> sum over the W walks)
//
// Data layout (all fp32, row-major):
//   G   [n_seq][T][2][4H]   gate pre-activations x_t W_ih^T + b_ih + b_hh of both directions, written by the input
//                           projection GEMM (gemm.cu), overwritten in place by the recurrent kernel with the gate
//                           activations (i, f, g, o) and, in the backward pass, by d(pre-activation).
//   OUT [n_seq*T + 1][2H]   h_t of both directions ([:, :H] forward, [:, H:] reverse) == the next layer's input
//                           == nn.LSTM's output; the extra last row stays zero (h_{-1} for the dW_hh gather).
//   CS  [n_seq][T][2][H]    cell states.
// One CTA owns S_TILE sequences of one direction for the whole recurrence: h and c live in shared memory,
// thread j owns gate column j (coalesced reads of the transposed W_hh, L1/L2 resident), the four gates of a
// hidden unit are exchanged through shared memory.  PAD steps are fed as zero vectors (no packing), exactly like
// the reference.
#include "common.cuh"
#include "../../include/subgnn_b200.h"
#include "lstm_reg.cuh"

#define S_TILE 8

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// One-instruction staging of the recurrent weights: a 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) global -> shared that
// completes on an mbarrier, issued by one thread while the CTA sets up its state.  The round-1 profile showed 25% of the
// recurrence kernels' stall samples on the per-thread load -> st.shared staging loop this replaces.
__device__ __forceinline__ uint32_t lstm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s_issue(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(lstm_smem_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(lstm_smem_u32(bar)), "r"(bytes) : "memory");
  // chunks of at most 32 KB keep every copy well inside the instruction's size field
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(lstm_smem_u32(reinterpret_cast<const char*>(dst) + off)), "l"(reinterpret_cast<const char*>(src) + off), "r"(n),
                    "r"(lstm_smem_u32(bar)) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s_wait(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLSTM_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t"
      "@P1 bra LSTM_DONE;\n\tbra LSTM_WAIT;\n\tLSTM_DONE:\n\t}\n"
      :: "r"(lstm_smem_u32(bar)) : "memory");
}

// grid (ceil(n_seq / S_TILE), 2 directions); block = 4H rounded up to 32.
// W_SMEM: the transposed recurrent weights of this direction (H x 4H fp32) are staged once in shared memory
// (H <= 64: 64 KB); otherwise they are streamed from L1/L2 every step.  The gate pre-activations of step t+1 are
// prefetched into registers while step t's matvec runs.
template <bool W_SMEM, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
lstm_recur_fwd_kernel(float* __restrict__ G, const float* __restrict__ WhhT /*[2][H][4H]*/, float* __restrict__ OUT,
                      float* __restrict__ CS, int n_seq, int T, int H, int steps_fwd, int steps_rev) {
  extern __shared__ float sm[];
  float* hbuf = sm;                       // [S_TILE][H]
  float* cbuf = hbuf + S_TILE * H;        // [S_TILE][H]
  float* gates = cbuf + S_TILE * H;       // [S_TILE][4H]
  float* wsm = gates + S_TILE * 4 * H;    // [H][4H] (W_SMEM only)
  const int dir = blockIdx.y;
  const int seq0 = blockIdx.x * S_TILE;
  const int ns = min(S_TILE, n_seq - seq0);
  const int j = threadIdx.x;
  const int H4 = 4 * H;
  const float* wt = WhhT + (size_t)dir * H * H4;
  const int n_steps = dir == 0 ? steps_fwd : steps_rev;
  __shared__ uint64_t wbar;
  if (W_SMEM) {
    if (threadIdx.x == 0) bulk_g2s_issue(wsm, wt, (uint32_t)(H * H4 * sizeof(float)), &wbar);
    wt = wsm;
  }
  for (int e = threadIdx.x; e < S_TILE * H; e += blockDim.x) { hbuf[e] = 0.f; cbuf[e] = 0.f; }
  float gnext[S_TILE];
  if (j < H4 && n_steps > 0) {
    const int t0 = dir == 0 ? 0 : T - 1;
#pragma unroll
    for (int s = 0; s < S_TILE; ++s) gnext[s] = s < ns ? G[(((size_t)(seq0 + s) * T + t0) * 2 + dir) * H4 + j] : 0.f;
  }
  __syncthreads();                       // also publishes the mbarrier initialisation to the waiting threads
  if (W_SMEM) bulk_g2s_wait(&wbar);
  for (int st = 0; st < n_steps; ++st) {
    const int t = dir == 0 ? st : T - 1 - st;
    if (j < H4) {
      float acc[S_TILE];
#pragma unroll
      for (int s = 0; s < S_TILE; ++s) acc[s] = gnext[s];
      if (st + 1 < n_steps) {
        const int tn = dir == 0 ? st + 1 : T - 2 - st;
#pragma unroll
        for (int s = 0; s < S_TILE; ++s) gnext[s] = s < ns ? G[(((size_t)(seq0 + s) * T + tn) * 2 + dir) * H4 + j] : 0.f;
      }
      if ((H & 3) == 0) {
        for (int k = 0; k < H; k += 4) {
          float w0, w1, w2, w3;
          if (W_SMEM) { w0 = wt[k * H4 + j]; w1 = wt[(k + 1) * H4 + j]; w2 = wt[(k + 2) * H4 + j]; w3 = wt[(k + 3) * H4 + j]; }
          else { w0 = __ldg(wt + (size_t)k * H4 + j); w1 = __ldg(wt + (size_t)(k + 1) * H4 + j);
                 w2 = __ldg(wt + (size_t)(k + 2) * H4 + j); w3 = __ldg(wt + (size_t)(k + 3) * H4 + j); }
#pragma unroll
          for (int s = 0; s < S_TILE; ++s) {
            const float4 h4 = *reinterpret_cast<const float4*>(hbuf + s * H + k);
            acc[s] = fmaf(h4.x, w0, acc[s]); acc[s] = fmaf(h4.y, w1, acc[s]);
            acc[s] = fmaf(h4.z, w2, acc[s]); acc[s] = fmaf(h4.w, w3, acc[s]);
          }
        }
      } else {
        for (int k = 0; k < H; ++k) {
          const float w = W_SMEM ? wt[k * H4 + j] : __ldg(wt + (size_t)k * H4 + j);
#pragma unroll
          for (int s = 0; s < S_TILE; ++s) acc[s] = fmaf(hbuf[s * H + k], w, acc[s]);
        }
      }
      const bool is_tanh = (j / H) == 2;
#pragma unroll
      for (int s = 0; s < S_TILE; ++s) {
        const float a = is_tanh ? tanhf(acc[s]) : sigmoidf_(acc[s]);
        gates[s * H4 + j] = a;
        if (s < ns) G[(((size_t)(seq0 + s) * T + t) * 2 + dir) * H4 + j] = a;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ns * H; e += blockDim.x) {
      const int s = e / H, u = e % H;
      const float ig = gates[s * H4 + u], fg = gates[s * H4 + H + u], gg = gates[s * H4 + 2 * H + u], og = gates[s * H4 + 3 * H + u];
      const float c = fmaf(fg, cbuf[e], ig * gg);
      const float h = og * tanhf(c);
      cbuf[e] = c;
      hbuf[e] = h;
      CS[(((size_t)(seq0 + s) * T + t) * 2 + dir) * H + u] = c;
      OUT[((size_t)(seq0 + s) * T + t) * 2 * H + dir * H + u] = h;
    }
    __syncthreads();
  }
  // steps not taken (top-layer reverse direction with the 'last' aggregator): outputs are never read
}

// BPTT.  dOUT [n_seq*T][2H] gradient w.r.t. this layer's outputs; G holds gate activations on entry and
// d(pre-activation) on exit (zero for steps that were not taken).  Whh [2][4H][H] native layout.
// db_ih / db_hh (optional, [2][4H]): the bias gradients (column sums of d(pre-activation) over every row this CTA owns,
// identical for the two bias vectors) are accumulated in registers along the recurrence and added once per CTA.
template <bool W_SMEM, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
lstm_recur_bwd_kernel(float* __restrict__ G, const float* __restrict__ Whh, const float* __restrict__ OUT,
                      const float* __restrict__ CS, const float* __restrict__ dOUT, int n_seq, int T, int H, int steps_fwd,
                      int steps_rev, int zero_untaken, float* __restrict__ db_ih, float* __restrict__ db_hh) {
  extern __shared__ float sm[];
  const int H4 = 4 * H;
  float* dgate = sm;                        // [S_TILE][4H]
  float* dh_rec = dgate + S_TILE * H4;      // [S_TILE][H]
  float* dc_rec = dh_rec + S_TILE * H;      // [S_TILE][H]
  float* red = dc_rec + S_TILE * H;         // [4][S_TILE][H]
  float* wsm = red + 4 * S_TILE * H;        // [4H][H] (W_SMEM only)
  const int dir = blockIdx.y;
  const int seq0 = blockIdx.x * S_TILE;
  const int ns = min(S_TILE, n_seq - seq0);
  const float* w = Whh + (size_t)dir * H4 * H;
  __shared__ uint64_t wbar;
  if (W_SMEM) {
    if (threadIdx.x == 0) bulk_g2s_issue(wsm, w, (uint32_t)(H4 * H * sizeof(float)), &wbar);
    w = wsm;
  }
  const int n_steps = dir == 0 ? steps_fwd : steps_rev;
  for (int e = threadIdx.x; e < S_TILE * H; e += blockDim.x) { dh_rec[e] = 0.f; dc_rec[e] = 0.f; }
  for (int e = threadIdx.x; e < S_TILE * H4; e += blockDim.x) dgate[e] = 0.f;
  float bias_acc = 0.f;
  __syncthreads();
  if (W_SMEM) bulk_g2s_wait(&wbar);
  for (int st = n_steps - 1; st >= 0; --st) {
    const int t = dir == 0 ? st : T - 1 - st;
    const int t_prev = dir == 0 ? t - 1 : t + 1;
    for (int e = threadIdx.x; e < ns * H; e += blockDim.x) {
      const int s = e / H, u = e % H;
      const size_t gbase = (((size_t)(seq0 + s) * T + t) * 2 + dir) * H4;
      const float ig = G[gbase + u], fg = G[gbase + H + u], gg = G[gbase + 2 * H + u], og = G[gbase + 3 * H + u];
      const float c = CS[(((size_t)(seq0 + s) * T + t) * 2 + dir) * H + u];
      const float c_prev = st > 0 ? CS[(((size_t)(seq0 + s) * T + t_prev) * 2 + dir) * H + u] : 0.f;
      const float tc = tanhf(c);
      const float dh = dOUT[((size_t)(seq0 + s) * T + t) * 2 * H + dir * H + u] + dh_rec[e];
      const float dc = dc_rec[e] + dh * og * (1.f - tc * tc);
      const float dai = dc * gg * ig * (1.f - ig);
      const float daf = dc * c_prev * fg * (1.f - fg);
      const float dag = dc * ig * (1.f - gg * gg);
      const float dao = dh * tc * og * (1.f - og);
      dc_rec[e] = dc * fg;
      dgate[s * H4 + u] = dai; dgate[s * H4 + H + u] = daf; dgate[s * H4 + 2 * H + u] = dag; dgate[s * H4 + 3 * H + u] = dao;
      G[gbase + u] = dai; G[gbase + H + u] = daf; G[gbase + 2 * H + u] = dag; G[gbase + 3 * H + u] = dao;
    }
    __syncthreads();
    // dh_rec[s][k] = sum_j dgate[s][j] * Whh[j][k]; thread -> (k, quarter of j)
    if (threadIdx.x < H4) {
      {
        float bs = 0.f;                                         // rows s >= ns of dgate stay zero
#pragma unroll
        for (int s = 0; s < S_TILE; ++s) bs += dgate[s * H4 + threadIdx.x];
        bias_acc += bs;
      }
      const int k = threadIdx.x % H, part = threadIdx.x / H;
      float acc[S_TILE];
#pragma unroll
      for (int s = 0; s < S_TILE; ++s) acc[s] = 0.f;
      if ((H & 3) == 0) {
        for (int jj = part * H; jj < (part + 1) * H; jj += 4) {
          const float w0 = w[(size_t)jj * H + k], w1 = w[(size_t)(jj + 1) * H + k];
          const float w2 = w[(size_t)(jj + 2) * H + k], w3 = w[(size_t)(jj + 3) * H + k];
#pragma unroll
          for (int s = 0; s < S_TILE; ++s) {
            const float4 g4 = *reinterpret_cast<const float4*>(dgate + s * H4 + jj);
            acc[s] = fmaf(g4.x, w0, acc[s]); acc[s] = fmaf(g4.y, w1, acc[s]);
            acc[s] = fmaf(g4.z, w2, acc[s]); acc[s] = fmaf(g4.w, w3, acc[s]);
          }
        }
      } else {
        for (int jj = part * H; jj < (part + 1) * H; ++jj) {
          const float wv = w[(size_t)jj * H + k];
#pragma unroll
          for (int s = 0; s < S_TILE; ++s) acc[s] = fmaf(dgate[s * H4 + jj], wv, acc[s]);
        }
      }
#pragma unroll
      for (int s = 0; s < S_TILE; ++s) red[(part * S_TILE + s) * H + k] = acc[s];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < S_TILE * H; e += blockDim.x)
      dh_rec[e] = red[e] + red[S_TILE * H + e] + red[2 * S_TILE * H + e] + red[3 * S_TILE * H + e];
    __syncthreads();
  }
  if (threadIdx.x < H4 && bias_acc != 0.f) {
    if (db_ih) atomicAdd(db_ih + (size_t)dir * H4 + threadIdx.x, bias_acc);
    if (db_hh) atomicAdd(db_hh + (size_t)dir * H4 + threadIdx.x, bias_acc);
  }
  // steps that were never taken carry no gradient: zero their slots (they still hold pre-activations) unless the
  // caller never reads them (zero_untaken == 0)
  for (int st = n_steps; zero_untaken && st < T; ++st) {
    const int t = dir == 0 ? st : T - 1 - st;
    for (int e = threadIdx.x; e < ns * H4; e += blockDim.x) {
      const int s = e / H4, jg = e % H4;
      G[(((size_t)(seq0 + s) * T + t) * 2 + dir) * H4 + jg] = 0.f;
    }
  }
}

// Walk-group head (anchor_patch_samplers.py:429-433: patch embedding = sum over its walks of Linear(agg(lstm_out))).  The head
// is linear, so the walks are summed first and the Linear runs once per patch.  One CTA per patch g:
//   AGG[g][c] = sum_{w<group} agg(OUT[g*group + w])[c]          (agg = row T-1 for 'last', SubGNN.py:83; sum over t for 'sum', :85)
//   EMB[g][d] = sum_c AGG[g][c] W[d][c] + group * bias[d]        (warp per output, lanes along c: coalesced weight rows)
// group = 1 is the plain per-sequence head of LSTM.forward (SubGNN.py:76-88).
__global__ void __launch_bounds__(512) lstm_head_fwd_kernel(const float* __restrict__ OUT, float* __restrict__ AGG, float* __restrict__ EMB,
                                                            const float* __restrict__ W, const float* __restrict__ bias, int n_groups,
                                                            int group, int T, int H2, int D, int sum_mode) {
  sg_pdl_sync();
  extern __shared__ float sm_head[];
  const int g = blockIdx.x;
  const int rows = sum_mode ? group * T : group;          // rows of OUT that are summed: every (w, t) or (w, T-1)
  for (int c = threadIdx.x; c < H2; c += blockDim.x) {
    float v = 0.f;
    for (int r0 = 0; r0 < rows; r0 += 8) {
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = r0 + u;
        const size_t row = sum_mode ? (size_t)g * rows + r : ((size_t)g * group + r) * T + T - 1;
        x[u] = r < rows ? OUT[row * H2 + c] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) v += x[u];
    }
    sm_head[c] = v;
    AGG[(size_t)g * H2 + c] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int d0 = warp * 4; d0 < D; d0 += nw * 4) {         // 4 outputs per warp pass: their weight rows load together
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = lane; c < H2; c += 32) {
      const float a = sm_head[c];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (d0 + u < D) acc[u] = fmaf(a, __ldg(W + (size_t)(d0 + u) * H2 + c), acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float t = warp_sum(acc[u]);
      if (lane == 0 && d0 + u < D) EMB[(size_t)g * D + d0 + u] = t + (float)group * bias[d0 + u];
    }
  }
}

// backward of the head for patch g: dAGG[c] = sum_d dEMB[g][d] W[d][c]; dOUT[(g, w)][t][:] = dAGG for every t ('sum') or only
// t = T-1 ('last', zero elsewhere); db[d] += group * dEMB[g][d].  (dW = dEMB^T AGG is a separate GEMM off the critical chain.)
__global__ void __launch_bounds__(512) lstm_head_bwd_kernel(const float* __restrict__ dEMB, const float* __restrict__ W, float* __restrict__ dOUT,
                                                            float* __restrict__ db, int n_groups, int group, int T, int H2, int D,
                                                            int sum_mode) {
  extern __shared__ float sm_head[];
  float* dy = sm_head;          // [D]
  float* dagg = sm_head + D;    // [H2]
  const int g = blockIdx.x;
  // thread (c, part): the D reduction is split over blockDim / H2 parts; the thread's weights (a parameter, <= 16 values) are
  // loaded BEFORE the dependency wait, i.e. while the kernel that produces dEMB is still running
  const int parts = max(1, (int)blockDim.x / H2);
  const int c = threadIdx.x % H2, part = threadIdx.x / H2;
  const int d_lo = (int)((long long)D * part / parts), d_hi = (int)((long long)D * (part + 1) / parts);
  const bool pre = part < parts && d_hi - d_lo <= 16;
  float wreg[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) wreg[i] = (pre && d_lo + i < d_hi) ? __ldg(W + (size_t)(d_lo + i) * H2 + c) : 0.f;
  sg_pdl_sync();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float v = dEMB[(size_t)g * D + d];
    dy[d] = v;
    if (db && v != 0.f) atomicAdd(db + d, (float)group * v);
  }
  for (int c2 = threadIdx.x; c2 < H2; c2 += blockDim.x) dagg[c2] = 0.f;
  __syncthreads();
  {
    if (part < parts) {
      float acc = 0.f;
      if (pre) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc = fmaf(dy[min(d_lo + i, D - 1)], wreg[i], acc);      // zero-padded tail: exact
      } else {
#pragma unroll 16
        for (int d = d_lo; d < d_hi; ++d) acc = fmaf(dy[d], __ldg(W + (size_t)d * H2 + c), acc);
      }
      if (parts == 1) dagg[c] = acc; else atomicAdd(dagg + c, acc);
    }
    if (parts == 1)
      for (int c2 = threadIdx.x + blockDim.x; c2 < H2; c2 += blockDim.x) {
        float acc = 0.f;
        for (int d = 0; d < D; ++d) acc = fmaf(dy[d], __ldg(W + (size_t)d * H2 + c2), acc);
        dagg[c2] = acc;
      }
  }
  __syncthreads();
  const int rows = group * T;
  float* base = dOUT + (size_t)g * rows * H2;
  if ((H2 & 3) == 0 && ((((size_t)base) & 15) == 0)) {           // 16-byte stores
    const int h4 = H2 >> 2;
    for (int e = threadIdx.x; e < rows * h4; e += blockDim.x) {
      const int c4 = (e % h4) * 4, t = (e / h4) % T;
      const bool live = sum_mode == 1 || t == T - 1;
      if (!live && sum_mode == SUBGNN_HEAD_LAST_NO_FILL) continue;
      const float4 v = live ? make_float4(dagg[c4], dagg[c4 + 1], dagg[c4 + 2], dagg[c4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(base)[e] = v;
    }
  } else {
    for (int e = threadIdx.x; e < rows * H2; e += blockDim.x) {
      const int c2 = e % H2, t = (e / H2) % T;
      const bool live = sum_mode == 1 || t == T - 1;
      if (!live && sum_mode == SUBGNN_HEAD_LAST_NO_FILL) continue;
      base[e] = live ? dagg[c2] : 0.f;
    }
  }
}

// inter-layer dropout (nn.LSTM dropout=p, training only): y = x * keep/(1-p); the same kernel applies the mask
// to gradients.  Counter-based mask: element index within the buffer, salt = layer / step counter.
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float p, unsigned long long seed,
                               unsigned salt, const int* __restrict__ step_dev) {
  sg_pdl_sync();
  if (step_dev) salt += 64u * (unsigned)*step_dev + 0x80000000u;
  // one Philox4x32 draw serves the four elements 4q .. 4q+3 (sg_dropout_scale uses word (idx & 3) of draw idx >> 2)
  const float keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const long long nq = n >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq && vec; q += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(x)[q];
    if (p > 0.f) {
      const Philox4 r = sg_draw(seed, (uint64_t)q, salt, SG_TAG_DROP);
      v.x *= sg_unit(r.x) >= p ? keep : 0.f;
      v.y *= sg_unit(r.y) >= p ? keep : 0.f;
      v.z *= sg_unit(r.z) >= p ? keep : 0.f;
      v.w *= sg_unit(r.w) >= p ? keep : 0.f;
    }
    reinterpret_cast<float4*>(y)[q] = v;
  }
  for (long long e = (vec ? nq * 4 : 0) + blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    y[e] = x[e] * sg_dropout_scale(seed, salt, (uint64_t)e, p);
}

// per-step weight staging: bsum[dir][j] = b_ih[dir][j] + b_hh[dir][j] and the transposed recurrent weights, either
//   cl == 0: WhhT[dir][k][j] = Whh[dir][j][k]                                   (streaming kernels above), or
//   cl >= 1: Wp[dir][rank][k][half][p][g2][o] = Whh[dir][(2 half + g2)*H + rank*U + 2p + o][k], U = H / cl   (lstm_reg.cu: every
//            CTA's slice is one contiguous block for its bulk copy; a thread's two float4 per k are contiguous across the warp).
__device__ __forceinline__ void lstm_prep_body(const float* __restrict__ Whh, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                               float* __restrict__ WhhT, float* __restrict__ bsum, int H, int cl) {
  const int H4 = 4 * H;
  const long long total = (long long)2 * H4 * H;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int dir = (int)(e / ((long long)H4 * H));
    const int r = (int)(e % ((long long)H4 * H));
    if (cl == 0) {
      const int k = r / H4, j = r % H4;
      WhhT[e] = Whh[((size_t)dir * H4 + j) * H + k];
    } else {
      const int U = H / cl, U4 = 4 * U;
      const int rank = r / (H * U4), rr = r % (H * U4);
      const int k = rr / U4, c = rr % U4;
      const int half = c / (2 * U), p = (c % (2 * U)) / 4, gate = 2 * half + (c % 4) / 2, o = c % 2;
      WhhT[e] = Whh[((size_t)dir * H4 + gate * H + rank * U + 2 * p + o) * H + k];
    }
    if (e < 2 * H4) bsum[e] = b_ih[e] + b_hh[e];
  }
}
__global__ void lstm_prep_kernel(const float* __restrict__ Whh, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                 float* __restrict__ WhhT, float* __restrict__ bsum, int H, int cl) {
  sg_pdl_wait_only();        // the recurrence kernels stage WhhT before their dependency wait (lstm_reg.cu)
  lstm_prep_body(Whh, b_ih, b_hh, WhhT, bsum, H, cl);
}
// every layer of the stack in ONE launch (grid.y = layer): the packing sits in front of the first projection on the step's chain
#define LSTM_PREP_MAX_LAYERS 8
struct LstmPrepLayers {
  const float* whh[LSTM_PREP_MAX_LAYERS];
  const float* b_ih[LSTM_PREP_MAX_LAYERS];
  const float* b_hh[LSTM_PREP_MAX_LAYERS];
  float* whh_t[LSTM_PREP_MAX_LAYERS];
  float* bsum[LSTM_PREP_MAX_LAYERS];
};
__global__ void lstm_prep_layers_kernel(LstmPrepLayers a, int H, int cl) {
  sg_pdl_wait_only();
  const int l = blockIdx.y;
  lstm_prep_body(a.whh[l], a.b_ih[l], a.b_hh[l], a.whh_t[l], a.bsum[l], H, cl);
}

extern "C" {

static int lstm_block(int H) { return ((4 * H + 31) / 32) * 32; }

int subgnn_lstm_prep(const float* whh, const float* b_ih, const float* b_hh, float* whh_t, float* bsum, int H, void* stream) {
  SG_REQUIRE(H >= 1 && H <= 256, "hidden size must be in [1, 256]");
  sg_launch_pdl(lstm_prep_kernel, dim3(sg_grid_for((long long)8 * H * H, 256, 4)), dim3(256), 0, (cudaStream_t)stream, whh, b_ih, b_hh, whh_t, bsum, H,
                                                                                                  lstm_reg_supported(H) ? lstm_reg_cluster(H) : 0);
  return subgnn_check_launch("lstm_prep_kernel");
}

int subgnn_lstm_prep_layers(const float* const* whh, const float* const* b_ih, const float* const* b_hh, float* const* whh_t,
                            float* const* bsum, int n_layers, int H, void* stream) {
  SG_REQUIRE(H >= 1 && H <= 256, "hidden size must be in [1, 256]");
  SG_REQUIRE(n_layers >= 1 && n_layers <= LSTM_PREP_MAX_LAYERS, "1 .. 8 layers per launch");
  LstmPrepLayers a = {};
  for (int l = 0; l < n_layers; ++l) { a.whh[l] = whh[l]; a.b_ih[l] = b_ih[l]; a.b_hh[l] = b_hh[l]; a.whh_t[l] = whh_t[l]; a.bsum[l] = bsum[l]; }
  sg_launch_pdl(lstm_prep_layers_kernel, dim3(sg_grid_for((long long)8 * H * H, 256, 4), n_layers), dim3(256), 0, (cudaStream_t)stream, a, H,
                lstm_reg_supported(H) ? lstm_reg_cluster(H) : 0);
  return subgnn_check_launch("lstm_prep_layers_kernel");
}

int subgnn_lstm_recur_fwd(float* G, const float* whh_t, float* OUT, float* CS, int n_seq, int T, int H, int steps_fwd, int steps_rev,
                          void* stream) {
  SG_REQUIRE(H >= 1 && H <= 256 && n_seq >= 0 && T >= 1, "bad sizes");
  SG_REQUIRE(steps_fwd >= 0 && steps_fwd <= T && steps_rev >= 0 && steps_rev <= T, "bad step counts");
  if (n_seq == 0) return SUBGNN_OK;
  if (lstm_reg_supported(H)) return lstm_reg_fwd(G, whh_t, OUT, CS, n_seq, T, H, steps_fwd, steps_rev, nullptr, 0.f, 0ull, 0u, nullptr, (cudaStream_t)stream);
  size_t smem = (size_t)(2 * S_TILE * H + S_TILE * 4 * H) * sizeof(float);
  const size_t wbytes = (size_t)4 * H * H * sizeof(float);
  const bool w_smem = smem + wbytes <= 100 * 1024 && (H % 2 == 0) && (((size_t)whh_t) & 15) == 0;   // two CTAs per SM stay resident; bulk copy: 16-byte granules
  dim3 grid(sg_div_up(n_seq, S_TILE), 2);
  const int block = lstm_block(H);
  cudaStream_t st = (cudaStream_t)stream;
#define LSTM_FWD_LAUNCH(WS, BL)                                                                                               \
  do {                                                                                                                        \
    cudaFuncSetAttribute(lstm_recur_fwd_kernel<WS, BL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    lstm_recur_fwd_kernel<WS, BL><<<grid, block, smem, st>>>(G, whh_t, OUT, CS, n_seq, T, H, steps_fwd, steps_rev);           \
  } while (0)
  if (w_smem) {
    smem += wbytes;
    LSTM_FWD_LAUNCH(true, 256);                               // w_smem implies H <= 64, i.e. at most 256 threads
  } else if (block <= 256) LSTM_FWD_LAUNCH(false, 256);
  else if (block <= 512) LSTM_FWD_LAUNCH(false, 512);
  else LSTM_FWD_LAUNCH(false, 1024);
#undef LSTM_FWD_LAUNCH
  return subgnn_check_launch("lstm_recur_fwd_kernel");
}

int subgnn_lstm_fused_dropout_supported(int H) { return lstm_reg_supported(H) ? 1 : 0; }

int subgnn_lstm_recur_fwd_drop(float* G, const float* whh_t, float* OUT, float* CS, int n_seq, int T, int H, int steps_fwd, int steps_rev,
                               float* xdrop, float p, unsigned long long seed, unsigned salt, const int* step_dev, void* stream) {
  SG_REQUIRE(lstm_reg_supported(H), "fused inter-layer dropout needs the register-tiled recurrence (H % 8 == 0, H <= 128)");
  SG_REQUIRE(n_seq >= 0 && T >= 1 && p >= 0.f && p < 1.f, "bad sizes");
  SG_REQUIRE(steps_fwd >= 0 && steps_fwd <= T && steps_rev >= 0 && steps_rev <= T, "bad step counts");
  if (n_seq == 0) return SUBGNN_OK;
  return lstm_reg_fwd(G, whh_t, OUT, CS, n_seq, T, H, steps_fwd, steps_rev, xdrop, p, seed, salt, step_dev, (cudaStream_t)stream);
}

int subgnn_lstm_recur_bwd_drop(float* G, const float* whh, const float* OUT, const float* CS, const float* dOUT, int n_seq, int T, int H,
                               int steps_fwd, int steps_rev, int zero_untaken, float* db_ih, float* db_hh, float p, unsigned long long seed,
                               unsigned salt, const int* step_dev, void* stream) {
  SG_REQUIRE(lstm_reg_supported(H), "fused inter-layer dropout needs the register-tiled recurrence (H % 8 == 0, H <= 128)");
  SG_REQUIRE(n_seq >= 0 && T >= 1 && p >= 0.f && p < 1.f, "bad sizes");
  if (n_seq == 0) return SUBGNN_OK;
  return lstm_reg_bwd(G, whh, OUT, CS, dOUT, nullptr, n_seq, T, H, steps_fwd, steps_rev, zero_untaken, db_ih, db_hh, p, seed, salt, step_dev,
                      (cudaStream_t)stream);
}

int subgnn_lstm_recur_bwd_add(float* G, const float* whh, const float* OUT, const float* CS, const float* dOUT, const float* dOUT_add, int n_seq,
                              int T, int H, int steps_fwd, int steps_rev, int zero_untaken, float* db_ih, float* db_hh, float p,
                              unsigned long long seed, unsigned salt, const int* step_dev, void* stream) {
  SG_REQUIRE(lstm_reg_supported(H), "the two-source BPTT needs the register-tiled recurrence (H % 8 == 0, H <= 128)");
  SG_REQUIRE(n_seq >= 0 && T >= 1 && p >= 0.f && p < 1.f, "bad sizes");
  if (n_seq == 0) return SUBGNN_OK;
  return lstm_reg_bwd(G, whh, OUT, CS, dOUT, dOUT_add, n_seq, T, H, steps_fwd, steps_rev, zero_untaken, db_ih, db_hh, p, seed, salt, step_dev,
                      (cudaStream_t)stream);
}

int subgnn_lstm_recur_bwd(float* G, const float* whh, const float* OUT, const float* CS, const float* dOUT, int n_seq, int T, int H,
                          int steps_fwd, int steps_rev, int zero_untaken, float* db_ih, float* db_hh, void* stream) {
  SG_REQUIRE(H >= 1 && H <= 256 && n_seq >= 0 && T >= 1, "bad sizes");
  if (n_seq == 0) return SUBGNN_OK;
  if (lstm_reg_supported(H))
    return lstm_reg_bwd(G, whh, OUT, CS, dOUT, nullptr, n_seq, T, H, steps_fwd, steps_rev, zero_untaken, db_ih, db_hh, 0.f, 0ull, 0u, nullptr,
                        (cudaStream_t)stream);
  SG_REQUIRE(!(zero_untaken & SUBGNN_LSTM_DOUT_LAST_ONLY), "SUBGNN_LSTM_DOUT_LAST_ONLY needs the register-tiled recurrence (H % 8 == 0, H <= 128)");
  zero_untaken &= SUBGNN_LSTM_ZERO_UNTAKEN;
  size_t smem = (size_t)(S_TILE * 4 * H + 2 * S_TILE * H + 4 * S_TILE * H) * sizeof(float);
  const size_t wbytes = (size_t)4 * H * H * sizeof(float);
  const bool w_smem = smem + wbytes <= 100 * 1024 && (H % 2 == 0) && (((size_t)whh) & 15) == 0;
  dim3 grid(sg_div_up(n_seq, S_TILE), 2);
  const int block = lstm_block(H);
  cudaStream_t st = (cudaStream_t)stream;
#define LSTM_BWD_LAUNCH(WS, BL)                                                                                               \
  do {                                                                                                                        \
    cudaFuncSetAttribute(lstm_recur_bwd_kernel<WS, BL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    lstm_recur_bwd_kernel<WS, BL><<<grid, block, smem, st>>>(G, whh, OUT, CS, dOUT, n_seq, T, H, steps_fwd, steps_rev,        \
                                                             zero_untaken, db_ih, db_hh);                                     \
  } while (0)
  if (w_smem) {
    smem += wbytes;
    LSTM_BWD_LAUNCH(true, 256);
  } else if (block <= 256) LSTM_BWD_LAUNCH(false, 256);
  else if (block <= 512) LSTM_BWD_LAUNCH(false, 512);
  else LSTM_BWD_LAUNCH(false, 1024);
#undef LSTM_BWD_LAUNCH
  return subgnn_check_launch("lstm_recur_bwd_kernel");
}

int subgnn_lstm_head_fwd(const float* OUT, float* AGG, float* EMB, const float* W, const float* bias, int n_groups, int group, int T, int H2,
                         int D, int sum_mode, void* stream) {
  if (n_groups == 0) return SUBGNN_OK;
  sg_launch_pdl<SG_PDL_CHAIN>(lstm_head_fwd_kernel, dim3(n_groups), dim3(512), (size_t)H2 * sizeof(float), (cudaStream_t)stream, OUT, AGG, EMB, W, bias, n_groups, group, T, H2, D,
                                                                                          sum_mode);
  return subgnn_check_launch("lstm_head_fwd_kernel");
}

int subgnn_lstm_head_bwd(const float* dEMB, const float* W, float* dOUT, float* db, int n_groups, int group, int T, int H2, int D,
                         int sum_mode, void* stream) {
  if (n_groups == 0) return SUBGNN_OK;
  sg_launch_pdl<SG_PDL_CHAIN>(lstm_head_bwd_kernel, dim3(n_groups), dim3(512), (size_t)(H2 + D) * sizeof(float), (cudaStream_t)stream, dEMB, W, dOUT, db, n_groups, group, T, H2, D,
                                                                                                sum_mode);
  return subgnn_check_launch("lstm_head_bwd_kernel");
}

int subgnn_dropout(const float* x, float* y, long long n, float p, unsigned long long seed, unsigned salt, const int* step_dev,
                   void* stream) {
  if (n == 0) return SUBGNN_OK;
  sg_launch_pdl(dropout_kernel, dim3(sg_grid_for((n + 3) / 4, 256, 8)), dim3(256), 0, (cudaStream_t)stream, x, y, n, p, seed, salt, step_dev);
  return subgnn_check_launch("dropout_kernel");
}

}  // extern "C"
