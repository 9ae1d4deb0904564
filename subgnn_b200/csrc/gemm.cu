// fp32 tiled GEMMs for the dense parts of the path: MPN projection, readout MLP, LSTM input
// projections / weight gradients, with the row gathers / scatters of the embedding table fused into the
// operand loads and the epilogue.
//
// Replaces the cuBLAS/ATen calls behind (reference):
//   subgraph_mpn.py:239        F.relu(self.linear(cat[x, aggr]))           -> subgnn_linear_fwd (+ bwd)
//   SubGNN.py:306-310          lin / lin2 / lin3                            -> subgnn_linear_fwd (+ bwd)
//   SubGNN.py:73-78 nn.LSTM    x_t W_ih^T (all t at once), dW_ih, dW_hh, dX -> subgnn_linear_fwd/bwd_input/bwd_weight
//   anchor_patch_samplers.py:409 node_matrix(ids) feeding the LSTM          -> row gather fused into the A-operand load
//   embedding_dense_backward                                                -> atomic row scatter fused into the epilogue
//
// fp32 FFMA on purpose (round 1): the stated parity tolerance for the deterministic stages is fp32
// (rtol 1e-4); tf32/bf16 tcgen05 tiles would not hold it for the gradient GEMMs.
#include "common.cuh"
#include "../../include/subgnn_b200.h"

#include "gemm_tile.cuh"

// y[m][n] = act( sum_k X[row(m)][k] * W[n][k] + bias[n] )
__global__ void __launch_bounds__(256)
linear_fwd_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ ids, const float* __restrict__ w, int ldw,
                  const float* __restrict__ bias, float* __restrict__ y, int ldy, int M, int N, int K, int relu) {
  gemm_tile<true, true>(
      M, N, 0, K,
      [&](int m, int k) { const long long r = ids ? (long long)ids[m] : m; return __ldg(x + r * ldx + k); },
      [&](int k, int n) { return __ldg(w + (long long)n * ldw + k); },
      [&](int m, int n, float v) {
        if (bias) v += bias[n];
        if (relu) v = fmaxf(v, 0.f);
        y[(long long)m * ldy + n] = v;
      });
}

// dx[row(m)][k] (+)= sum_n dy[m][n] * W[n][k]     (scatter_ids: atomic row scatter, PAD row 0 skipped)
__global__ void __launch_bounds__(256)
linear_bwd_input_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ w, int ldw, float* __restrict__ dx, int lddx,
                        const int* __restrict__ scatter_ids, int M, int N, int K, int accumulate) {
  // output tile is (M x K); reduction over N
  gemm_tile<true, false>(
      M, K, 0, N,
      [&](int m, int n) { return __ldg(dy + (long long)m * ldy + n); },
      [&](int n, int k) { return __ldg(w + (long long)n * ldw + k); },
      [&](int m, int k, float v) {
        if (scatter_ids) {
          const int r = scatter_ids[m];
          if (r != 0) atomicAdd(dx + (long long)r * lddx + k, v);
        } else if (accumulate) {
          dx[(long long)m * lddx + k] += v;
        } else {
          dx[(long long)m * lddx + k] = v;
        }
      });
}

// dW[n][k] += sum_m dy[m][n] * X[row(m)][k]   (split over m across blockIdx.z, atomic accumulate)
__global__ void __launch_bounds__(256)
linear_bwd_weight_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ x, int ldx, const int* __restrict__ ids,
                         float* __restrict__ dw, int lddw, float* __restrict__ db, int M, int N, int K, int m_chunk,
                         const int* __restrict__ m_dev) {
  sg_pdl_sync();
  if (m_dev) M = min(M, *m_dev);
  const int m_beg = blockIdx.z * m_chunk;
  const int m_end = min(M, m_beg + m_chunk);
  if (m_beg >= m_end) return;
  if (db && blockIdx.x == 0) {                                   // fused bias gradient: column sums of this tile's rows of dy^T
    const int n = blockIdx.y * BM + (threadIdx.x % BM), part = threadIdx.x / BM;
    if (n < N) {
      float s = 0.f;
      for (int m = m_beg + part; m < m_end; m += 256 / BM) s += dy[(long long)m * ldy + n];
      if (s != 0.f) atomicAdd(db + n, s);
    }
  }
  gemm_tile<false, false>(
      N, K, m_beg, m_end,
      [&](int n, int m) { return __ldg(dy + (long long)m * ldy + n); },
      [&](int m, int k) { const long long r = ids ? (long long)ids[m] : m; return __ldg(x + r * ldx + k); },
      [&](int n, int k, float v) { atomicAdd(dw + (long long)n * lddw + k, v); });
}

// db[n] += sum_m dy[m][n]
__global__ void colsum_kernel(const float* __restrict__ dy, int ldy, float* __restrict__ db, int M, int N, int m_chunk,
                              const int* __restrict__ m_dev) {
  if (m_dev) M = min(M, *m_dev);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m_beg = blockIdx.y * m_chunk, m_end = min(M, m_beg + m_chunk);
  float s = 0.f;
  for (int m = m_beg; m < m_end; ++m) s += dy[(long long)m * ldy + n];
  if (m_beg < m_end) atomicAdd(db + n, s);
}

extern "C" {

int subgnn_linear_fwd(const float* x, int ldx, const int* gather_ids, const float* w, int ldw, const float* bias, float* y, int ldy,
                      int M, int N, int K, int relu, void* stream) {
  SG_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad sizes");
  if (M == 0) return SUBGNN_OK;
  dim3 grid(sg_div_up(N, BN), sg_div_up(M, BM));
  linear_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, gather_ids, w, ldw, bias, y, ldy, M, N, K, relu);
  return subgnn_check_launch("linear_fwd_kernel");
}

int subgnn_linear_bwd_input(const float* dy, int ldy, const float* w, int ldw, float* dx, int lddx, const int* scatter_ids, int M, int N,
                            int K, int accumulate, void* stream) {
  SG_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad sizes");
  if (M == 0) return SUBGNN_OK;
  dim3 grid(sg_div_up(K, BN), sg_div_up(M, BM));
  linear_bwd_input_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, ldy, w, ldw, dx, lddx, scatter_ids, M, N, K, accumulate);
  return subgnn_check_launch("linear_bwd_input_kernel");
}

int subgnn_linear_bwd_weight(const float* dy, int ldy, const float* x, int ldx, const int* gather_ids, float* dw, int lddw, float* db,
                             int M, int N, int K, const int* m_dev, void* stream) {
  SG_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad sizes");
  if (M == 0) return SUBGNN_OK;
  // split the reduction so that the grid covers the machine: tiles * splits ~ 2 waves
  const int tiles = sg_div_up(K, BN) * sg_div_up(N, BM);
  int splits = (2 * subgnn_sm_count() + tiles - 1) / tiles;
  const int max_splits = sg_div_up(M, 4 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int m_chunk = sg_div_up(sg_div_up(M, splits), BK) * BK;
  splits = sg_div_up(M, m_chunk);
  dim3 grid(sg_div_up(K, BN), sg_div_up(N, BM), splits);
  sg_launch_pdl(linear_bwd_weight_kernel, grid, dim3(256), 0, (cudaStream_t)stream, dy, ldy, x, ldx, gather_ids, dw, lddw, db, M, N, K, m_chunk, m_dev);
  return subgnn_check_launch("linear_bwd_weight_kernel");
}

int subgnn_colsum(const float* dy, int ldy, float* db, int M, int N, const int* m_dev, void* stream) {
  if (M == 0) return SUBGNN_OK;
  const int cs = sg_div_up(M, 64) > 64 ? 64 : sg_div_up(M, 64);
  const int chunk = sg_div_up(M, cs);
  dim3 g2(sg_div_up(N, 128), sg_div_up(M, chunk));
  colsum_kernel<<<g2, 128, 0, (cudaStream_t)stream>>>(dy, ldy, db, M, N, chunk, m_dev);
  return subgnn_check_launch("colsum_kernel");
}

}  // extern "C"
