// Generic fp32 smem-tiled GEMM engine shared by gemm.cu and model.cu.
#pragma once
#define BM 64
#define BN 64
#define BK 16
#define TM 4
#define TN 4

// Generic tile engine.  a(m, k) / b(k, n) fetch operand elements (bounds already checked by the caller),
// epi(m, n, acc) consumes one output element.  A_KC / B_KC: operand is contiguous along k (choose the
// thread->element map so that global loads coalesce).
template <bool A_KC, bool B_KC, class AF, class BF, class EF>
__device__ __forceinline__ void gemm_tile(int M, int N, int k0, int k1, AF a, BF b, EF epi) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int kb = k0; kb < k1; kb += BK) {
#pragma unroll
    for (int p = 0; p < (BM * BK) / 256; ++p) {
      const int e = p * 256 + tid;
      int mm, kk;
      if (A_KC) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      const int m = m0 + mm, k = kb + kk;
      As[kk][mm] = (m < M && k < k1) ? a(m, k) : 0.f;
    }
#pragma unroll
    for (int p = 0; p < (BN * BK) / 256; ++p) {
      const int e = p * 256 + tid;
      int nn, kk;
      if (B_KC) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      const int n = n0 + nn, k = kb + kk;
      Bs[kk][nn] = (n < N && k < k1) ? b(k, n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int m = m0 + ty * TM + i, n = n0 + tx * TN + j;
      if (m < M && n < N) epi(m, n, acc[i][j]);
    }
}

