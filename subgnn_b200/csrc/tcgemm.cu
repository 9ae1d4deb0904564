// tcgen05 (5th-gen tensor core) GEMMs for the walk-encoder LSTM projections, fp32-accurate via 3xTF32.
//
// Replaces the cuBLAS sgemm behind nn.LSTM's input projections and their gradients (reference SubGNN.py:73-78,
// called from anchor_patch_samplers.py:429) on B200: tcgen05.mma.kind::tf32 with the accumulator in TMEM.
// fp32 parity is kept by error-compensated splitting: every operand x is staged in shared memory as
//   x_hi = x with the 13 low mantissa bits cleared (exactly what the TF32 datapath reads)
//   x_lo = x - x_hi                                 (exact in fp32)
// and D += A_hi B_hi + A_lo B_hi + A_hi B_lo (fp32 accumulate in TMEM); the dropped lo*lo term is ~2^-22 relative.
//
// Operand staging is done by the CTA's threads (generic-proxy st.shared into the canonical K-major SWIZZLE_128B
// layout, then fence.proxy.async) instead of TMA because the A operand is a row GATHER of the embedding table
// (fused anchor_patch_samplers.py:409 lookup) and the hi/lo split has to be applied on the way in.
// Layout facts used (cute/arch/mma_sm100_desc.hpp, cute/atom/mma_traits_sm100.hpp of the vendored CUTLASS):
//   K-major SWIZZLE_128B tile: row r = 128 bytes (32 fp32 of K); 16-byte chunk c of row r lives at chunk c ^ (r & 7);
//   8-row groups are 1024 B apart (SBO = 64 in 16-B units), descriptor version 1, layout_type 2; one MMA consumes
//   K = 8 tf32 = 32 B, so MMA k of a 32-wide K block starts 2*k 16-B units further.
//   Accumulator of an M = 128, cta_group::1 MMA: row m <-> TMEM lane m, column n <-> TMEM column n.
#include "common.cuh"
#include "../../include/subgnn_b200.h"

#define TC_M 128
#define TC_KB 32          // K elements per staged block (128 bytes per row)
#define TC_THREADS 128

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major; canonical value 1)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8-row groups are 1024 B apart
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // D=F32, A=B=TF32, K-major
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// split a float4 into TF32-exact high parts and fp32 remainders
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
}

// byte offset of (row r, 16-byte chunk c) inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// stage ROWS x 32 floats (K-contiguous source rows) into hi / lo tiles; fetch(row, k) -> const float* or nullptr
template <int ROWS, class RowPtr>
__device__ __forceinline__ void stage_kmajor(unsigned char* hi, unsigned char* lo, int k0, int K, RowPtr rowptr) {
#pragma unroll
  for (int i = 0; i < (ROWS * 8) / TC_THREADS; ++i) {
    const int idx = i * TC_THREADS + threadIdx.x;
    const int r = idx >> 3, c = idx & 7;
    const int k = k0 + c * 4;
    const float* src = rowptr(r);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src) {
      if (k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(src + k));
      else {
        if (k < K) v.x = src[k];
        if (k + 1 < K) v.y = src[k + 1];
        if (k + 2 < K) v.z = src[k + 2];
      }
    }
    float4 h, l;
    split4(v, h, l);
    const uint32_t off = sw128(r, c);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

// ------------------------------------------------------------------------------------------------
// y[m][n] = act(sum_k X[row(m)][k] W[n][k] + bias[n])      (same contract as subgnn_linear_fwd)
template <int N_TILE>
__global__ void __launch_bounds__(TC_THREADS)
tc_linear_fwd_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ ids, const float* __restrict__ w, int ldw,
                     const float* __restrict__ bias, float* __restrict__ y, int ldy, int M, int N, int K, int relu) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* a_hi = base;
  unsigned char* a_lo = a_hi + TC_M * 128;
  unsigned char* b_hi = a_lo + TC_M * 128;
  unsigned char* b_lo = b_hi + N_TILE * 128;
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_M, n0 = blockIdx.x * N_TILE;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(N_TILE) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    mbar_init(&mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = umma_idesc_tf32(TC_M, N_TILE);
  const int n_kb = (K + TC_KB - 1) / TC_KB;
  for (int kb = 0; kb < n_kb; ++kb) {
    if (kb > 0) mbar_wait(&mma_bar, (uint32_t)((kb - 1) & 1));        // MMAs of the previous block have consumed the tiles
    stage_kmajor<TC_M>(a_hi, a_lo, kb * TC_KB, K, [&](int r) -> const float* {
      const int m = m0 + r;
      if (m >= M) return nullptr;
      const long long row = ids ? (long long)ids[m] : m;
      return x + row * ldx;
    });
    stage_kmajor<N_TILE>(b_hi, b_lo, kb * TC_KB, K, [&](int r) -> const float* {
      const int n = n0 + r;
      return n < N ? w + (long long)n * ldw : nullptr;
    });
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t dah = umma_desc_k_sw128(smem_u32(a_hi)), dal = umma_desc_k_sw128(smem_u32(a_lo));
      const uint64_t dbh = umma_desc_k_sw128(smem_u32(b_hi)), dbl = umma_desc_k_sw128(smem_u32(b_lo));
#pragma unroll
      for (int ks = 0; ks < TC_KB / 8; ++ks) {
        const uint64_t adv = (uint64_t)(2 * ks);                        // 32 bytes per K=8 step, in 16-byte units
        umma_tf32(tmem_d, dah + adv, dbh + adv, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
        umma_tf32(tmem_d, dal + adv, dbh + adv, idesc, 1u);
        umma_tf32(tmem_d, dah + adv, dbl + adv, idesc, 1u);
      }
      umma_commit(&mma_bar);                                            // arrives when every MMA issued so far has completed
    }
  }
  mbar_wait(&mma_bar, (uint32_t)((n_kb - 1) & 1));
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w owns TMEM lanes (= output rows) 32w .. 32w+31, 32 columns per tcgen05.ld
  const int m = m0 + warp * 32 + lane;
#pragma unroll 1
  for (int c0 = 0; c0 < N_TILE; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (m < M) {
      float* dst = y + (long long)m * ldy + n0 + c0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = n0 + c0 + j;
        if (n < N) {
          float v = __uint_as_float(r[j]);
          if (bias) v += bias[n];
          if (relu) v = fmaxf(v, 0.f);
          dst[j] = v;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(N_TILE) : "memory");
}

extern "C" {

int subgnn_tc_linear_fwd(const float* x, int ldx, const int* gather_ids, const float* w, int ldw, const float* bias, float* y, int ldy,
                         int M, int N, int K, int relu, void* stream) {
  SG_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad sizes");
  SG_REQUIRE((ldx % 4) == 0 && (ldw % 4) == 0 && (((size_t)x | (size_t)w) & 15) == 0, "tensor-core path needs 16-byte aligned rows");
  if (M == 0) return SUBGNN_OK;
  constexpr int NT = 128;
  const size_t smem = (size_t)(2 * TC_M * 128 + 2 * NT * 128) + 1024;
  cudaFuncSetAttribute(tc_linear_fwd_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(sg_div_up(N, NT), sg_div_up(M, TC_M));
  tc_linear_fwd_kernel<NT><<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(x, ldx, gather_ids, w, ldw, bias, y, ldy, M, N, K, relu);
  return subgnn_check_launch("tc_linear_fwd_kernel");
}

}  // extern "C"
