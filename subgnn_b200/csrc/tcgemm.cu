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
#include <cstdlib>
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

// ------------------------------------------------------------------------------------------------
// Operand stagers.  A stager first ISSUES all of a thread's global loads for one 32-wide reduction block into
// registers (load), and later splits them into hi / lo and writes the SWIZZLE_128B tiles (store).  Keeping the two
// phases apart lets the tile loop issue block kb+1's loads before it waits on block kb's MMAs, and keeps the 8-16
// loads of one thread in flight together instead of one load -> use chain per row (the round-1 profile showed the
// kernels stalled on exactly those chains).

// K-major source: tile row r is a K-contiguous row in memory.  rowptr(r) -> const float* or nullptr (zero row).
template <int ROWS>
struct KMajorStager {
  static constexpr int IT = (ROWS * 8) / TC_THREADS;
  const float* rp[IT];
  float4 v[IT];
  int K;
  template <class RowPtr>
  __device__ __forceinline__ void init(int K_, RowPtr rowptr) {
    K = K_;
#pragma unroll
    for (int i = 0; i < IT; ++i) rp[i] = rowptr((i * TC_THREADS + (int)threadIdx.x) >> 3);
  }
  __device__ __forceinline__ void load(int k0) {
    const int k = k0 + ((int)threadIdx.x & 7) * 4;
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rp[i]) {
        if (k + 3 < K) t = __ldg(reinterpret_cast<const float4*>(rp[i] + k));
        else {
          if (k < K) t.x = rp[i][k];
          if (k + 1 < K) t.y = rp[i][k + 1];
          if (k + 2 < K) t.z = rp[i][k + 2];
        }
      }
      v[i] = t;
    }
  }
  __device__ __forceinline__ void store(unsigned char* hi, unsigned char* lo) const {
    const int c = (int)threadIdx.x & 7;
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      const int r = (i * TC_THREADS + (int)threadIdx.x) >> 3;
      float4 h, l;
      split4(v[i], h, l);
      const uint32_t off = sw128(r, c);
      *reinterpret_cast<float4*>(hi + off) = h;
      *reinterpret_cast<float4*>(lo + off) = l;
    }
  }
};

// "MN-major" source: for reduction index kk the tile rows are contiguous in memory (colptr(k) -> pointer to tile row 0
// at absolute reduction index k, or nullptr).  Lanes run along the reduction index, so the 4-byte shared-memory stores
// of a warp fall into one 128-byte row: conflict free; the 16-byte global loads of neighbouring row chunks share
// 32-byte sectors through L1.
template <int ROWS>
struct MNMajorStager {
  static constexpr int IT = (ROWS * 8) / TC_THREADS;
  float4 v[IT];
  template <class ColPtr>
  __device__ __forceinline__ void load(int k0, int rows_valid, ColPtr colptr) {
    const int kk = (int)threadIdx.x & 31;
    const float* src = colptr(k0 + kk);
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      const int r0 = (i * (TC_THREADS / 32) + ((int)threadIdx.x >> 5)) * 4;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src) {
        if (r0 + 3 < rows_valid) t = __ldg(reinterpret_cast<const float4*>(src + r0));
        else {
          if (r0 < rows_valid) t.x = src[r0];
          if (r0 + 1 < rows_valid) t.y = src[r0 + 1];
          if (r0 + 2 < rows_valid) t.z = src[r0 + 2];
        }
      }
      v[i] = t;
    }
  }
  __device__ __forceinline__ void store(unsigned char* hi, unsigned char* lo) const {
    const int kk = (int)threadIdx.x & 31;
    const int c = kk >> 2, w4 = (kk & 3) * 4;
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      const int r0 = (i * (TC_THREADS / 32) + ((int)threadIdx.x >> 5)) * 4;
      float4 h, l;
      split4(v[i], h, l);
      *reinterpret_cast<float*>(hi + sw128(r0, c) + w4) = h.x;     *reinterpret_cast<float*>(lo + sw128(r0, c) + w4) = l.x;
      *reinterpret_cast<float*>(hi + sw128(r0 + 1, c) + w4) = h.y; *reinterpret_cast<float*>(lo + sw128(r0 + 1, c) + w4) = l.y;
      *reinterpret_cast<float*>(hi + sw128(r0 + 2, c) + w4) = h.z; *reinterpret_cast<float*>(lo + sw128(r0 + 2, c) + w4) = l.z;
      *reinterpret_cast<float*>(hi + sw128(r0 + 3, c) + w4) = h.w; *reinterpret_cast<float*>(lo + sw128(r0 + 3, c) + w4) = l.w;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Tile engine: D[128 x N_TILE] = sum over reduction blocks [kr0, kr1) of A_tile B_tile^T, 3xTF32, accumulator in TMEM.
// load_a(k0) / load_b(k0) issue the global loads of the 32-wide reduction block starting at k0 into the stagers'
// registers, store_a / store_b write the operand tiles; epi(row_in_tile, col_in_tile, float4 of 4 consecutive columns)
// consumes the result.  The epilogue transposes each warp's 32 x 32 TMEM slab through shared memory so that 8 lanes
// cover 128 contiguous bytes of one output row (coalesced 16-byte stores / vector atomics).
#define TC_EPI_LD 36      // padded row length (floats) of the per-warp transposition slab: conflict-free float4 in both directions
// STAGES = 2 double-buffers the operand tiles: the threads split / store block kb+1 while the tensor core still works on block
// kb (the single-buffered loop serialises load latency, store and MMA of every block: ~1.5 us per 32-wide block measured
// in the step timeline, profiles/r01_timeline_ppi_bp_v10.txt); STAGES = 1 keeps 3 CTAs per SM for the short reductions.
template <int N_TILE, int STAGES, class LoadA, class LoadB, class StoreA, class StoreB, class Epi>
__device__ __forceinline__ void tc_tile(int kr0, int kr1, LoadA load_a, LoadB load_b, StoreA store_a, StoreB store_b, Epi epi) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int STAGE_BYTES = 2 * TC_M * 128 + 2 * N_TILE * 128;
  __shared__ uint64_t mma_bar[STAGES];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kb = (kr1 - kr0 + TC_KB - 1) / TC_KB;
  if (n_kb > 0) { load_a(kr0); load_b(kr0); }                           // first block's loads fly during the TMEM allocation
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(N_TILE < 32 ? 32 : N_TILE) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) mbar_init(&mma_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = umma_idesc_tf32(TC_M, N_TILE);
  for (int kb = 0; kb < n_kb; ++kb) {
    const int stg = kb % STAGES;
    unsigned char* a_hi = base + stg * STAGE_BYTES;
    unsigned char* a_lo = a_hi + TC_M * 128;
    unsigned char* b_hi = a_lo + TC_M * 128;
    unsigned char* b_lo = b_hi + N_TILE * 128;
    // the MMAs of block kb - STAGES (the previous user of this buffer, completion kb/STAGES - 1 of its barrier) have consumed the tiles
    if (kb >= STAGES) mbar_wait(&mma_bar[stg], (uint32_t)((kb / STAGES - 1) & 1));
    store_a(a_hi, a_lo);
    store_b(b_hi, b_lo);
    if (kb + 1 < n_kb) { load_a(kr0 + (kb + 1) * TC_KB); load_b(kr0 + (kb + 1) * TC_KB); }   // in flight while the tensor core works
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
      umma_commit(&mma_bar[stg]);                                       // arrives when every MMA issued so far has completed
    }
  }
  if (n_kb > 0) mbar_wait(&mma_bar[(n_kb - 1) % STAGES], (uint32_t)(((n_kb - 1) / STAGES) & 1));
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w owns TMEM lanes (= tile rows) 32w .. 32w+31, 32 columns per tcgen05.ld; the operand tiles are free now
  float* slab = reinterpret_cast<float*>(base) + warp * (32 * TC_EPI_LD);
#pragma unroll 1
  for (int c0 = 0; c0 < N_TILE && n_kb > 0; c0 += 32) {
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
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(slab + lane * TC_EPI_LD + j) =
          make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = it * 4 + (lane >> 3), cc = (lane & 7) * 4;
      if (c0 + cc < N_TILE) epi(warp * 32 + row, c0 + cc, *reinterpret_cast<const float4*>(slab + row * TC_EPI_LD + cc));
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(N_TILE < 32 ? 32 : N_TILE) : "memory");
}

// y[m][n] = act(sum_k X[row(m)][k] W[n][k] + bias[n])      (same contract as subgnn_linear_fwd)
template <int N_TILE, int STAGES>
__global__ void __launch_bounds__(TC_THREADS)
tc_linear_fwd_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ ids, const float* __restrict__ w, int ldw,
                     const float* __restrict__ bias, float* __restrict__ y, int ldy, int M, int N, int K, int relu) {
  sg_pdl_sync();
  __shared__ __align__(16) float s_bias[N_TILE];
  const int m0 = blockIdx.y * TC_M, n0 = blockIdx.x * N_TILE;
  for (int i = threadIdx.x; i < N_TILE; i += TC_THREADS) s_bias[i] = (bias && n0 + i < N) ? bias[n0 + i] : 0.f;   // visible after tc_tile's barriers
  KMajorStager<TC_M> sa;
  KMajorStager<N_TILE> sb;
  sa.init(K, [&](int r) -> const float* {
    const int m = m0 + r;
    if (m >= M) return nullptr;
    const long long row = ids ? (long long)ids[m] : m;
    return x + row * ldx;
  });
  sb.init(K, [&](int r) -> const float* { return n0 + r < N ? w + (long long)(n0 + r) * ldw : nullptr; });
  tc_tile<N_TILE, STAGES>(
      0, K, [&](int k0) { sa.load(k0); }, [&](int k0) { sb.load(k0); },
      [&](unsigned char* hi, unsigned char* lo) { sa.store(hi, lo); }, [&](unsigned char* hi, unsigned char* lo) { sb.store(hi, lo); },
      [&](int r, int c, float4 v) {
        const int m = m0 + r, n = n0 + c;
        if (m >= M || n >= N) return;
        float* dst = y + (long long)m * ldy + n;
        const float4 bq = *reinterpret_cast<const float4*>(s_bias + c);
        float t[4] = {v.x + bq.x, v.y + bq.y, v.z + bq.z, v.w + bq.w};
        if (relu) {
#pragma unroll
          for (int j = 0; j < 4; ++j) t[j] = fmaxf(t[j], 0.f);
        }
        if (n + 3 < N && (((size_t)dst) & 15) == 0) *reinterpret_cast<float4*>(dst) = make_float4(t[0], t[1], t[2], t[3]);
        else for (int j = 0; j < 4 && n + j < N; ++j) dst[j] = t[j];
      });
}

// dx[row(m)][k] (+)= sum_n dy[m][n] W[n][k]       (same contract as subgnn_linear_bwd_input)
// grid.z splits the reduction over n; with more than one split every output is added atomically (dx must then hold the
// value to accumulate onto: the host wrapper only splits in scatter / accumulate mode).
template <int N_TILE, int STAGES>
__global__ void __launch_bounds__(TC_THREADS)
tc_linear_bwd_input_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ w, int ldw, float* __restrict__ dx, int lddx,
                           const int* __restrict__ scatter_ids, int M, int N, int K, int accumulate, int n_chunk) {
  sg_pdl_sync();
  const int m0 = blockIdx.y * TC_M, k0o = blockIdx.x * N_TILE;      // output tile: rows m, columns k
  const int nr0 = blockIdx.z * n_chunk, nr1 = min(N, nr0 + n_chunk);
  if (nr0 >= nr1) return;
  const bool atomic_out = gridDim.z > 1;
  KMajorStager<TC_M> sa;
  MNMajorStager<N_TILE> sb;
  sa.init(nr1, [&](int r) -> const float* { return m0 + r < M ? dy + (long long)(m0 + r) * ldy : nullptr; });
  tc_tile<N_TILE, STAGES>(
      nr0, nr1, [&](int n0) { sa.load(n0); },
      [&](int n0) {                                                    // B[k][n] = W[n][k]: rows k contiguous in memory for fixed n
        sb.load(n0, K - k0o, [&](int n) -> const float* { return n < nr1 ? w + (long long)n * ldw + k0o : nullptr; });
      },
      [&](unsigned char* hi, unsigned char* lo) { sa.store(hi, lo); }, [&](unsigned char* hi, unsigned char* lo) { sb.store(hi, lo); },
      [&](int r, int c, float4 v) {
        const int m = m0 + r, k = k0o + c;
        if (m >= M || k >= K) return;
        const float t[4] = {v.x, v.y, v.z, v.w};
        long long row = m;
        if (scatter_ids) {
          row = scatter_ids[m];
          if (row == 0) return;
        }
        float* dst = dx + row * lddx + k;
        if (scatter_ids || atomic_out) {
          if (k + 3 < K && (((size_t)dst) & 15) == 0) atomicAdd(reinterpret_cast<float4*>(dst), v);     // red.global.add.v4.f32
          else for (int j = 0; j < 4 && k + j < K; ++j) atomicAdd(dst + j, t[j]);
          return;
        }
        if (k + 3 < K && (((size_t)dst) & 15) == 0) {
          float4 o = v;
          if (accumulate) { const float4 p = *reinterpret_cast<const float4*>(dst); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
          *reinterpret_cast<float4*>(dst) = o;
        } else {
          for (int j = 0; j < 4 && k + j < K; ++j) dst[j] = accumulate ? dst[j] + t[j] : t[j];
        }
      });
}

// dW[n][k] += sum_m dy[m][n] X[row(m)][k]          (same contract as subgnn_linear_bwd_weight, bias gradient excluded)
template <int N_TILE, int STAGES>
__global__ void __launch_bounds__(TC_THREADS)
tc_linear_bwd_weight_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ x, int ldx, const int* __restrict__ ids,
                            float* __restrict__ dw, int lddw, int M, int N, int K, int m_chunk) {
  sg_pdl_sync();
  const int n0 = blockIdx.y * TC_M, k0o = blockIdx.x * N_TILE;       // output tile: rows n, columns k
  const int mr0 = blockIdx.z * m_chunk, mr1 = min(M, mr0 + m_chunk);
  if (mr0 >= mr1) return;
  MNMajorStager<TC_M> sa;
  MNMajorStager<N_TILE> sb;
  tc_tile<N_TILE, STAGES>(
      mr0, mr1,
      [&](int mb) {                                                    // A[n][m] = dy[m][n]
        sa.load(mb, N - n0, [&](int m) -> const float* { return m < mr1 ? dy + (long long)m * ldy + n0 : nullptr; });
      },
      [&](int mb) {                                                    // B[k][m] = X[row(m)][k]
        sb.load(mb, K - k0o, [&](int m) -> const float* {
          if (m >= mr1) return nullptr;
          const long long row = ids ? (long long)ids[m] : m;
          return x + row * ldx + k0o;
        });
      },
      [&](unsigned char* hi, unsigned char* lo) { sa.store(hi, lo); }, [&](unsigned char* hi, unsigned char* lo) { sb.store(hi, lo); },
      [&](int r, int c, float4 v) {
        const int n = n0 + r, k = k0o + c;
        if (n >= N || k >= K) return;
        float* dst = dw + (long long)n * lddw + k;
        if (k + 3 < K && (((size_t)dst) & 15) == 0) atomicAdd(reinterpret_cast<float4*>(dst), v);
        else {
          const float t[4] = {v.x, v.y, v.z, v.w};
          for (int j = 0; j < 4 && k + j < K; ++j) atomicAdd(dst + j, t[j]);
        }
      });
}

bool tc_ws_usable(const void* a, int lda, const void* b, int ldb);      // tcgemm_ws.cu: TMA-fed warp-specialised kernel
static subgnn_gemm_desc ws_desc(int op, const float* a, int lda, const float* b, int ldb, float* out, int ldo, int M, int N, int K) {
  subgnn_gemm_desc d;
  d.a = a; d.b = b; d.out = out; d.bias = nullptr; d.scatter_ids = nullptr;
  d.op = op; d.lda = lda; d.ldb = ldb; d.ldo = ldo; d.M = M; d.N = N; d.K = K; d.relu = 0; d.accumulate = 0; d.shift = 0; d.period = 0;
  return d;
}

static bool tc_aligned(const void* p, int ld) { return (ld % 4) == 0 && (((size_t)p) & 15) == 0; }
template <int NT, int STAGES> static size_t tc_smem() { return (size_t)STAGES * (2 * TC_M * 128 + 2 * NT * 128) + 1024; }

// tuning aids (tools/gemm_bench.py sweeps them): SUBGNN_TC_STAGES, SUBGNN_TC_NT_{FWD,BWI,BWW}, SUBGNN_TC_SPLITS_{BWI,BWW}
static int tc_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}
// operand buffers: measured on B200 (tools/gemm_sweep.py, profiles/r01_gemm_sweep_ppi_bp.txt) double buffering never pays at the
// LSTM shapes — these GEMMs stream 10-20 MB per launch and are bound by loads in flight per SM, which the extra 64 KB of
// shared memory per CTA reduces (fewer resident CTAs) — so one buffer is the default and two stay available as a switch
static int tc_stages(int n_kb) { (void)n_kb; return tc_env("SUBGNN_TC_STAGES", 1) >= 2 ? 2 : 1; }

#define TC_LAUNCH(KERNEL, NT, STAGES, grid, st, ...)                                                                          \
  do {                                                                                                                        \
    subgnn_note_variant(#KERNEL "<%d,%d>", NT, STAGES);                                                                       \
    cudaFuncSetAttribute(KERNEL<NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem<NT, STAGES>());          \
    sg_launch_pdl<SG_PDL_GEMM>(KERNEL<NT, STAGES>, grid, dim3(TC_THREADS), tc_smem<NT, STAGES>(), st, __VA_ARGS__);             \
  } while (0)
#define TC_DISPATCH(KERNEL, nt, stages, grid, st, ...)                                                                        \
  do {                                                                                                                        \
    if (stages == 2) {                                                                                                        \
      if (nt == 32) TC_LAUNCH(KERNEL, 32, 2, grid, st, __VA_ARGS__);                                                          \
      else if (nt == 64) TC_LAUNCH(KERNEL, 64, 2, grid, st, __VA_ARGS__);                                                     \
      else TC_LAUNCH(KERNEL, 128, 2, grid, st, __VA_ARGS__);                                                                  \
    } else {                                                                                                                  \
      if (nt == 32) TC_LAUNCH(KERNEL, 32, 1, grid, st, __VA_ARGS__);                                                          \
      else if (nt == 64) TC_LAUNCH(KERNEL, 64, 1, grid, st, __VA_ARGS__);                                                     \
      else TC_LAUNCH(KERNEL, 128, 1, grid, st, __VA_ARGS__);                                                                  \
    }                                                                                                                         \
  } while (0)

extern "C" {

int subgnn_tc_linear_fwd(const float* x, int ldx, const int* gather_ids, const float* w, int ldw, const float* bias, float* y, int ldy,
                         int M, int N, int K, int relu, void* stream) {
  SG_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad sizes");
  SG_REQUIRE(tc_aligned(x, ldx) && tc_aligned(w, ldw), "tensor-core path needs 16-byte aligned rows");
  if (M == 0) return SUBGNN_OK;
  if (!gather_ids && tc_ws_usable(x, ldx, w, ldw)) {
    subgnn_gemm_desc d = ws_desc(SUBGNN_GEMM_FWD, x, ldx, w, ldw, y, ldy, M, N, K);
    d.bias = bias; d.relu = relu;
    return subgnn_tc_gemm_group(&d, 1, 0, stream);
  }
  // (stand-alone a 64-wide tile is 1 us faster at M = 10000, N = 256 — tools/gemm_sweep.py — but inside the step graph, next to
  // the row kernels, the 128-wide launch with half the CTAs wins by 4 us per step: tools/ab_bench.sh)
  int nt = tc_env("SUBGNN_TC_NT_FWD", N <= 64 ? 64 : 128);
  if (nt != 32 && nt != 64) nt = 128;
  const int stages = tc_stages(sg_div_up(K, TC_KB));
  dim3 grid(sg_div_up(N, nt), sg_div_up(M, TC_M));
  cudaStream_t st = (cudaStream_t)stream;
  TC_DISPATCH(tc_linear_fwd_kernel, nt, stages, grid, st, x, ldx, gather_ids, w, ldw, bias, y, ldy, M, N, K, relu);
  return subgnn_check_launch("tc_linear_fwd_kernel");
}

int subgnn_tc_linear_bwd_input(const float* dy, int ldy, const float* w, int ldw, float* dx, int lddx, const int* scatter_ids, int M, int N,
                               int K, int accumulate, void* stream) {
  SG_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad sizes");
  SG_REQUIRE(tc_aligned(dy, ldy) && tc_aligned(w, ldw), "tensor-core path needs 16-byte aligned rows");
  if (M == 0) return SUBGNN_OK;
  if (tc_ws_usable(dy, ldy, w, ldw)) {
    subgnn_gemm_desc d = ws_desc(SUBGNN_GEMM_BWD_INPUT, dy, ldy, w, ldw, dx, lddx, M, N, K);
    d.scatter_ids = scatter_ids; d.accumulate = accumulate;
    return subgnn_tc_gemm_group(&d, 1, 0, stream);
  }
  // output tile width: the widest that still gives every SM a CTA.  (Stand-alone, the scatter GEMM of layer 0 is faster with a
  // 64-wide tile and 4 reduction splits, 16.9 against 26.9 us; inside the step graph, where it shares the GPU with three
  // weight-gradient GEMMs, the narrow tile with 2 splits is 4 us per step better: tools/gemm_sweep.py, tools/ab_bench.sh)
  const int sms = subgnn_sm_count(), m_tiles = sg_div_up(M, TC_M);
  int nt = K <= 32 ? 32 : (K <= 64 ? 64 : 128);
  while (nt > 32 && sg_div_up(K, nt) * m_tiles < sms) nt >>= 1;
  nt = tc_env("SUBGNN_TC_NT_BWI", nt);
  if (nt != 32 && nt != 64) nt = 128;
  // reduction splits (atomic output) only where the destination already holds the value to add onto
  int splits = 1;
  if (scatter_ids || accumulate) {
    const int ctas = sg_div_up(K, nt) * m_tiles;
    splits = sg_div_up(2 * sms, ctas);
    const int max_splits = sg_div_up(N, 4 * TC_KB);
    if (splits > max_splits) splits = max_splits;
    splits = tc_env("SUBGNN_TC_SPLITS_BWI", splits);
    if (splits < 1) splits = 1;
  }
  const int n_chunk = sg_div_up(sg_div_up(N, splits), TC_KB) * TC_KB;
  splits = sg_div_up(N, n_chunk);
  const int stages = tc_stages(n_chunk / TC_KB);
  dim3 grid(sg_div_up(K, nt), m_tiles, splits);
  cudaStream_t st = (cudaStream_t)stream;
  TC_DISPATCH(tc_linear_bwd_input_kernel, nt, stages, grid, st, dy, ldy, w, ldw, dx, lddx, scatter_ids, M, N, K, accumulate, n_chunk);
  return subgnn_check_launch("tc_linear_bwd_input_kernel");
}

int subgnn_tc_linear_bwd_weight(const float* dy, int ldy, const float* x, int ldx, const int* gather_ids, float* dw, int lddw, float* db,
                                int M, int N, int K, void* stream) {
  SG_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad sizes");
  SG_REQUIRE(tc_aligned(dy, ldy) && tc_aligned(x, ldx), "tensor-core path needs 16-byte aligned rows");
  if (M == 0) return SUBGNN_OK;
  if (!gather_ids && tc_ws_usable(dy, ldy, x, ldx)) {
    subgnn_gemm_desc d = ws_desc(SUBGNN_GEMM_BWD_WEIGHT, dy, ldy, x, ldx, dw, lddw, M, N, K);
    int rc = subgnn_tc_gemm_group(&d, 1, 0, stream);
    if (!rc && db) rc = subgnn_colsum(dy, ldy, db, M, N, nullptr, stream);
    return rc;
  }
  int nt = tc_env("SUBGNN_TC_NT_BWW", K <= 64 ? 64 : 128);
  if (nt != 32 && nt != 64) nt = 128;
  const int tiles = sg_div_up(K, nt) * sg_div_up(N, TC_M);
  int splits = (3 * subgnn_sm_count() + tiles - 1) / tiles;          // ~3 resident CTAs per SM
  const int max_splits = sg_div_up(M, 4 * TC_KB);                    // (in-graph A/B: 79 splits 0.405 ms/step, 40: 0.412, 20: 0.420)
  if (splits > max_splits) splits = max_splits;
  splits = tc_env("SUBGNN_TC_SPLITS_BWW", splits);
  if (splits < 1) splits = 1;
  const int m_chunk = sg_div_up(sg_div_up(M, splits), TC_KB) * TC_KB;
  splits = sg_div_up(M, m_chunk);
  const int stages = tc_stages(m_chunk / TC_KB);
  dim3 grid(sg_div_up(K, nt), sg_div_up(N, TC_M), splits);
  cudaStream_t st = (cudaStream_t)stream;
  TC_DISPATCH(tc_linear_bwd_weight_kernel, nt, stages, grid, st, dy, ldy, x, ldx, gather_ids, dw, lddw, M, N, K, m_chunk);
  int rc = subgnn_check_launch("tc_linear_bwd_weight_kernel");
  if (rc) return rc;
  if (db) rc = subgnn_colsum(dy, ldy, db, M, N, nullptr, stream);
  return rc;
}

}  // extern "C"
