// Warp-specialised, TMA-fed, persistent tcgen05 GEMM for the walk-encoder LSTM projections and their gradients
// (reference: the cuBLAS sgemms behind nn.LSTM, SubGNN.py:73-78, called from anchor_patch_samplers.py:429, and their autograd).
//
// One kernel, one instantiation, GROUPED: a launch carries up to 6 independent products ("problems"); every CTA (one per SM,
// persistent) walks a static list of work items = (problem, output tile, reduction split).  fp32 parity through 3xTF32
// (D += A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32 accumulation in TMEM; see tcgemm.cu for the error analysis).
//
// Pipeline (per CTA, 10 warps):
//   warp 4      TMA producer   cp.async.bulk.tensor.2d of the RAW fp32 operand tiles into a 3-stage ring (SWIZZLE_128B boxes);
//                               the raw tile IS the hi operand: kind::tf32 reads the top 19 bits of every fp32 word
//   warps 6-9   converters     lo = x - trunc_tf32(x) element-wise from the raw tile into the lo tile of the same stage
//                               (same swizzled offsets: layout-agnostic), fence.proxy.async, arrive
//   warp 5      MMA issuer     one elected thread: 4 K-steps x 3 tcgen05.mma.kind::tf32 per stage, tcgen05.commit frees the stage;
//                               also owns the TMEM allocation (2 accumulators x 128 columns: the next tile's MMAs run while
//                               the previous tile drains)
//   warps 0-3   epilogue       tcgen05.ld 32 lanes x 32 columns -> per-warp smem transposition -> coalesced 16-byte stores /
//                               vector atomics (bias, relu, read-add-store, split-reduction atomics, row scatter into dE)
//
// Operand majors are run-time properties of a problem.  With the reduction index contiguous in memory the operand is K-major
// (one 128-row x 32-float box per stage); with the output index contiguous it is MN-major (one 32 x 32 box per 32 output
// rows): the weight-gradient products dW = dY^T X read dY and X exactly as they lie in memory, no transposition pass, and the
// input-gradient product reads W the same way.  Canonical layouts (cute/atom/mma_traits_sm100.hpp, units of 16 bytes):
//   K-major  SWIZZLE_128B: ((8,n),2):((8,SBO),1)          row = 128 B, 8-row groups SBO = 1024 B apart; K step of 8 = +32 B
//   MN-major SWIZZLE_128B_BASE32B: ((8,n),(4,k)):((1,LBO),(8,SBO)) 32 floats of the output index = 128 B, 4 reduction rows = one
//                          512 B atom (32-byte chunks swizzled by row & 3), atoms SBO = 512 B apart along the reduction, 32-wide
//                          chunks LBO = 4096 B apart; K step of 8 = +1024 B
//
// The recurrent-weight gradient dW_hh = sum_t dG_t^T h_{t-1} pairs row m of dG with row m -/+ 1 of the layer output: the B box
// is fetched one row up / down (b_shift) and the converter warps zero the dG rows of the first / last time step of every
// sequence (zero_mod / zero_rem) — the h_{t-1} = 0 boundary — so no gathered operand is needed on the step path.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include "../../include/subgnn_b200.h"

#define WS_THREADS 320
#define WS_KB 32            // reduction elements per stage (128 bytes)
#define WS_STAGES 3
#define WS_M 128
#define WS_NT_MAX 128
#define WS_MAX_PROBLEMS 6
#define WS_TILE_BYTES (WS_M * 128)                    // one operand tile of a stage: 128 rows x 128 B
#define WS_STAGE_BYTES (4 * WS_TILE_BYTES)            // A raw | A lo | B raw | B lo
#define WS_EPI_LD 36
#define WS_SMEM_BYTES (WS_STAGES * WS_STAGE_BYTES + 4 * 32 * WS_EPI_LD * 4 + 1024)

enum { WS_STORE = 0, WS_ACCUM = 1, WS_ATOMIC = 2, WS_SCATTER = 3 };

struct WsProblem {
  CUtensorMap tmA;
  CUtensorMap tmB;
  float* out;
  const float* bias;
  const int* scatter_ids;
  int ldo, out_rows, out_cols;       // D is out_rows x out_cols: rows = output index of A, cols = output index of B
  int a_mn, b_mn;                    // 0 = K-major (reduction contiguous), 1 = MN-major (output index contiguous)
  int nt;                            // output tile width: 32, 64 or 128
  int tiles_m, tiles_n, splits;
  int red_len, red_chunk;            // reduction length, reduction elements per split (multiple of WS_KB)
  int b_shift, zero_mod, zero_rem;   // recurrent-weight gradient: B rows fetched at (reduction index + b_shift); A rows with
                                     // (reduction index % zero_mod) == zero_rem are taken as zero (zero_mod = 0: off)
  int mode, relu;
  int item_begin, n_items;
};

struct alignas(64) WsParams {
  WsProblem p[WS_MAX_PROBLEMS];
  int n_problems, n_items;
};

// -DWS_TRACE: per-role event clocks of CTA 0 (tools/gemm_trace.py) — debugging aid, compiled out of the product build
#ifdef WS_TRACE
__device__ long long ws_trace_buf[4 * 1024];
__device__ int ws_trace_n[4];
#define WS_TR(role, tag)                                                              \
  do {                                                                                \
    if (blockIdx.x == 0) {                                                            \
      const int i_ = ws_trace_n[role]++;                                              \
      if (i_ < 512) { ws_trace_buf[(role) * 1024 + 2 * i_] = (tag); ws_trace_buf[(role) * 1024 + 2 * i_ + 1] = clock64(); }  \
    }                                                                                 \
  } while (0)
#else
#define WS_TR(role, tag) do { } while (0)
#endif

namespace {

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(bar)), "r"(count) : "memory");
}
// bounded spin: a protocol error traps (a launch failure the host reports) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = s32(bar);
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(s32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptors, SWIZZLE_128B, descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t desc_k_major(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major fp32 / tf32 operands have ONE legal swizzled layout: SWIZZLE_128B_BASE32B (layout type 1) — 32-byte chunks of a 128-byte
// row permuted by (row & 3), atoms of 4 reduction rows (512 B): cutlass sm100_common.inl "for mn-major tf32 operands, SW128_32B is
// the only available smem layout"; the matching TMA mode is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)1 << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32(int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(WS_M >> 4) << 24);                       // D = F32, A = B = TF32, M = 128
}

struct Item {
  int q, tm, tn, r0, n_kb;
};
__device__ __forceinline__ Item decode(const WsParams& P, int item) {
  Item it;
  int q = 0;
#pragma unroll
  for (int i = 1; i < WS_MAX_PROBLEMS; ++i)
    if (i < P.n_problems && item >= P.p[i].item_begin) q = i;
  const WsProblem& pr = P.p[q];
  const int li = item - pr.item_begin;
  it.q = q;
  it.tn = li % pr.tiles_n;
  it.tm = (li / pr.tiles_n) % pr.tiles_m;
  const int sp = li / (pr.tiles_n * pr.tiles_m);
  it.r0 = sp * pr.red_chunk;
  const int r1 = min(pr.red_len, it.r0 + pr.red_chunk);
  it.n_kb = (r1 - it.r0 + WS_KB - 1) / WS_KB;
  return it;
}

// explicit shared-window accesses: the tile pointers are carved out of the dynamic shared buffer by integer arithmetic, so plain
// C++ dereferences compile to GENERIC LD / ST (and every store orders against the following load: same-buffer aliasing)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float4 lo_part(const float4 v) {
  float4 l;
  l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  return l;
}

// column tails / unaligned destinations: element-wise (rare: shapes whose width is not a multiple of 4, odd pitches)
__device__ __noinline__ void ws_epilogue_tail(uint32_t slab_rd, int row0, int out_rows, const int* ids, float* out, const float* bias, int ldo,
                                              int col, int out_cols, int mode, int relu) {
  for (int i8 = 0; i8 < 8; ++i8) {
    const int row = row0 + i8 * 4;
    if (row >= out_rows) continue;
    long long orow = row;
    if (mode == WS_SCATTER) {
      orow = ids[row];
      if (orow == 0) continue;
    }
    const float4 v4 = lds128(slab_rd + i8 * (4 * WS_EPI_LD * 4));
    const float t[4] = {v4.x, v4.y, v4.z, v4.w};
    float* dst = out + orow * ldo + col;
    for (int j = 0; j < 4 && col + j < out_cols; ++j) {
      float v = t[j];
      if (mode == WS_STORE) {
        if (bias) v += bias[col + j];
        if (relu) v = fmaxf(v, 0.f);
        dst[j] = v;
      } else if (mode == WS_ACCUM) dst[j] += v;
      else atomicAdd(dst + j, v);
    }
  }
}

__global__ void __launch_bounds__(WS_THREADS, 1) tc_gemm_ws_kernel(const __grid_constant__ WsParams P) {
  extern __shared__ unsigned char ws_smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)ws_smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_raw[WS_STAGES], full_lo[WS_STAGES], empty[WS_STAGES], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    for (int q = 0; q < P.n_problems; ++q) {
      asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&P.p[q].tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&P.p[q].tmB)) : "memory");
    }
  }
  if (warp == 5) {
    if (lane == 0) {
      for (int s = 0; s < WS_STAGES; ++s) { mbar_init(&full_raw[s], 1); mbar_init(&full_lo[s], 128); mbar_init(&empty[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(&tmem_base_s)), "n"(2 * WS_NT_MAX) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  if (threadIdx.x == 0) WS_TR(3, 9);
  sg_pdl_sync();                                   // everything above overlaps the tail of the preceding kernel
  if (threadIdx.x == 0) WS_TR(3, 8);

  // Every role copies the fields of its problem into registers ONCE per work item: the problem table lives in the kernel parameter
  // bank and is indexed by a run-time problem id, so every `P.p[q].x` is a constant-bank load with a register offset — re-issued
  // after each asm volatile (the first build spent ~1400 instructions per 32-column epilogue block that way).
  if (warp == 4) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const Item it = decode(P, item);
        const WsProblem& pr = P.p[it.q];
        const CUtensorMap* tmA = &pr.tmA;
        const CUtensorMap* tmB = &pr.tmB;
        const int a_mn = pr.a_mn, b_mn = pr.b_mn, nt = pr.nt, b_shift = pr.b_shift, n_kb = it.n_kb, r0 = it.r0;
        const int a_row0 = it.tm * WS_M, b_row0 = it.tn * nt;
        // only boxes that intersect the tensor are fetched (a fully out-of-range 32-wide chunk of an MN-major operand feeds
        // output rows / columns beyond the matrix, which the epilogue never stores)
        int a_chunks = 1, b_chunks = 1;
        if (a_mn) a_chunks = min(WS_M / 32, (pr.out_rows - a_row0 + 31) / 32);
        if (b_mn) b_chunks = min(nt / 32, (pr.out_cols - b_row0 + 31) / 32);
        const uint32_t bytes = (a_mn ? a_chunks * 4096u : (uint32_t)WS_TILE_BYTES) + (b_mn ? b_chunks * 4096u : (uint32_t)nt * 128u);
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          WS_TR(0, 1);
          unsigned char* a_raw = base + stage * WS_STAGE_BYTES;
          unsigned char* b_raw = a_raw + 2 * WS_TILE_BYTES;
          mbar_arrive_expect_tx(&full_raw[stage], bytes);
          const int red = r0 + kb * WS_KB;
          if (!a_mn) tma_load_2d(tmA, &full_raw[stage], a_raw, red, a_row0);
          else {
#pragma unroll 1
            for (int c = 0; c < a_chunks; ++c) tma_load_2d(tmA, &full_raw[stage], a_raw + c * 4096, a_row0 + 32 * c, red);
          }
          if (!b_mn) tma_load_2d(tmB, &full_raw[stage], b_raw, red, b_row0);
          else {
#pragma unroll 1
            for (int c = 0; c < b_chunks; ++c) tma_load_2d(tmB, &full_raw[stage], b_raw + c * 4096, b_row0 + 32 * c, red + b_shift);
          }
          if (++stage == WS_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 5) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t smem0 = s32(base);
      for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const Item it = decode(P, item);
        const WsProblem& pr = P.p[it.q];
        const int a_mn = pr.a_mn, b_mn = pr.b_mn, n_kb = it.n_kb;
        const uint32_t idesc = idesc_tf32(pr.nt, a_mn, b_mn);
        // descriptor = constant part | start address; the address field advances by whole 16-byte units
        const uint64_t a_const = (a_mn ? desc_mn_major(0) : desc_k_major(0)), b_const = (b_mn ? desc_mn_major(0) : desc_k_major(0));
        const uint32_t a_step = a_mn ? 64u : 2u, b_step = b_mn ? 64u : 2u;     // K = 8 per MMA: +1024 B (MN-major) or +32 B (K-major)
        const uint32_t d_addr = tmem_d + (uint32_t)(acc * WS_NT_MAX);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);                  // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&full_lo[stage], phase);
          WS_TR(1, 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a16 = (smem0 + stage * WS_STAGE_BYTES) >> 4;           // tiles are 1024-byte aligned, < 256 KB: fits the 14-bit field
          const uint64_t dah = a_const | a16, dal = a_const | (a16 + (WS_TILE_BYTES >> 4));
          const uint64_t dbh = b_const | (a16 + 2 * (WS_TILE_BYTES >> 4)), dbl = b_const | (a16 + 3 * (WS_TILE_BYTES >> 4));
#pragma unroll
          for (int ks = 0; ks < WS_KB / 8; ++ks) {
            umma_tf32(d_addr, dah + a_step * ks, dbh + b_step * ks, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
            umma_tf32(d_addr, dal + a_step * ks, dbh + b_step * ks, idesc, 1u);
            umma_tf32(d_addr, dah + a_step * ks, dbl + b_step * ks, idesc, 1u);
          }
          umma_commit(&empty[stage]);                                  // arrives when the MMAs above have read the stage
          WS_TR(1, 2);
          if (++stage == WS_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full[acc]);                                  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 6) {
    // ===================================== converters =====================================
    const int tc = threadIdx.x - 6 * 32;                               // 0 .. 127
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t smem0 = s32(base) + tc * 16;
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = decode(P, item);
      const WsProblem& pr = P.p[it.q];
      const int b_vec = pr.nt * 8;                                     // 16-byte words of the B tile
      const int zero_mod = pr.zero_mod, zero_rem = pr.zero_rem, n_kb = it.n_kb, r0 = it.r0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&full_raw[stage], phase);
        if (tc == 0) WS_TR(2, 1);
        const uint32_t a_raw = smem0 + stage * WS_STAGE_BYTES, a_lo = a_raw + WS_TILE_BYTES;
        const uint32_t b_raw = a_raw + 2 * WS_TILE_BYTES, b_lo = b_raw + WS_TILE_BYTES;
        // thread tc owns the 16-byte words tc, tc + 128, ... of either tile; all loads in flight first, then the stores
        float4 va[8], vb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) va[i] = lds128(a_raw + i * 2048);
#pragma unroll
        for (int i = 0; i < 8; ++i) vb[i] = (tc + i * 128 < b_vec) ? lds128(b_raw + i * 2048) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (zero_mod > 0) {
          // MN-major A tile: 16-byte word w lives in chunk w / 256, reduction row (w % 256) / 8
          const int red = r0 + kb * WS_KB;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int w = tc + i * 128;
            const int m = red + ((w & 255) >> 3);
            if (m % zero_mod == zero_rem) {
              va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              sts128(a_raw + i * 2048, va[i]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) sts128(a_lo + i * 2048, lo_part(va[i]));
#pragma unroll
        for (int i = 0; i < 8; ++i) if (tc + i * 128 < b_vec) sts128(b_lo + i * 2048, lo_part(vb[i]));
        if (tc == 0) WS_TR(2, 2);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        mbar_arrive(&full_lo[stage]);
        if (tc == 0) WS_TR(2, 3);
        if (++stage == WS_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ===================================== epilogue (warps 0-3 <-> TMEM lane quarters) =====================================
    const uint32_t slab = s32(base + WS_STAGES * WS_STAGE_BYTES) + warp * (32 * WS_EPI_LD * 4);
    const uint32_t slab_wr = slab + lane * (WS_EPI_LD * 4);                      // my row of the transposition slab
    const int rsub = lane >> 3, cc = (lane & 7) * 4;
    const uint32_t slab_rd = slab + (rsub * WS_EPI_LD + cc) * 4;                 // + i8 * 4 rows
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = decode(P, item);
      const WsProblem& pr = P.p[it.q];
      float* const out = pr.out;
      const float* const bias = pr.bias;
      const int* const ids = pr.scatter_ids;
      const int ldo = pr.ldo, out_rows = pr.out_rows, out_cols = pr.out_cols, nt = pr.nt, mode = pr.mode, relu = pr.relu;
      const int row0 = it.tm * WS_M + warp * 32 + rsub, col_base = it.tn * nt + cc;
      const bool aligned = ((((size_t)out) & 15) == 0) && ((ldo & 3) == 0) && (!bias || ((((size_t)bias) & 15) == 0));
      // bias of the whole tile (<= 4 column blocks), fetched BEFORE the wait for the accumulator: a per-block fetch left its L2
      // latency in front of every block's stores (ncu source page, round 2: 11 % of the kernel's stall samples on the first bias add)
      float4 bq_all[WS_NT_MAX / 32];
#pragma unroll
      for (int j = 0; j < WS_NT_MAX / 32; ++j) {
        const int col = col_base + 32 * j;
        bq_all[j] = (mode == WS_STORE && bias && aligned && 32 * j < nt && col + 3 < out_cols) ? __ldg(reinterpret_cast<const float4*>(bias + col))
                                                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      if (threadIdx.x == 0) WS_TR(3, 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c0 = 0; c0 < nt; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * WS_NT_MAX + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        const int col = col_base + c0;
        const bool fast = aligned && (col + 3 < out_cols);
        float4 bq = bq_all[0];                                         // this block's bias (register select: the loop is not unrolled)
#pragma unroll
        for (int j = 1; j < WS_NT_MAX / 32; ++j)
          if (c0 == 32 * j) bq = bq_all[j];
        int orow[8];
        if (mode == WS_SCATTER) {                                      // row targets: 8 independent loads, PAD row (id 0) skipped
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int row = row0 + i8 * 4;
            const int o = (row < out_rows) ? __ldg(ids + row) : 0;
            orow[i8] = o == 0 ? -1 : o;
          }
        } else {
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) { const int row = row0 + i8 * 4; orow[i8] = (row < out_rows) ? row : -1; }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 32 >= nt) {                                           // last column block read: the accumulator may be overwritten
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&tmem_empty[acc]);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(slab_wr + j * 4, make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
        __syncwarp();
        if (fast) {
          float* const dst0 = out + col;
          if (mode == WS_STORE) {
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
              float4 v = lds128(slab_rd + i8 * (4 * WS_EPI_LD * 4));
              v.x += bq.x; v.y += bq.y; v.z += bq.z; v.w += bq.w;
              if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              if (orow[i8] >= 0) *reinterpret_cast<float4*>(dst0 + (size_t)orow[i8] * ldo) = v;
            }
          } else if (mode == WS_ACCUM) {
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
              const float4 v = lds128(slab_rd + i8 * (4 * WS_EPI_LD * 4));
              if (orow[i8] >= 0) {
                float4* d4 = reinterpret_cast<float4*>(dst0 + (size_t)orow[i8] * ldo);
                float4 o = *d4;
                o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
                *d4 = o;
              }
            }
          } else {
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
              const float4 v = lds128(slab_rd + i8 * (4 * WS_EPI_LD * 4));
              if (orow[i8] >= 0) atomicAdd(reinterpret_cast<float4*>(dst0 + (size_t)orow[i8] * ldo), v);   // red.global.add.v4.f32
            }
          }
        } else if (col < out_cols) {
          ws_epilogue_tail(slab_rd, row0, out_rows, ids, out, bias, ldo, col, out_cols, mode, relu);
        }
        __syncwarp();
        if (threadIdx.x == 0) WS_TR(3, 2);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(2 * WS_NT_MAX) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp32 matrix with `inner` contiguous elements per row, `outer` rows `ld` floats apart; box = 32 x box_rows; K-major operands use
// SWIZZLE_128B (16-byte chunks), MN-major operands SWIZZLE_128B_ATOM_32B (32-byte chunks)
int make_map(CUtensorMap* tm, const float* ptr, long long inner, long long outer, long long ld, int box_rows, bool mn_major = false) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) { subgnn_set_error("cuTensorMapEncodeTiled is not available from this driver"); return SUBGNN_ERR_CUDA; }
  if ((((size_t)ptr) & 15) != 0 || (ld % 4) != 0 || inner < 1 || outer < 1) {
    subgnn_set_error("TMA operand needs a 16-byte aligned base and row pitch (ptr %p, ld %lld)", (const void*)ptr, ld);
    return SUBGNN_ERR_ARG;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)WS_KB, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { subgnn_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return SUBGNN_ERR_CUDA; }
  return SUBGNN_OK;
}

int pick_nt(int cols) { return cols > 64 ? 128 : (cols > 32 ? 64 : 32); }

}  // namespace

bool tc_ws_usable(const void* a, int lda, const void* b, int ldb) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("SUBGNN_TC_LEGACY"); off = (e && *e && atoi(e) != 0) ? 1 : 0; }
  if (const char* f = getenv("SUBGNN_TC_LEGACY_FORCE")) { if (*f) return atoi(f) == 0 && encode_fn() != nullptr; }   // tools/gemm_bench.py: per-call A/B
  return !off && (lda % 4) == 0 && (ldb % 4) == 0 && (((size_t)a | (size_t)b) & 15) == 0 && encode_fn() != nullptr;
}

#ifdef WS_TRACE
extern "C" int subgnn_ws_trace_read(long long* host_buf, int* host_n) {       // host_buf[4096], host_n[4]; also resets the counters
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host_buf, ws_trace_buf, sizeof(long long) * 4096);
  cudaMemcpyFromSymbol(host_n, ws_trace_n, sizeof(int) * 4);
  const int z[4] = {0, 0, 0, 0};
  cudaMemcpyToSymbol(ws_trace_n, z, sizeof(z));
  return 0;
}
#endif

extern "C" {

int subgnn_gemm_desc_size(void) { return (int)sizeof(subgnn_gemm_desc); }
int subgnn_tc_ws_available(void) { return encode_fn() != nullptr ? 1 : 0; }

int subgnn_tc_gemm_group(const subgnn_gemm_desc* g, int n, int max_ctas, void* stream) {
  SG_REQUIRE(g && n >= 1 && n <= WS_MAX_PROBLEMS, "1 .. 6 problems per group");
  static WsParams P;                                                   // (host staging; the launch copies it by value)
  memset(&P, 0, sizeof(P));
  const int sms = subgnn_sm_count();
  int items = 0, n_used = 0;
  // target number of work items of the split reductions: every SM gets about two, each at least 4 stages long
  int n_split_problems = 0;
  for (int i = 0; i < n; ++i) n_split_problems += (g[i].op >= SUBGNN_GEMM_BWD_WEIGHT || (g[i].op == SUBGNN_GEMM_BWD_INPUT && (g[i].scatter_ids || g[i].accumulate))) ? 1 : 0;
  for (int i = 0; i < n; ++i) {
    const subgnn_gemm_desc& d = g[i];
    SG_REQUIRE(d.M >= 0 && d.N >= 1 && d.K >= 1, "bad sizes");
    if (d.M == 0) continue;
    WsProblem& pr = P.p[n_used];
    int rc = SUBGNN_OK;
    pr.bias = nullptr; pr.scatter_ids = nullptr; pr.relu = 0; pr.b_shift = 0; pr.zero_mod = 0; pr.zero_rem = 0;
    bool splittable = false;
    if (d.op == SUBGNN_GEMM_FWD) {                 // out[M][N] = a[M][K] . b[N][K]^T + bias
      pr.a_mn = 0; pr.b_mn = 0; pr.out_rows = d.M; pr.out_cols = d.N; pr.red_len = d.K; pr.nt = pick_nt(d.N);
      rc = make_map(&pr.tmA, d.a, d.K, d.M, d.lda, WS_M);
      if (!rc) rc = make_map(&pr.tmB, d.b, d.K, d.N, d.ldb, pr.nt);
      pr.mode = WS_STORE; pr.bias = d.bias; pr.relu = d.relu;
    } else if (d.op == SUBGNN_GEMM_BWD_INPUT) {    // out[row(m)][K] (+)= a[M][N] . b[N][K]
      pr.a_mn = 0; pr.b_mn = 1; pr.out_rows = d.M; pr.out_cols = d.K; pr.red_len = d.N; pr.nt = pick_nt(d.K);
      rc = make_map(&pr.tmA, d.a, d.N, d.M, d.lda, WS_M);
      if (!rc) rc = make_map(&pr.tmB, d.b, d.K, d.N, d.ldb, 32, true);
      pr.scatter_ids = d.scatter_ids;
      // accumulate: 1 = read-add-store (one writer per element; atomics when the reduction is split), 2 = atomics from the start
      // (another product of the SAME group adds into the same rows, e.g. the reverse direction's last-step rows of a layer)
      pr.mode = d.scatter_ids ? WS_SCATTER : (d.accumulate == 2 ? WS_ATOMIC : d.accumulate ? WS_ACCUM : WS_STORE);
      splittable = d.scatter_ids != nullptr || d.accumulate == 1;
    } else {                                       // out[N][K] += a[M][N]^T . b[row + shift][K]
      SG_REQUIRE(d.op == SUBGNN_GEMM_BWD_WEIGHT || d.op == SUBGNN_GEMM_BWD_WEIGHT_SHIFT, "unknown op");
      pr.a_mn = 1; pr.b_mn = 1; pr.out_rows = d.N; pr.out_cols = d.K; pr.red_len = d.M; pr.nt = pick_nt(d.K);
      rc = make_map(&pr.tmA, d.a, d.N, d.M, d.lda, 32, true);
      if (!rc) rc = make_map(&pr.tmB, d.b, d.K, d.M, d.ldb, 32, true);
      pr.mode = WS_ATOMIC;
      splittable = true;
      if (d.op == SUBGNN_GEMM_BWD_WEIGHT_SHIFT) {
        SG_REQUIRE((d.shift == 1 || d.shift == -1) && d.period >= 1, "shift must be +-1 with a positive period");
        pr.b_shift = d.shift; pr.zero_mod = d.period; pr.zero_rem = d.shift < 0 ? 0 : d.period - 1;
      }
    }
    if (rc) return rc;
    pr.out = d.out; pr.ldo = d.ldo;
    pr.tiles_m = sg_div_up(pr.out_rows, WS_M);
    pr.tiles_n = sg_div_up(pr.out_cols, pr.nt);
    int splits = 1;
    if (splittable) {
      const int tiles = pr.tiles_m * pr.tiles_n;
      const int want = (2 * sms) / (n_split_problems > 0 ? n_split_problems : 1);
      splits = sg_div_up(want, tiles);
      const int max_splits = sg_div_up(pr.red_len, 4 * WS_KB);
      if (splits > max_splits) splits = max_splits;
      if (const char* e = getenv("SUBGNN_TC_WS_SPLITS")) { if (*e) splits = atoi(e); }
      if (splits < 1) splits = 1;
      if (splits > 1 && pr.mode == WS_ACCUM) pr.mode = WS_ATOMIC;      // the destination holds the value to add onto
    }
    pr.red_chunk = sg_div_up(sg_div_up(pr.red_len, splits), WS_KB) * WS_KB;
    pr.splits = sg_div_up(pr.red_len, pr.red_chunk);
    pr.item_begin = items;
    pr.n_items = pr.tiles_m * pr.tiles_n * pr.splits;
    items += pr.n_items;
    ++n_used;
  }
  if (n_used == 0) return SUBGNN_OK;
  P.n_problems = n_used;
  P.n_items = items;
  int grid = items < sms ? items : sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tc_gemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES);
    attr_set = true;
  }
  subgnn_note_variant("tc_gemm_ws_kernel<%d>", n_used);
  sg_launch_pdl<SG_PDL_GEMM>(tc_gemm_ws_kernel, dim3(grid), dim3(WS_THREADS), (size_t)WS_SMEM_BYTES, (cudaStream_t)stream, P);
  return subgnn_check_launch("tc_gemm_ws_kernel");
}

// out[m][:] = table[ids[m]][:]  (anchor_patch_samplers.py:409 embedding lookup of the walk nodes, once per step; the projection
// GEMMs then read a dense matrix through TMA)
__global__ void gather_rows_kernel(const float* __restrict__ table, const int* __restrict__ ids, float* __restrict__ out, int M, int D4) {
  sg_pdl_sync();
  const long long total = (long long)M * D4;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(e / D4), c = (int)(e % D4);
    reinterpret_cast<float4*>(out)[e] = __ldg(reinterpret_cast<const float4*>(table) + (long long)ids[m] * D4 + c);
  }
}

int subgnn_gather_rows(const float* table, const int* ids, float* out, int M, int D, void* stream) {
  SG_REQUIRE(M >= 0 && D >= 4 && (D % 4) == 0, "row length must be a multiple of 4 floats");
  SG_REQUIRE((((size_t)table | (size_t)out) & 15) == 0, "buffers must be 16-byte aligned");
  if (M == 0) return SUBGNN_OK;
  sg_launch_pdl<SG_PDL_CHAIN>(gather_rows_kernel, dim3(sg_grid_for((long long)M * (D / 4), 256, 8)), dim3(256), 0, (cudaStream_t)stream, table, ids, out, M, D / 4);
  return subgnn_check_launch("gather_rows_kernel");
}

}  // extern "C"
