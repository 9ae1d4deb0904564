// Walk-encoder LSTM forward recurrence on the tensor cores (tcgen05, 3xTF32) — the A/B of VERDICT r1 item 6 against the
// register-tiled FFMA2 kernel of lstm_reg.cu.  Same contract (reference: SubGNN.py:60-88 nn.LSTM inside class LSTM, called from
// anchor_patch_samplers.py:413-433): G holds the input projections x_t W_ih^T + b on entry and the gate ACTIVATIONS on exit,
// OUT / CS receive h_t / c_t, xdrop (optional) the inter-layer dropout of OUT.  H = 64 only.
//
// Shape of one time step: P[n_seq][4H] = h_{t-1}[n_seq][H] . W_hh^T — 2000 sequences x 256 gate columns x 64.  On tensor cores the
// product itself is nothing (24 MMAs); what bounds a step is everything around it: 10 MUFU operations per cell (3 sigmoids + 2
// tanh = 5 ex2 + 5 rcp), the TMEM read-back, the hi/lo split of the new h into the swizzled operand layout and one cluster-wide
// hand-over.  A 128-sequence tile on ONE SM would spend 128 x 64 x 10 / 16 = 5120 cycles per step in the MUFU pipe alone, so a
// tile is spread over a CLUSTER of 4 CTAs by hidden unit: CTA r owns units 16r .. 16r+15, i.e. 64 gate columns
// [i(16) f(16) g(16) o(16)], and needs ALL 64 units of h_{t-1} as its A operand:
//   A  h_{t-1} tile, 128 rows x 64 k, K-major SWIZZLE_128B (two 32-float atoms), hi and lo copies, DOUBLE buffered — written by
//      the epilogue threads of all 4 CTAs through distributed shared memory (every CTA stores its 16 units into every peer)
//   B  W_hh rows of the CTA's 64 gate columns, 64 x 64, K-major, hi / lo, staged once
//   D  128 lanes x 64 columns of TMEM
//   per step: [warp 8, one thread] 24 x tcgen05.mma (hi.hi + lo.hi + hi.lo) -> commit -> [warps 0-7] tcgen05.ld (8 units x 4 gates
//   per thread), + G, activations, cell update (c in registers), h -> hi / lo -> 16 DSMEM stores, fence.proxy.async,
//   barrier.cluster.arrive — gate / cell / output stores to global — barrier.cluster.wait.
// grid (8 tiles x 4, 2 directions) = 64 CTAs for 1000 sequences.
#include "common.cuh"
#include "../../include/subgnn_b200.h"
#include <cstdlib>

namespace {

constexpr int TC_H = 64, TC_CL = 4, TC_U = 16, TC_NG = 64, TC_M = 128;
constexpr int TC_THREADS = 288;            // 8 epilogue warps (two per TMEM lane quarter) + the MMA warp
constexpr int A_ATOM = TC_M * 128;          // 128 rows x 32 floats
constexpr int A_BUF = 2 * A_ATOM;           // K = 64
constexpr int B_ATOM = TC_NG * 128;
constexpr int B_BUF = 2 * B_ATOM;
constexpr int OFF_AHI = 0, OFF_ALO = 2 * A_BUF, OFF_BHI = 4 * A_BUF, OFF_BLO = 4 * A_BUF + B_BUF;
constexpr int TC_SMEM = 4 * A_BUF + 2 * B_BUF + 1024;

// -DTC_TRACE: clock64 stamps of cluster 0 / direction 0 (tools/lstm_tc_trace.py) — debugging aid, compiled out of the product build
#ifdef TC_TRACE
__device__ long long tc_trace_buf[2 * 16 * 64];
#define TC_TR(role, st, tag) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (st) < 64) tc_trace_buf[((role) * 64 + (st)) * 16 + (tag)] = clock64(); } while (0)
#else
#define TC_TR(role, st, tag) do { } while (0)
#endif

struct TcDrop {
  float* xd;
  float p;
  unsigned long long seed;
  unsigned salt;
  const int* step_dev;
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_fast(1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - 2.f * rcp_fast(1.f + __expf(2.f * x)); }
__device__ __forceinline__ float lo_of(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// byte offset of element (row, k) inside a K-major SWIZZLE_128B operand of `rows` rows: atoms of 32 k, 8-row groups of 1024 B,
// 16-byte chunks permuted by (row & 7)
__device__ __forceinline__ uint32_t sw_off(int row, int k, int atom_bytes) {
  const int kk = k & 31;
  return (uint32_t)((k >> 5) * atom_bytes + (row >> 3) * 1024 + (row & 7) * 128 + ((((kk >> 2) ^ (row & 7))) << 4) + (kk & 3) * 4);
}
__device__ __forceinline__ uint64_t desc_k_major(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
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
__device__ __forceinline__ void st_peer_v4(uint32_t local_saddr, unsigned rank, float4 v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_saddr), "r"(rank));
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(remote), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __cluster_dims__(TC_CL, 1, 1) __launch_bounds__(TC_THREADS, 1)
lstm_fwd_tc_kernel(float* __restrict__ G, const float* __restrict__ Whh, float* __restrict__ OUT, float* __restrict__ CS, int n_seq, int T,
                   int steps_fwd, int steps_rev, TcDrop drop) {
  constexpr int H = TC_H, H4 = 4 * TC_H;
  extern __shared__ unsigned char tc_smem_raw[];
  const uint32_t base = (s32(tc_smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int seq0 = (blockIdx.x / TC_CL) * TC_M;
  const int ns = min(TC_M, n_seq - seq0);
  const int n_steps = dir == 0 ? steps_fwd : steps_rev;

  // ---- prologue (reads parameters only: runs before the programmatic-dependency wait) ----
  if (warp == 8) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&mma_bar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(&tmem_base_s)), "n"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B: rows n = gate * 16 + u  <->  W_hh[dir][gate * H + rank * 16 + u][0 .. 63]; 64 x 64 floats = 1024 float4
  for (int e = tid; e < TC_NG * (H / 4); e += TC_THREADS) {
    const int n = e / (H / 4), k4 = (e % (H / 4)) * 4;
    const int wrow = (n / TC_U) * H + (int)rank * TC_U + (n % TC_U);
    const float4 w = *reinterpret_cast<const float4*>(Whh + ((size_t)dir * H4 + wrow) * H + k4);
    const uint32_t off = sw_off(n, k4, B_ATOM);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(base + OFF_BHI + off), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w) : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(base + OFF_BLO + off), "f"(lo_of(w.x)), "f"(lo_of(w.y)), "f"(lo_of(w.z)),
                 "f"(lo_of(w.w)) : "memory");
  }
  // A buffer 0 (h_{-1} = 0), hi and lo
  for (int e = tid; e < A_BUF / 16; e += TC_THREADS) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" :: "r"(base + OFF_AHI + e * 16), "f"(0.f) : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" :: "r"(base + OFF_ALO + e * 16), "f"(0.f) : "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  sg_pdl_sync();
  // every CTA of the cluster is running and has initialised its buffers before a peer stores into them
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");

  // ---- epilogue-thread state ----
  const int row = (warp & 3) * 32 + lane, half = (warp >> 2) & 1;
  const int unit0 = (int)rank * TC_U + half * 8;               // first of this thread's 8 hidden units
  const bool live = warp < 8 && row < ns;
  const size_t seq = (size_t)(seq0 + min(row, max(ns - 1, 0)));
  float c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = 0.f;
  const unsigned dsalt = drop.xd ? (drop.step_dev ? drop.salt + 64u * (unsigned)*drop.step_dev + 0x80000000u : drop.salt) : 0u;
  const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 8);
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_NG >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  float4 gin[4][2];
  auto load_gin = [&](int st) {
    const int t = dir == 0 ? st : T - 1 - st;
    const float* gp = G + ((seq * T + t) * 2 + dir) * H4 + unit0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      gin[g][0] = live ? *reinterpret_cast<const float4*>(gp + g * H) : make_float4(0.f, 0.f, 0.f, 0.f);
      gin[g][1] = live ? *reinterpret_cast<const float4*>(gp + g * H + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (warp < 8 && n_steps > 0) load_gin(0);

  for (int st = 0; st < n_steps; ++st) {
    const int t = dir == 0 ? st : T - 1 - st;
    const int cur = st & 1;
    if (warp == 8) {
      if (lane == 0) {
        TC_TR(0, st, 0);
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        TC_TR(0, st, 1);
        const uint32_t ahi = base + OFF_AHI + cur * A_BUF, alo = base + OFF_ALO + cur * A_BUF;
        const uint32_t bhi = base + OFF_BHI, blo = base + OFF_BLO;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t ao = a * A_ATOM + ks * 32, bo = a * B_ATOM + ks * 32;
            umma_tf32(tmem_d, desc_k_major(ahi + ao), desc_k_major(bhi + bo), idesc, (a | ks) ? 1u : 0u);
            umma_tf32(tmem_d, desc_k_major(alo + ao), desc_k_major(bhi + bo), idesc, 1u);
            umma_tf32(tmem_d, desc_k_major(ahi + ao), desc_k_major(blo + bo), idesc, 1u);
          }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(s32(&mma_bar)) : "memory");
        TC_TR(0, st, 2);
      }
      __syncwarp();
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    } else {
      float pi[8], pf[8], pg[8], po[8];
      if (tid == 0) TC_TR(1, st, 0);
      mbar_wait(&mma_bar, (uint32_t)(st & 1));
      if (tid == 0) TC_TR(1, st, 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tmem_ld8(taddr + 0 * TC_U, pi);
      tmem_ld8(taddr + 1 * TC_U, pf);
      tmem_ld8(taddr + 2 * TC_U, pg);
      tmem_ld8(taddr + 3 * TC_U, po);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (tid == 0) TC_TR(1, st, 2);
      const float* gi = reinterpret_cast<const float*>(&gin[0][0]);
      float hn[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ai = sigmoid_fast(pi[i] + gi[i]);
        const float af = sigmoid_fast(pf[i] + gi[8 + i]);
        const float ag = tanh_fast(pg[i] + gi[16 + i]);
        const float ao = sigmoid_fast(po[i] + gi[24 + i]);
        pi[i] = ai; pf[i] = af; pg[i] = ag; po[i] = ao;
        hn[i] = 0.f;
        if (live) {
          c[i] = fmaf(af, c[i], ai * ag);
          hn[i] = ao * tanh_fast(c[i]);
        }
      }
      if (tid == 0) TC_TR(1, st, 3);
      // h_t -> the A operand (buffer cur ^ 1) of every CTA of the cluster: rows >= ns stay zero
      {
        const uint32_t nb = (uint32_t)((cur ^ 1) * A_BUF);
        const uint32_t o0 = sw_off(row, unit0, A_ATOM), o1 = sw_off(row, unit0 + 4, A_ATOM);
        const float4 h0 = make_float4(hn[0], hn[1], hn[2], hn[3]), h1 = make_float4(hn[4], hn[5], hn[6], hn[7]);
        const float4 l0 = make_float4(lo_of(hn[0]), lo_of(hn[1]), lo_of(hn[2]), lo_of(hn[3]));
        const float4 l1 = make_float4(lo_of(hn[4]), lo_of(hn[5]), lo_of(hn[6]), lo_of(hn[7]));
#pragma unroll
        for (unsigned r = 0; r < (unsigned)TC_CL; ++r) {
          st_peer_v4(base + OFF_AHI + nb + o0, r, h0);
          st_peer_v4(base + OFF_AHI + nb + o1, r, h1);
          st_peer_v4(base + OFF_ALO + nb + o0, r, l0);
          st_peer_v4(base + OFF_ALO + nb + o1, r, l1);
        }
      }
      if (tid == 0) TC_TR(1, st, 4);
      asm volatile("fence.proxy.async;" ::: "memory");
      if (tid == 0) TC_TR(1, st, 5);
      __syncwarp();
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      if (tid == 0) TC_TR(1, st, 6);
      // fire-and-forget global stores of the step between arrive and wait; then the next step's input projections
      if (live) {
        float* gr = G + ((seq * T + t) * 2 + dir) * H4 + unit0;
        *reinterpret_cast<float4*>(gr) = make_float4(pi[0], pi[1], pi[2], pi[3]);
        *reinterpret_cast<float4*>(gr + 4) = make_float4(pi[4], pi[5], pi[6], pi[7]);
        *reinterpret_cast<float4*>(gr + H) = make_float4(pf[0], pf[1], pf[2], pf[3]);
        *reinterpret_cast<float4*>(gr + H + 4) = make_float4(pf[4], pf[5], pf[6], pf[7]);
        *reinterpret_cast<float4*>(gr + 2 * H) = make_float4(pg[0], pg[1], pg[2], pg[3]);
        *reinterpret_cast<float4*>(gr + 2 * H + 4) = make_float4(pg[4], pg[5], pg[6], pg[7]);
        *reinterpret_cast<float4*>(gr + 3 * H) = make_float4(po[0], po[1], po[2], po[3]);
        *reinterpret_cast<float4*>(gr + 3 * H + 4) = make_float4(po[4], po[5], po[6], po[7]);
        const size_t e = ((seq * T + t) * 2 + dir) * H + unit0;              // multiple of 8
        *reinterpret_cast<float4*>(CS + e) = make_float4(c[0], c[1], c[2], c[3]);
        *reinterpret_cast<float4*>(CS + e + 4) = make_float4(c[4], c[5], c[6], c[7]);
        *reinterpret_cast<float4*>(OUT + e) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        *reinterpret_cast<float4*>(OUT + e + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
        if (drop.xd) {
          float hd[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) hd[i] = hn[i];
          if (drop.p > 0.f) {
            const float keep = 1.f / (1.f - drop.p);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const Philox4 r = sg_draw(drop.seed, (uint64_t)((e >> 2) + q), dsalt, SG_TAG_DROP);
              hd[4 * q + 0] *= sg_unit(r.x) >= drop.p ? keep : 0.f;
              hd[4 * q + 1] *= sg_unit(r.y) >= drop.p ? keep : 0.f;
              hd[4 * q + 2] *= sg_unit(r.z) >= drop.p ? keep : 0.f;
              hd[4 * q + 3] *= sg_unit(r.w) >= drop.p ? keep : 0.f;
            }
          }
          *reinterpret_cast<float4*>(drop.xd + e) = make_float4(hd[0], hd[1], hd[2], hd[3]);
          *reinterpret_cast<float4*>(drop.xd + e + 4) = make_float4(hd[4], hd[5], hd[6], hd[7]);
        }
      }
      if (tid == 0) TC_TR(1, st, 7);
      if (st + 1 < n_steps) load_gin(st + 1);
      if (tid == 0) TC_TR(1, st, 8);
    }
    __syncwarp();
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (tid == 0) TC_TR(1, st, 9);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(64) : "memory");
  // no CTA leaves while a peer may still store into its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace

extern "C" {

#ifdef TC_TRACE
int subgnn_tc_trace_read(long long* out) { return (int)cudaMemcpyFromSymbol(out, tc_trace_buf, sizeof(tc_trace_buf)); }
#endif

int subgnn_lstm_recur_fwd_tc_supported(int H) { return H == TC_H ? 1 : 0; }

int subgnn_lstm_recur_fwd_tc(float* G, const float* whh, float* OUT, float* CS, int n_seq, int T, int H, int steps_fwd, int steps_rev,
                             float* xdrop, float p, unsigned long long seed, unsigned salt, const int* step_dev, void* stream) {
  SG_REQUIRE(H == TC_H, "the tensor-core recurrence is built for H = 64");
  SG_REQUIRE(n_seq >= 0 && T >= 1 && p >= 0.f && p < 1.f, "bad sizes");
  SG_REQUIRE(steps_fwd >= 0 && steps_fwd <= T && steps_rev >= 0 && steps_rev <= T, "bad step counts");
  SG_REQUIRE((((size_t)G | (size_t)whh | (size_t)OUT | (size_t)CS | (size_t)xdrop) & 15) == 0, "buffers must be 16-byte aligned");
  if (n_seq == 0) return SUBGNN_OK;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(lstm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    attr_set = true;
  }
  const TcDrop drop = {xdrop, p, seed, salt, step_dev};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sg_div_up(n_seq, TC_M) * TC_CL, 2, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = TC_SMEM;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = subgnn_pdl_enabled(SG_PDL_RECUR);
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  subgnn_note_variant("lstm_fwd_tc_kernel<4,64>");
  cudaLaunchKernelEx(&cfg, lstm_fwd_tc_kernel, G, whh, OUT, CS, n_seq, T, steps_fwd, steps_rev, drop);
  return subgnn_check_launch("lstm_fwd_tc_kernel");
}

}  // extern "C"
