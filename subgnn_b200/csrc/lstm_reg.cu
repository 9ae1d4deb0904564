// Walk-encoder LSTM recurrence as a register-tiled per-step GEMM with packed fp32 FMAs (FFMA2) and thread-block clusters.
//
// Same contract and data layout as the streaming kernels in lstm.cu (reference: SubGNN.py:60-88 nn.LSTM inside class LSTM,
// called from anchor_patch_samplers.py:413-433); chosen by the host wrappers there whenever lstm_reg_supported(H).
//
// Why this shape (round-1 profile of the streaming kernel, profiles/r01_ncu_lstm_recur_v3.txt): a thread-per-gate-column matvec reads every h value once per
// FMA through a broadcast LDS.128 and is bound by the 128 B/clk shared-memory return path (fma pipe 23 % active).  Here each
// thread owns a 4-sequence x 8-column register tile, so one loaded operand feeds 4-8 FMAs, and the FMAs are issued as packed
// fma.rn.f32x2 (FFMA2, sm_100): 16 issues per k for 32 FMAs.
//
// Forward.  A CTA owns `tile` sequences of one direction and the gate columns of U hidden units (U = H, or H/2 with a
// cluster of two CTAs when H > 64 so that the recurrent weights of a CTA — H x 4U fp32, staged once by a 1-D TMA bulk copy —
// fit in shared memory).  Thread (sgroup, p) owns sequences 4 sgroup .. +3 and the FOUR gates of units 2p, 2p+1: after the
// k loop it holds i, f, g, o of its units in registers, so the cell update needs no exchange and the cell state lives in
// registers for the whole recurrence.  The new h is written (k-major, double buffered) into the h tile of every CTA of the
// cluster through distributed shared memory: ONE (cluster) barrier per time step.
//
// Backward.  Same tiling for d h_{t-1} = d gates . W_hh: thread (sgroup, gate, kgroup) reduces over the U columns of one gate
// for 4 sequences x 8 hidden units; the four per-gate partials (and, in a cluster, the peer CTA's) land in double-buffered
// reduction slots of the CTA that owns the units (DSMEM), where the element-wise BPTT step sums them.
#include "common.cuh"
#include "../../include/subgnn_b200.h"
#include "lstm_reg.cuh"
#include <cstdlib>

namespace {

// inter-layer dropout fused into the recurrence (nn.LSTM dropout=p between layers, training only; same Philox stream as
// dropout_kernel in lstm.cu: element index within the (n_seq*T, 2H) buffer, salt = layer + 64 * step counter)
struct DropArgs {
  float* xd;                  // forward: dropped-out copy of OUT (input of the next layer), or nullptr
  float p;                    // 0: no dropout
  unsigned long long seed;
  unsigned salt;
  const int* step_dev;
};
__device__ __forceinline__ unsigned drop_salt(const DropArgs& d) { return d.step_dev ? d.salt + 64u * (unsigned)*d.step_dev + 0x80000000u : d.salt; }

// sigmoid / tanh from ex2.approx + rcp.approx: relative error ~1e-7 on the sigmoid, absolute error ~2e-7 on tanh — inside the
// fp32 tolerance of the parity tests; 5 instructions instead of ~25 for the libm versions (40 activations per thread and step).
// (rcp.approx.ftz is a bare MUFU.RCP, 1 ulp; __frcp_rn adds a fix-up sequence with a divergent slow path: BSSY / BRA / BSYNC were
// 13 % of the forward kernel's stall samples, profiles/r01_ncu_lstm_recur_v4.txt)
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_fast(1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - 2.f * rcp_fast(1.f + __expf(2.f * x)); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_v4(const float* local_addr, unsigned rank, float4 v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_addr)), "r"(rank));
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(remote), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(const float* local_addr, unsigned rank, float2 v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_addr)), "r"(rank));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" :: "r"(remote), "f"(v.x), "f"(v.y) : "memory");
}
template <int CL>
__device__ __forceinline__ void step_sync() {
  if (CL > 1) cluster_sync_all(); else __syncthreads();
}
// split form: everything written to (distributed) shared memory before step_arrive is visible to the cluster after step_wait.
// The release fence of the arrive waits for the thread's outstanding stores, so fire-and-forget GLOBAL stores of a step are
// issued between the two (profiles/r01_ncu_lstm_h128.txt: MEMBAR + ERRBAR of the arrive were ~8 % of the samples).
template <int CL>
__device__ __forceinline__ void step_arrive() {
  if (CL > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
template <int CL>
__device__ __forceinline__ void step_wait() {
  if (CL > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); else __syncthreads();
}

// 1-D TMA bulk copies global -> shared completing on one mbarrier (SASS UBLKCP); issued by one thread.
__device__ __forceinline__ void bulk_begin(uint64_t* bar, uint32_t total_bytes) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(total_bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(reinterpret_cast<const char*>(dst) + off)), "l"(reinterpret_cast<const char*>(src) + off), "r"(n),
                    "r"(smem_u32(bar)) : "memory");
  }
}
__device__ __forceinline__ void bulk_wait(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLSTMR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t"
      "@P1 bra LSTMR_DONE;\n\tbra LSTMR_WAIT;\n\tLSTMR_DONE:\n\t}\n"
      :: "r"(smem_u32(bar)) : "memory");
}

// acc[s][c] (4 sequences x 4 column pairs) += x[s] * w[c]: 16 FFMA2
__device__ __forceinline__ void tile_fma(float2 (&acc)[4][4], const float4 x, const float4 wa, const float4 wb) {
  const float2 w0 = make_float2(wa.x, wa.y), w1 = make_float2(wa.z, wa.w), w2 = make_float2(wb.x, wb.y), w3 = make_float2(wb.z, wb.w);
  const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const float2 xx = make_float2(xs[s], xs[s]);
    acc[s][0] = __ffma2_rn(xx, w0, acc[s][0]);
    acc[s][1] = __ffma2_rn(xx, w1, acc[s][1]);
    acc[s][2] = __ffma2_rn(xx, w2, acc[s][2]);
    acc[s][3] = __ffma2_rn(xx, w3, acc[s][3]);
  }
}

// ------------------------------------------------------------------------------------------------
// Forward.  grid (n_tiles * CL, 2 directions), cluster (CL, 1, 1); block = KS * (tile/4) * (U/2) threads:
// thread (ks, sgroup, p).  The k range is split over KS warps per register tile (the whole problem is only ~500 such tiles:
// one warp per SM sub-partition cannot hide the FFMA2 / LDS latencies: measured with tools/lstm_bench.py), the partial tiles meet in
// shared memory and each of the KS threads finishes 4/KS of the tile's sequences (activations, cell update, outputs).
// Wp: permuted transposed weights written by lstm_prep_kernel: [dir][rank][k][half][p][4] with the 4 = (gate 2 half + {0,1}) x
// (unit 2p + {0,1}): a thread's two float4 loads per k are contiguous across the warp (conflict free).
template <int CL, int KS, int HC>
__global__ void __launch_bounds__(512)
lstm_fwd_tile_kernel(float* __restrict__ G, const float* __restrict__ Wp, float* __restrict__ OUT, float* __restrict__ CS,
                     int n_seq, int T, int H_arg, int steps_fwd, int steps_rev, int tile, DropArgs drop) {
  const int H = HC > 0 ? HC : H_arg;         // HC: hidden size known at compile time (32 / 64 / 128), 0: generic
  extern __shared__ __align__(16) float sm[];
  __shared__ uint64_t wbar;
  constexpr int NF = 4 / KS;                // sequences finished per thread
  const int U = H / CL, U4 = 4 * U, H4 = 4 * H;
  float* wsm = sm;                          // [H][2][U/2][4]
  float* hS = wsm + (size_t)H * U4;         // [2][tile][H]   h_{t-1} of ALL hidden units, double buffered
  float* part = hS + (size_t)2 * tile * H;  // [KS][tile][4U] partial gate pre-activations
  const int dir = blockIdx.y;
  const unsigned rank = CL > 1 ? cluster_rank() : 0u;
  const int seq0 = (blockIdx.x / CL) * tile;
  const int ns = min(tile, n_seq - seq0);
  const int n_steps = dir == 0 ? steps_fwd : steps_rev;
  const int half_u = U / 2, n_sg = tile / 4;
  const int p = threadIdx.x % half_u, sg = (threadIdx.x / half_u) % n_sg, ks = threadIdx.x / (half_u * n_sg);
  const int s0 = sg * 4;
  const int unit0 = (int)rank * U + 2 * p;              // first of the two hidden units of this thread
  const int kb = ks * (H / KS), ke = kb + H / KS;
  if (threadIdx.x == 0) {
    bulk_begin(&wbar, (uint32_t)((size_t)H * U4 * sizeof(float)));
    bulk_copy(wsm, Wp + ((size_t)dir * CL + rank) * H * U4, (uint32_t)((size_t)H * U4 * sizeof(float)), &wbar);
  }
  for (int e = threadIdx.x; e < 2 * tile * H; e += blockDim.x) hS[e] = 0.f;
  float2 c[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) c[f] = make_float2(0.f, 0.f);
  // Everything above reads only data that is at least two launches old (the permuted weights) or nothing at all, so it runs
  // BEFORE the programmatic-dependency wait, i.e. while the projection GEMM that produces G is still executing (the trigger
  // comes after the wait, so a dependent of this kernel still starts only once G's producer has completed: common.cuh)
  sg_pdl_sync();
  const unsigned dsalt = drop.xd ? drop_salt(drop) : 0u;
  step_sync<CL>();                          // every CTA's h tile is zeroed before a peer writes into it; mbarrier init published
  bulk_wait(&wbar);
  int cur = 0;
  // gate pre-activations of the sequences this thread finishes, one step ahead: the loads of step st+1 are issued after the
  // k loop of step st (its accumulators are dead by then) and land during the activations, the barrier and the next k loop.
  // G is rewritten in place, but only by the thread that read the same addresses.
  float2 gin[NF][4];
  // element offsets of this thread's (sequence, unit pair) at t = 0; a step adds t * 2*H4 (gates) or t * 2*H (cell / output)
  size_t g_off[NF], o_off[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    const int s = min(s0 + ks + f * KS, max(ns - 1, 0));
    g_off[f] = ((size_t)(seq0 + s) * T * 2 + dir) * H4 + unit0;
    o_off[f] = ((size_t)(seq0 + s) * T * 2 + dir) * H + unit0;
  }
  auto load_gin = [&](int st, float2 (&gv)[NF][4]) {
    const int t = dir == 0 ? st : T - 1 - st;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int s = s0 + ks + f * KS;
      const float* gp = G + g_off[f] + (size_t)t * 2 * H4;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        gv[f][g] = s < ns ? *reinterpret_cast<const float2*>(gp + g * H) : make_float2(0.f, 0.f);
    }
  };
  if (n_steps > 0) load_gin(0, gin);
  for (int st = 0; st < n_steps; ++st) {
    const int t = dir == 0 ? st : T - 1 - st;
    float2 acc[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[s][g] = make_float2(0.f, 0.f);
    const float* hk = hS + ((size_t)cur * tile + s0) * H;
    const float* wk = wsm + p * 4;
#pragma unroll 2
    for (int k = kb; k < ke; k += 4) {
      float4 x[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) x[s] = *reinterpret_cast<const float4*>(hk + s * H + k);
      const float* xf = reinterpret_cast<const float*>(x);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 wa = *reinterpret_cast<const float4*>(wk + (size_t)(k + kk) * U4);
        const float4 wb = *reinterpret_cast<const float4*>(wk + (size_t)(k + kk) * U4 + 2 * U);
        tile_fma(acc, make_float4(xf[kk], xf[4 + kk], xf[8 + kk], xf[12 + kk]), wa, wb);
      }
    }
    {
      float* pw = part + ((size_t)ks * tile + s0) * U4 + p * 4;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        *reinterpret_cast<float4*>(pw + (size_t)s * U4) = make_float4(acc[s][0].x, acc[s][0].y, acc[s][1].x, acc[s][1].y);
        *reinterpret_cast<float4*>(pw + (size_t)s * U4 + 2 * U) = make_float4(acc[s][2].x, acc[s][2].y, acc[s][3].x, acc[s][3].y);
      }
    }
    float2 gnx[NF][4];
    if (st + 1 < n_steps) load_gin(st + 1, gnx);
    __syncthreads();
    float* hnxt = hS + (size_t)(cur ^ 1) * tile * H;
    float2 gate_i[NF], gate_f[NF], gate_g[NF], gate_o[NF], h_new[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int s = s0 + ks + f * KS;
      float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;
#pragma unroll
      for (int q = 0; q < KS; ++q) {
        const float4 va = *reinterpret_cast<const float4*>(part + ((size_t)q * tile + s) * U4 + p * 4);
        const float4 vb = *reinterpret_cast<const float4*>(part + ((size_t)q * tile + s) * U4 + 2 * U + p * 4);
        pa.x += va.x; pa.y += va.y; pa.z += va.z; pa.w += va.w;
        pb.x += vb.x; pb.y += vb.y; pb.z += vb.z; pb.w += vb.w;
      }
      // gates (PyTorch order i, f, g, o) of units unit0, unit0 + 1
      const float2 ai = make_float2(sigmoid_fast(pa.x + gin[f][0].x), sigmoid_fast(pa.y + gin[f][0].y));
      const float2 af = make_float2(sigmoid_fast(pa.z + gin[f][1].x), sigmoid_fast(pa.w + gin[f][1].y));
      const float2 ag = make_float2(tanh_fast(pb.x + gin[f][2].x), tanh_fast(pb.y + gin[f][2].y));
      const float2 ao = make_float2(sigmoid_fast(pb.z + gin[f][3].x), sigmoid_fast(pb.w + gin[f][3].y));
      float2 hn = make_float2(0.f, 0.f);
      if (s < ns) {
        c[f].x = fmaf(af.x, c[f].x, ai.x * ag.x);
        c[f].y = fmaf(af.y, c[f].y, ai.y * ag.y);
        hn = make_float2(ao.x * tanh_fast(c[f].x), ao.y * tanh_fast(c[f].y));
      }
      if (CL == 1) *reinterpret_cast<float2*>(hnxt + (size_t)s * H + unit0) = hn;      // rows s >= ns stay zero
      else {
#pragma unroll
        for (unsigned r = 0; r < (unsigned)CL; ++r) st_cluster_v2(hnxt + (size_t)s * H + unit0, r, hn);
      }
      gate_i[f] = ai; gate_f[f] = af; gate_g[f] = ag; gate_o[f] = ao; h_new[f] = hn;
    }
    step_arrive<CL>();
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int s = s0 + ks + f * KS;
      if (s < ns) {
        float* gr = G + g_off[f] + (size_t)t * 2 * H4;
        *reinterpret_cast<float2*>(gr) = gate_i[f];
        *reinterpret_cast<float2*>(gr + H) = gate_f[f];
        *reinterpret_cast<float2*>(gr + 2 * H) = gate_g[f];
        *reinterpret_cast<float2*>(gr + 3 * H) = gate_o[f];
        const size_t e = o_off[f] + (size_t)t * 2 * H;                        // even: units come in pairs
        *reinterpret_cast<float2*>(CS + e) = c[f];
        *reinterpret_cast<float2*>(OUT + e) = h_new[f];
        if (drop.xd) {
          float2 hd = h_new[f];
          if (drop.p > 0.f) {
            const Philox4 r = sg_draw(drop.seed, (uint64_t)(e >> 2), dsalt, SG_TAG_DROP);
            const float keep = 1.f / (1.f - drop.p);
            hd.x *= sg_unit((e & 2) ? r.z : r.x) >= drop.p ? keep : 0.f;
            hd.y *= sg_unit((e & 2) ? r.w : r.y) >= drop.p ? keep : 0.f;
          }
          *reinterpret_cast<float2*>(drop.xd + e) = hd;
        }
      }
    }
    step_wait<CL>();
    cur ^= 1;
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
      for (int g = 0; g < 4; ++g) gin[f][g] = gnx[f][g];
  }
  // steps not taken (top-layer reverse direction with the 'last' aggregator): outputs are never read
}

// ------------------------------------------------------------------------------------------------
// BPTT.  dOUT [n_seq*T][2H]; G holds gate activations on entry and d(pre-activation) on exit.  Whh [2][4H][H] native layout.
// block = KS * (tile/4) * 4 * (H/8) threads: thread (ks, sgroup, gate, kgroup) reduces over a 1/KS slice of the U columns of one
// gate for 4 sequences x 8 hidden units (k in kg*4..+3 and H/2 + kg*4..+3: both weight loads contiguous across the warp).
template <int CL, int KS, int HC>
__global__ void __launch_bounds__(512)
lstm_bwd_tile_kernel(float* __restrict__ G, const float* __restrict__ Whh, const float* __restrict__ OUT, const float* __restrict__ CS,
                     const float* __restrict__ dOUT, const float* __restrict__ dOUT_add, int n_seq, int T, int H_arg, int steps_fwd,
                     int steps_rev, int zero_untaken, float* __restrict__ db_ih, float* __restrict__ db_hh, int tile, DropArgs drop) {
  const int H = HC > 0 ? HC : H_arg;
  extern __shared__ __align__(16) float sm[];
  __shared__ uint64_t wbar;
  const int U = H / CL, U4 = 4 * U, H4 = 4 * H;
  constexpr int SLOTS = 4 * KS * CL;
  float* wsm = sm;                                  // [4 gates][U][H]  rows of W_hh for this CTA's units
  float* dgS = wsm + (size_t)U4 * H;                // [tile][4U]       d(pre-activation) of this CTA's columns
  float* dc_rec = dgS + (size_t)U4 * tile;          // [tile][U]
  float* bias_sm = dc_rec + (size_t)tile * U;       // [4U]
  float* red = bias_sm + U4;                        // [SLOTS][tile][U] partial d h_{t-1} of this CTA's units (read by the element-wise
                                                    // phase, rewritten by the matvec phase: a barrier separates the two)
  const int dir = blockIdx.y;
  const unsigned rank = CL > 1 ? cluster_rank() : 0u;
  const int seq0 = (blockIdx.x / CL) * tile;
  const int ns = min(tile, n_seq - seq0);
  const int n_steps = dir == 0 ? steps_fwd : steps_rev;
  const int tid = threadIdx.x;
  const int KG = H / 8, n_sg = tile / 4;
  const int kg = tid % KG, gate = (tid / KG) % 4, sg = (tid / (4 * KG)) % n_sg, ks = tid / (4 * KG * n_sg);
  const int s0 = sg * 4;
  const int ka = kg * 4, kb2 = H / 2 + kg * 4;      // the two 4-wide groups of hidden units this thread produces
  const int jb = ks * (U / KS), je = jb + U / KS;
  if (tid == 0) {
    bulk_begin(&wbar, (uint32_t)((size_t)U4 * H * sizeof(float)));
    for (int g = 0; g < 4; ++g)
      bulk_copy(wsm + (size_t)g * U * H, Whh + ((size_t)dir * H4 + g * H + rank * U) * H, (uint32_t)((size_t)U * H * sizeof(float)), &wbar);
  }
  for (int e = tid; e < U4 * tile; e += blockDim.x) dgS[e] = 0.f;
  for (int e = tid; e < tile * U; e += blockDim.x) dc_rec[e] = 0.f;
  for (int e = tid; e < U4; e += blockDim.x) bias_sm[e] = 0.f;
  for (int e = tid; e < SLOTS * tile * U; e += blockDim.x) red[e] = 0.f;
  sg_pdl_sync();                            // weight staging (parameters) and the zero fills above overlap the producer of dOUT
  step_sync<CL>();
  bulk_wait(&wbar);
  const int n_el = ns * U;
  const unsigned dsalt = drop.p > 0.f ? drop_salt(drop) : 0u;
  // Global operands of the element-wise BPTT step (gate activations, cell states, output gradient), one step ahead: the loads of
  // step st-1 are issued at the end of the element-wise phase of step st and land during the bias / matvec phases and the two
  // barriers.  G is rewritten in place, but only by the thread that read the same addresses (element -> thread map is fixed).
  constexpr int NIT = (4 + KS * CL - 1) / (KS * CL);     // element-wise iterations per thread: n_el <= tile*U = NIT * 2 * blockDim
  struct Pre { float ig, fg, gg, og, cc, cp, dho; };
  Pre pv[NIT][2];
  // time-invariant part of the element -> address maps (the divisions by U happen once, not every step)
  bool on[NIT][2];
  int dg_off[NIT][2];                      // s * U4 + u: this element's column in dgS
  size_t g_off[NIT][2], o_off[NIT][2];     // element offsets at t = 0 into G, and into CS / OUT / dOUT
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = tid + (it * 2 + q) * (int)blockDim.x;
      on[it][q] = e < n_el;
      const int s = on[it][q] ? e / U : 0, u = on[it][q] ? e % U : 0;
      const int ug = (int)rank * U + u;
      dg_off[it][q] = s * U4 + u;
      g_off[it][q] = ((size_t)(seq0 + s) * T * 2 + dir) * H4 + ug;
      o_off[it][q] = ((size_t)(seq0 + s) * T * 2 + dir) * H + ug;
    }
  }
  auto load_step = [&](int st) {
    const int t = dir == 0 ? st : T - 1 - st;
    const int t_prev = dir == 0 ? t - 1 : t + 1;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float* gp = G + g_off[it][q] + (size_t)t * 2 * H4;
        const size_t oe = o_off[it][q] + (size_t)t * 2 * H;
        Pre& v = pv[it][q];
        v.ig = gp[0]; v.fg = gp[H]; v.gg = gp[2 * H]; v.og = gp[3 * H];
        v.cc = CS[oe];
        v.cp = st > 0 ? CS[o_off[it][q] + (size_t)t_prev * 2 * H] : 0.f;
        // flag 2: the caller's dOUT is non-zero (and written) only at t = T-1 ('last' aggregator, top layer)
        v.dho = (!(zero_untaken & SUBGNN_LSTM_DOUT_LAST_ONLY) || t == T - 1) ? dOUT[oe] : 0.f;
        // second gradient source for the rows t = T-1 (same layout, only those rows written): the layer above's reverse direction,
        // which read this layer's output at its last step only
        if (dOUT_add && t == T - 1) v.dho += dOUT_add[oe];
      }
    }
  };
  if (n_steps > 0) load_step(n_steps - 1);
  for (int st = n_steps - 1; st >= 0; --st) {
    const int t = dir == 0 ? st : T - 1 - st;
    // ---- element-wise BPTT step over this CTA's (sequence, unit) pairs, two per iteration ----
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      if (!on[it][0]) break;
      float dhr[2], dho[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int e = tid + (it * 2 + q) * (int)blockDim.x;
        dho[q] = pv[it][q].dho;
        // dOUT holds the gradient w.r.t. the DROPPED-OUT copy of this layer's output: apply the same mask
        if (drop.p > 0.f) dho[q] *= sg_dropout_scale(drop.seed, dsalt, (uint64_t)(o_off[it][q] + (size_t)t * 2 * H), drop.p);
        float r = 0.f;
        const float* rp = red + (on[it][q] ? e : 0);
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) r += rp[(size_t)sl * tile * U];
        dhr[q] = r;
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (!on[it][q]) continue;
        const int e = tid + (it * 2 + q) * (int)blockDim.x;
        const Pre& v = pv[it][q];
        const float tc = tanh_fast(v.cc);               // the forward kernel's tanh: same value as in h = o * tanh(c)
        const float dh = dho[q] + dhr[q];
        const float dc = dc_rec[e] + dh * v.og * (1.f - tc * tc);
        const float dai = dc * v.gg * v.ig * (1.f - v.ig);
        const float daf = dc * v.cp * v.fg * (1.f - v.fg);
        const float dag = dc * v.ig * (1.f - v.gg * v.gg);
        const float dao = dh * tc * v.og * (1.f - v.og);
        dc_rec[e] = dc * v.fg;
        float* dr = dgS + dg_off[it][q];
        dr[0] = dai; dr[U] = daf; dr[2 * U] = dag; dr[3 * U] = dao;
        float* gp = G + g_off[it][q] + (size_t)t * 2 * H4;
        gp[0] = dai; gp[H] = daf; gp[2 * H] = dag; gp[3 * H] = dao;
      }
    }
    if (st > 0) load_step(st - 1);
    step_sync<CL>();                        // dgS complete; every CTA of the cluster has finished reading its reduction slots
    // ---- bias gradient: column sums over the tile's sequences (rows s >= ns of dgS stay zero) ----
    for (int col = tid; col < U4; col += blockDim.x) {
      float bs = 0.f;
      for (int s = 0; s < ns; ++s) bs += dgS[(size_t)s * U4 + col];
      bias_sm[col] += bs;
    }
    // ---- d h_{t-1} partials ----
    {
      float2 acc[4][4];
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[s][q] = make_float2(0.f, 0.f);
      const float* dg = dgS + (size_t)s0 * U4 + gate * U;
      const float* wr = wsm + (size_t)gate * U * H;
#pragma unroll 2
      for (int jj = jb; jj < je; jj += 4) {
        float4 x[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) x[s] = *reinterpret_cast<const float4*>(dg + (size_t)s * U4 + jj);
        const float* xf = reinterpret_cast<const float*>(x);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wa = *reinterpret_cast<const float4*>(wr + (size_t)(jj + q) * H + ka);
          const float4 wb = *reinterpret_cast<const float4*>(wr + (size_t)(jj + q) * H + kb2);
          tile_fma(acc, make_float4(xf[q], xf[4 + q], xf[8 + q], xf[12 + q]), wa, wb);
        }
      }
      const unsigned owner_a = (unsigned)(ka / U), owner_b = (unsigned)(kb2 / U);
      float* rout = red + (size_t)(((int)rank * 4 + gate) * KS + ks) * tile * U;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s0 + s < ns) {
          const float4 va = make_float4(acc[s][0].x, acc[s][0].y, acc[s][1].x, acc[s][1].y);
          const float4 vb = make_float4(acc[s][2].x, acc[s][2].y, acc[s][3].x, acc[s][3].y);
          float* da = rout + (size_t)(s0 + s) * U + (ka % U);
          float* db = rout + (size_t)(s0 + s) * U + (kb2 % U);
          if (CL == 1) { *reinterpret_cast<float4*>(da) = va; *reinterpret_cast<float4*>(db) = vb; }
          else { st_cluster_v4(da, owner_a, va); st_cluster_v4(db, owner_b, vb); }
        }
      }
    }
    step_sync<CL>();
  }
  for (int lc = tid; lc < U4; lc += blockDim.x) {
    const float v = bias_sm[lc];
    if (v != 0.f) {
      const int col = (lc / U) * H + (int)rank * U + (lc % U);
      if (db_ih) atomicAdd(db_ih + (size_t)dir * H4 + col, v);
      if (db_hh) atomicAdd(db_hh + (size_t)dir * H4 + col, v);
    }
  }
  // steps that were never taken carry no gradient: zero their slots (they still hold pre-activations) unless the caller
  // never reads them (zero_untaken == 0)
  for (int st = n_steps; (zero_untaken & SUBGNN_LSTM_ZERO_UNTAKEN) && st < T; ++st) {
    const int t = dir == 0 ? st : T - 1 - st;
    for (int e = tid; e < ns * U4; e += blockDim.x) {
      const int s = e / U4, lc = e % U4;
      const int col = (lc / U) * H + (int)rank * U + (lc % U);
      G[(((size_t)(seq0 + s) * T + t) * 2 + dir) * H4 + col] = 0.f;
    }
  }
}

// sequences per CTA (multiple of 4, at most 32): fewest waves, then fewest warps sharing an SM sub-partition
int pick_tile(int n_seq, int cl, int threads_per_4seq, size_t smem_fixed, size_t smem_per_seq) {
  const int sms = subgnn_sm_count();
  int best = -1;
  long long best_cost = -1;
  int forced = 0;
  if (const char* e = getenv("SUBGNN_LSTM_TILE")) forced = atoi(e);     // tuning aid (tools/lstm_bench.py)
  for (int tile = 4; tile <= 32; tile += 4) {
    if (forced > 0 && tile != forced) continue;
    const size_t smem = smem_fixed + smem_per_seq * tile;
    if (smem > 220 * 1024) continue;
    const int threads = threads_per_4seq * (tile / 4);
    if (threads > 512) continue;
    int occ = (int)((220 * 1024) / smem);
    const int occ_regs = 65536 / (threads * 128);           // both kernels compile to ~128 registers per thread
    if (occ > occ_regs) occ = occ_regs;
    if (occ < 1) occ = 1;
    if (occ > 4) occ = 4;
    const long long ctas = 2LL * sg_div_up(n_seq, tile) * cl;
    const long long waves = (ctas + (long long)sms * occ - 1) / ((long long)sms * occ);
    // measured (tools/lstm_bench.py, B200): a step costs ~ fixed + c * tile per resident CTA and the CTAs of a wave run side
    // by side, so the fewest waves win and, among those, the smallest tile (most SMs busy): e.g. H = 128, 720 sequences:
    // tile 20 (144 CTAs) 9.0 us/step, 24 (120 CTAs) 9.6, 28 (104) 10.9, 16 (180 CTAs, two waves) 13.8; inside the training
    // step (other graph branches share the GPU) EM-USER shape: 0.715 ms with tile 20 against 0.744 ms with tile 24
    const long long cost = waves * 1000 + tile;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = tile; }
  }
  return best;
}

template <int CL, int KS, int HC>
int launch_fwd(float* G, const float* wp, float* OUT, float* CS, int n_seq, int T, int H, int sf, int sr, DropArgs drop, cudaStream_t st) {
  const int U = H / CL;
  const size_t fixed = (size_t)H * 4 * U * sizeof(float), per_seq = (size_t)(2 * H + KS * 4 * U) * sizeof(float);
  const int tile = pick_tile(n_seq, CL, KS * (U / 2), fixed, per_seq);
  if (tile < 0) { subgnn_set_error("lstm_fwd_tile: no feasible tile"); return SUBGNN_ERR_ARG; }
  const size_t smem = fixed + per_seq * tile;
  cudaFuncSetAttribute(lstm_fwd_tile_kernel<CL, KS, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sg_div_up(n_seq, tile) * CL, 2, 1);
  cfg.blockDim = dim3(KS * (U / 2) * (tile / 4), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;          // PDL: see sg_pdl_sync (common.cuh)
  attr[1].val.programmaticStreamSerializationAllowed = subgnn_pdl_enabled(SG_PDL_RECUR);
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  subgnn_note_variant("lstm_fwd_tile_kernel<%d,%d,%d>", CL, KS, HC);
  cudaLaunchKernelEx(&cfg, lstm_fwd_tile_kernel<CL, KS, HC>, G, wp, OUT, CS, n_seq, T, H, sf, sr, tile, drop);
  return subgnn_check_launch("lstm_fwd_tile_kernel");
}

template <int CL, int KS, int HC>
int launch_bwd(float* G, const float* whh, const float* OUT, const float* CS, const float* dOUT, const float* dOUT_add, int n_seq, int T, int H,
               int sf, int sr, int zero_untaken, float* db_ih, float* db_hh, DropArgs drop, cudaStream_t st) {
  const int U = H / CL;
  const size_t fixed = (size_t)(4 * U * H + 4 * U) * sizeof(float), per_seq = (size_t)(4 * U + U + 4 * KS * CL * U) * sizeof(float);
  const int tile = pick_tile(n_seq, CL, KS * 4 * (H / 8), fixed, per_seq);
  if (tile < 0) { subgnn_set_error("lstm_bwd_tile: no feasible tile"); return SUBGNN_ERR_ARG; }
  const size_t smem = fixed + per_seq * tile;
  cudaFuncSetAttribute(lstm_bwd_tile_kernel<CL, KS, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sg_div_up(n_seq, tile) * CL, 2, 1);
  cfg.blockDim = dim3(KS * (H / 8) * tile, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;          // PDL: see sg_pdl_sync (common.cuh)
  attr[1].val.programmaticStreamSerializationAllowed = subgnn_pdl_enabled(SG_PDL_RECUR);
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  subgnn_note_variant("lstm_bwd_tile_kernel<%d,%d,%d>", CL, KS, HC);
  cudaLaunchKernelEx(&cfg, lstm_bwd_tile_kernel<CL, KS, HC>, G, whh, OUT, CS, dOUT, dOUT_add, n_seq, T, H, sf, sr, zero_untaken, db_ih, db_hh, tile,
                     drop);
  return subgnn_check_launch("lstm_bwd_tile_kernel");
}

}  // namespace

// H % 8: float4 / 8-wide register tiles and 16-byte bulk-copy granules; H > 64 runs on 2-CTA clusters (U = H/2, a multiple of 8)
bool lstm_reg_supported(int H) { return H >= 8 && H <= 128 && (H % 8) == 0 && (H <= 64 || (H % 16) == 0); }
int lstm_reg_cluster(int H) { return H <= 64 ? 1 : 2; }

int lstm_reg_fwd(float* G, const float* wp, float* OUT, float* CS, int n_seq, int T, int H, int sf, int sr, float* xdrop, float p,
                 unsigned long long seed, unsigned salt, const int* step_dev, cudaStream_t st) {
  const DropArgs drop = {xdrop, p, seed, salt, step_dev};
  if (H == 128) return launch_fwd<2, 2, 128>(G, wp, OUT, CS, n_seq, T, H, sf, sr, drop, st);
  if (H == 64) return launch_fwd<1, 4, 64>(G, wp, OUT, CS, n_seq, T, H, sf, sr, drop, st);
  if (H == 32) return launch_fwd<1, 4, 32>(G, wp, OUT, CS, n_seq, T, H, sf, sr, drop, st);
  if (lstm_reg_cluster(H) == 2) return launch_fwd<2, 2, 0>(G, wp, OUT, CS, n_seq, T, H, sf, sr, drop, st);
  if (H % 16 == 0) return launch_fwd<1, 4, 0>(G, wp, OUT, CS, n_seq, T, H, sf, sr, drop, st);
  return launch_fwd<1, 2, 0>(G, wp, OUT, CS, n_seq, T, H, sf, sr, drop, st);
}

int lstm_reg_bwd(float* G, const float* whh, const float* OUT, const float* CS, const float* dOUT, const float* dOUT_add, int n_seq, int T, int H,
                 int sf, int sr, int zero_untaken, float* db_ih, float* db_hh, float p, unsigned long long seed, unsigned salt,
                 const int* step_dev, cudaStream_t st) {
  const DropArgs drop = {nullptr, p, seed, salt, step_dev};
#define SG_BWD_ARGS G, whh, OUT, CS, dOUT, dOUT_add, n_seq, T, H, sf, sr, zero_untaken, db_ih, db_hh, drop, st
  if (H == 128) return launch_bwd<2, 1, 128>(SG_BWD_ARGS);
  if (H == 64) return launch_bwd<1, 4, 64>(SG_BWD_ARGS);
  if (H == 32) return launch_bwd<1, 4, 32>(SG_BWD_ARGS);
  if (lstm_reg_cluster(H) == 2) return launch_bwd<2, 1, 0>(SG_BWD_ARGS);
  if (H % 16 == 0) return launch_bwd<1, 4, 0>(SG_BWD_ARGS);
  return launch_bwd<1, 2, 0>(SG_BWD_ARGS);
#undef SG_BWD_ARGS
}
