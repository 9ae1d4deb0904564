// Data-parallel gradient exchange fused with the optimizer, over NVLink peer memory (no NCCL on the step path).
//
// Reference semantics: every process steps its shard of the global batch and the gradients are averaged before
// torch.optim.Adam / clip_grad_norm_ (SubGNN.py:1156-1164 under Lightning's distributed back-ends).  Round 1 did
// all-reduce (NCCL) -> sum of squares -> Adam: the 4 flat arenas streamed three times and a library collective between two
// graph halves.  Here the gradient arenas of all ranks are SYMMETRIC allocations (torch.distributed._symmetric_memory: every rank
// maps every peer's arena), and the exchange is two kernels around cross-GPU barriers:
//
//   dp_reduce_scatter_kernel   rank r sums shard r of the gradient arena over all peers (16-byte loads straight from the peers'
//                              HBM through NVLink / NVSwitch, fixed summation order), keeps the sum in a local shard buffer and
//                              publishes the shard's sum of squares in slot r of every peer
//   dp_adam_allgather_kernel   global norm = fixed-order sum of the slots -> clip coefficient; Adam on shard r (the only copy of
//                              m / v for that shard lives on rank r); the updated parameters are written into shard r of EVERY
//                              rank's parameter arena (16-byte stores through NVLink)
//
// With NVLS multicast mappings (hdl.multicast_ptr) the shard sum is ONE multimem.ld_reduce per 16 bytes — the NVSwitch adds the ranks'
// values — and the parameter broadcast ONE multimem.st: fabric traffic per rank drops from (world-1)/world of an arena each way to
// 1/world.  Without multicast support the kernels fall back to per-peer loads / stores (same results up to summation order).
//
// Per step a rank reads (world-1)/world of one arena and writes (world-1)/world of one arena over the fabric — a reduce-scatter
// plus an all-gather, the bandwidth-optimal decomposition of an all-reduce — and the optimizer state is sharded (ZeRO-1 style).
// Every rank ends a step with bit-identical parameters: each shard has a single writer.
#include "common.cuh"
#include "../../include/subgnn_b200.h"

#define DP_MAX_WORLD 16
#define DP_MAX_BLOCKS 1024

struct DpPtrs {
  float* p[DP_MAX_WORLD];
};

__device__ float dp_part[DP_MAX_BLOCKS];
__device__ unsigned dp_ticket = 0;

// peer memory is read with ld.volatile (never served from a stale cache line of the previous step) and written with plain stores
// followed by a system-scope fence; the cross-GPU barriers around the kernels carry the release / acquire
__device__ __forceinline__ float4 ld_peer(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// NVLS: one load from the MULTICAST mapping of the gradient arenas returns the sum over all ranks, reduced inside the NVSwitch
// (multimem.ld_reduce), one store to the multicast mapping of the parameter arenas lands in every rank's copy (multimem.st)
__device__ __forceinline__ float4 ld_reduce_mc(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_mc(float* p, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256) dp_reduce_scatter_kernel(DpPtrs grads, DpPtrs slots, const float* __restrict__ mc_grads, int world, int rank,
                                                               long long n, long long shard, float* __restrict__ gsum) {
  sg_pdl_sync();
  const long long base = (long long)rank * shard;
  const long long len = max(0LL, min(shard, n - base));            // elements of my shard (multiple of 4 except at the arena's end)
  const long long len4 = len / 4;
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mc_grads) acc = ld_reduce_mc(mc_grads + base + 4 * i);      // in-switch reduction: one load instead of `world`
    else {
#pragma unroll 8
      for (int q = 0; q < world; ++q) {                             // fixed order: identical sums whatever the launch geometry
        const float4 v = ld_peer(grads.p[q] + base + 4 * i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    reinterpret_cast<float4*>(gsum)[i] = acc;
    s = fmaf(acc.x, acc.x, s); s = fmaf(acc.y, acc.y, s); s = fmaf(acc.z, acc.z, s); s = fmaf(acc.w, acc.w, s);
  }
  if (blockIdx.x == 0)
    for (long long i = len4 * 4 + threadIdx.x; i < len; i += blockDim.x) {
      float a = 0.f;
      for (int q = 0; q < world; ++q) a += *reinterpret_cast<volatile const float*>(grads.p[q] + base + i);
      gsum[i] = a;
      s = fmaf(a, a, s);
    }
  // deterministic block reduction: partials, last block adds them in a fixed order and publishes the shard's sum of squares
  __shared__ float ws[8];
  __shared__ bool last;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += ws[w];
    dp_part[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(&dp_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += 256) t += __ldcg(dp_part + i);
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += ws[w];
    for (int q = 0; q < world; ++q) slots.p[q][rank] = tot;         // slot `rank` of every peer
    __threadfence_system();
    dp_ticket = 0;
  }
}

__global__ void __launch_bounds__(256)
dp_adam_allgather_kernel(DpPtrs params, float* __restrict__ mc_params, int world, int rank, long long n, long long shard, const float* __restrict__ gsum, float* __restrict__ m,
                         float* __restrict__ v, float lr, float beta1, float beta2, float eps, const int* __restrict__ step_dev,
                         const float* slots_local, float clip_norm, float grad_scale) {
  sg_pdl_sync();
  const long long base = (long long)rank * shard;
  const long long len = max(0LL, min(shard, n - base));
  const int t = *step_dev;
  float sumsq = 0.f;
  for (int q = 0; q < world; ++q) sumsq += *reinterpret_cast<volatile const float*>(slots_local + q);     // fixed order on every rank
  float coef = grad_scale;
  if (clip_norm > 0.f) {
    const float total = sqrtf(sumsq) * grad_scale;                   // norm of the averaged gradient (clip_grad_norm_)
    const float c = clip_norm / (total + 1e-6f);
    if (c < 1.f) coef *= c;
  }
  const float bc1 = 1.f - powf(beta1, (float)t), bc2 = 1.f - powf(beta2, (float)t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  float* mine = params.p[rank] + base;
  float* ms = m + base;
  float* vs = v + base;
  auto upd = [&](float gi_raw, float& pi, float& mi_, float& vi_) {
    const float gi = gi_raw * coef;
    const float mi = beta1 * mi_ + (1.f - beta1) * gi;
    const float vi = beta2 * vi_ + (1.f - beta2) * gi * gi;
    mi_ = mi;
    vi_ = vi;
    pi -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
  };
  const long long len4 = len / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(gsum)[i];
    float4 p4 = reinterpret_cast<float4*>(mine)[i], m4 = reinterpret_cast<float4*>(ms)[i], v4 = reinterpret_cast<float4*>(vs)[i];
    upd(g4.x, p4.x, m4.x, v4.x); upd(g4.y, p4.y, m4.y, v4.y); upd(g4.z, p4.z, m4.z, v4.z); upd(g4.w, p4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(ms)[i] = m4;
    reinterpret_cast<float4*>(vs)[i] = v4;
    if (mc_params) st_mc(mc_params + base + 4 * i, p4);                                         // one store, every rank's arena
    else {
#pragma unroll 8
      for (int q = 0; q < world; ++q) reinterpret_cast<float4*>(params.p[q] + base)[i] = p4;    // my shard of every rank's arena
    }
  }
  if (blockIdx.x == 0)
    for (long long i = len4 * 4 + threadIdx.x; i < len; i += blockDim.x) {
      float pi = mine[i], mi = ms[i], vi = vs[i];
      upd(gsum[i], pi, mi, vi);
      ms[i] = mi;
      vs[i] = vi;
      for (int q = 0; q < world; ++q) params.p[q][base + i] = pi;
    }
  __threadfence_system();
}

extern "C" {

int subgnn_dp_reduce_scatter(const unsigned long long* peer_grads, const unsigned long long* peer_slots, const float* mc_grads, int world, int rank,
                             long long n, long long shard, float* gsum, void* stream) {
  SG_REQUIRE(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, "bad world / rank");
  SG_REQUIRE(shard > 0 && (shard % 4) == 0 && shard * world >= n, "shard must be a multiple of 4 floats covering the arena");
  DpPtrs g, s;
  for (int q = 0; q < DP_MAX_WORLD; ++q) {
    g.p[q] = q < world ? reinterpret_cast<float*>(peer_grads[q]) : nullptr;
    s.p[q] = q < world ? reinterpret_cast<float*>(peer_slots[q]) : nullptr;
  }
  int grid = sg_grid_for(shard / 4, 256, 4);
  if (grid > DP_MAX_BLOCKS) grid = DP_MAX_BLOCKS;
  sg_launch_pdl<SG_PDL_CHAIN>(dp_reduce_scatter_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, g, s, mc_grads, world, rank, n, shard, gsum);
  return subgnn_check_launch("dp_reduce_scatter_kernel");
}

int subgnn_dp_adam_allgather(const unsigned long long* peer_params, float* mc_params, int world, int rank, long long n, long long shard, const float* gsum, float* m,
                             float* v, float lr, float beta1, float beta2, float eps, const int* step_dev, const float* slots_local,
                             float clip_norm, float grad_scale, void* stream) {
  SG_REQUIRE(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, "bad world / rank");
  SG_REQUIRE(shard > 0 && (shard % 4) == 0 && shard * world >= n, "shard must be a multiple of 4 floats covering the arena");
  DpPtrs p;
  for (int q = 0; q < DP_MAX_WORLD; ++q) p.p[q] = q < world ? reinterpret_cast<float*>(peer_params[q]) : nullptr;
  int grid = sg_grid_for(shard / 4, 256, 4);
  sg_launch_pdl<SG_PDL_CHAIN>(dp_adam_allgather_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p, mc_params, world, rank, n, shard, gsum, m, v, lr, beta1,
                              beta2, eps, step_dev, slots_local, clip_norm, grad_scale);
  return subgnn_check_launch("dp_adam_allgather_kernel");
}

}  // extern "C"
