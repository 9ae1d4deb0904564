// Data-parallel gradient exchange fused with the optimizer, over NVLink peer memory (no NCCL on the step path).
//
// Reference semantics: every process steps its shard of the global batch and the gradients are averaged before
// torch.optim.Adam / clip_grad_norm_ (SubGNN.py:1156-1164 under Lightning's distributed back-ends).  Round 1 did
// all-reduce (NCCL) -> sum of squares -> Adam: the 4 flat arenas streamed three times and a library collective between two
// graph halves.  Here the gradient arenas of all ranks are SYMMETRIC allocations (torch.distributed._symmetric_memory: every rank
// maps every peer's arena), and the exchange is two kernels with the cross-GPU barriers inside them (epoch flags in symmetric
// memory: a system fence + flag stores into the peers' flag arrays, polls of the own one + a system fence — no barrier launches):
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
#define DP_UNROLL 4          // 16-byte words per thread and round: DP_UNROLL x world peer loads in flight

struct DpPtrs {
  float* p[DP_MAX_WORLD];
};
struct DpFlags {
  unsigned* p[DP_MAX_WORLD];   // every rank's flag array [3][DP_MAX_WORLD] (symmetric): [phase][sender]
};

__device__ float dp_part[DP_MAX_BLOCKS];
__device__ unsigned dp_ticket = 0, dp_ticket2 = 0;

// peer memory is read with ld.volatile (never served from a stale cache line of the previous step) and written with plain stores
// followed by a system-scope fence
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

// Cross-GPU barriers inside the kernels: flags[phase][sender] in every rank's symmetric flag array carry the epoch (a device
// counter every rank advances once per exchange).  signal = release store of the epoch into my slot of every peer's array,
// wait = acquire loads of my own array until every sender's slot has reached the epoch.  A waiter spins for at most ~10 s
// (a lost peer traps the kernel instead of hanging the GPU).
__device__ __forceinline__ void dp_signal(const DpFlags& f, int phase, int world, int rank, unsigned epoch) {
  __threadfence_system();                                            // ONE fence, then plain (volatile) flag stores: a release store per
  for (int q = 0; q < world; ++q)                                    // peer would drain the fabric `world` times in a row
    *reinterpret_cast<volatile unsigned*>(f.p[q] + phase * DP_MAX_WORLD + rank) = epoch;
}
__device__ __forceinline__ void dp_wait(const unsigned* mine, int phase, int world, unsigned epoch) {
  const long long t0 = clock64();
  for (int q = 0; q < world; ++q) {
    while ((int)(*reinterpret_cast<volatile const unsigned*>(mine + phase * DP_MAX_WORLD + q) - epoch) < 0)
      if (clock64() - t0 > 20000000000ll) __trap();                 // ~10 s: a peer is gone
  }
  __threadfence_system();                                            // acquire: the data the flags announce is visible to what follows
}

// phase 0: "my gradient arena is complete" (signalled by the first block, which runs after every kernel of my backward pass);
// phase 1: "my shard's sum of squares is in your slot and I am done reading your gradients" (last block).
__global__ void __launch_bounds__(256) dp_reduce_scatter_kernel(DpPtrs grads, DpPtrs slots, DpFlags flags, const float* __restrict__ mc_grads,
                                                               const unsigned* __restrict__ epoch_dev, int world, int rank, long long n, long long shard,
                                                               float* __restrict__ gsum) {
  sg_pdl_sync();
  const bool sync = epoch_dev != nullptr;                           // nullptr: the caller brackets the kernels with its own cross-GPU barriers
  const unsigned epoch = sync ? *epoch_dev : 0u;
  if (sync) {
    if (threadIdx.x == 0) {
      if (blockIdx.x == 0) dp_signal(flags, 0, world, rank, epoch);
      dp_wait(flags.p[rank], 0, world, epoch);
    }
    __syncthreads();
  }
  const long long base = (long long)rank * shard;
  const long long len = max(0LL, min(shard, n - base));            // elements of my shard (multiple of 4 except at the arena's end)
  const long long len4 = len / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float s = 0.f;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < len4; i0 += DP_UNROLL * stride) {
    float4 acc[DP_UNROLL];
    if (mc_grads) {
#pragma unroll
      for (int u = 0; u < DP_UNROLL; ++u)                            // in-switch reduction: one load instead of `world`
        acc[u] = (i0 + u * stride < len4) ? ld_reduce_mc(mc_grads + base + 4 * (i0 + u * stride)) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
      for (int u = 0; u < DP_UNROLL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < world; ++q) {                             // fixed order: identical sums whatever the launch geometry
        float4 v[DP_UNROLL];
#pragma unroll
        for (int u = 0; u < DP_UNROLL; ++u)
          v[u] = (i0 + u * stride < len4) ? ld_peer(grads.p[q] + base + 4 * (i0 + u * stride)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < DP_UNROLL; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
      }
    }
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u)
      if (i0 + u * stride < len4) {
        reinterpret_cast<float4*>(gsum)[i0 + u * stride] = acc[u];
        s = fmaf(acc[u].x, acc[u].x, s); s = fmaf(acc[u].y, acc[u].y, s); s = fmaf(acc[u].z, acc[u].z, s); s = fmaf(acc[u].w, acc[u].w, s);
      }
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
    dp_ticket = 0;
    if (sync) dp_signal(flags, 1, world, rank, epoch);              // (fence inside) every read of the peers' gradients has completed
    else __threadfence_system();
  }
}

// waits for phase 1 of every rank, updates shard `rank`, writes it into every rank's parameter arena, last block signals phase 2
__global__ void __launch_bounds__(256)
dp_adam_allgather_kernel(DpPtrs params, DpFlags flags, float* __restrict__ mc_params, const unsigned* __restrict__ epoch_dev, int world, int rank,
                         long long n, long long shard, const float* __restrict__ gsum, float* __restrict__ m, float* __restrict__ v, float lr,
                         float beta1, float beta2, float eps, const int* __restrict__ step_dev, const float* slots_local, float clip_norm,
                         float grad_scale) {
  sg_pdl_sync();
  const bool sync = epoch_dev != nullptr;
  const unsigned epoch = sync ? *epoch_dev : 0u;
  if (sync) {
    if (threadIdx.x == 0) dp_wait(flags.p[rank], 1, world, epoch);
    __syncthreads();
  }
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
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < len4; i0 += 2 * stride) {
    float4 g4[2], p4[2], m4[2], v4[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i = i0 + u * stride;
      if (i < len4) {
        g4[u] = reinterpret_cast<const float4*>(gsum)[i];
        p4[u] = reinterpret_cast<float4*>(mine)[i]; m4[u] = reinterpret_cast<float4*>(ms)[i]; v4[u] = reinterpret_cast<float4*>(vs)[i];
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i = i0 + u * stride;
      if (i >= len4) continue;
      upd(g4[u].x, p4[u].x, m4[u].x, v4[u].x); upd(g4[u].y, p4[u].y, m4[u].y, v4[u].y);
      upd(g4[u].z, p4[u].z, m4[u].z, v4[u].z); upd(g4[u].w, p4[u].w, m4[u].w, v4[u].w);
      reinterpret_cast<float4*>(ms)[i] = m4[u];
      reinterpret_cast<float4*>(vs)[i] = v4[u];
      if (mc_params) st_mc(mc_params + base + 4 * i, p4[u]);                                      // one store, every rank's arena
      else {
#pragma unroll 8
        for (int q = 0; q < world; ++q) reinterpret_cast<float4*>(params.p[q] + base)[i] = p4[u]; // my shard of every rank's arena
      }
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
  // every block fences its peer stores, the last one tells every rank that shard `rank` of its parameter arena is written
  __threadfence_system();
  if (!sync) return;
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&dp_ticket2, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    dp_ticket2 = 0;
    dp_signal(flags, 2, world, rank, epoch);
  }
}

// closes the exchange: every shard of my parameter arena has been written (phase 2 of every rank); advances the epoch
__global__ void dp_finish_kernel(const unsigned* mine, unsigned* epoch_dev, int world) {
  sg_pdl_sync();
  if (threadIdx.x == 0) {
    const unsigned epoch = *epoch_dev;
    dp_wait(mine, 2, world, epoch);
    *epoch_dev = epoch + 1;
  }
}

extern "C" {

int subgnn_dp_flag_words(void) { return 3 * DP_MAX_WORLD; }

int subgnn_dp_reduce_scatter(const unsigned long long* peer_grads, const unsigned long long* peer_slots, const unsigned long long* peer_flags,
                             const float* mc_grads, const unsigned* epoch_dev, int world, int rank, long long n, long long shard, float* gsum,
                             void* stream) {
  SG_REQUIRE(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, "bad world / rank");
  SG_REQUIRE(shard > 0 && (shard % 4) == 0 && shard * world >= n, "shard must be a multiple of 4 floats covering the arena");
  DpPtrs g, s;
  DpFlags f;
  for (int q = 0; q < DP_MAX_WORLD; ++q) {
    g.p[q] = q < world ? reinterpret_cast<float*>(peer_grads[q]) : nullptr;
    s.p[q] = q < world ? reinterpret_cast<float*>(peer_slots[q]) : nullptr;
    f.p[q] = q < world ? reinterpret_cast<unsigned*>(peer_flags[q]) : nullptr;
  }
  int grid = sg_grid_for(shard / 4, 256 * DP_UNROLL, 2);             // every block resident: waiting blocks cannot starve signalling ones
  if (grid > DP_MAX_BLOCKS) grid = DP_MAX_BLOCKS;
  sg_launch_pdl<SG_PDL_CHAIN>(dp_reduce_scatter_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, g, s, f, mc_grads, epoch_dev, world, rank, n,
                              shard, gsum);
  return subgnn_check_launch("dp_reduce_scatter_kernel");
}

int subgnn_dp_adam_allgather(const unsigned long long* peer_params, const unsigned long long* peer_flags, float* mc_params, unsigned* epoch_dev, int world,
                             int rank, long long n, long long shard, const float* gsum, float* m, float* v, float lr, float beta1, float beta2,
                             float eps, const int* step_dev, const float* slots_local, float clip_norm, float grad_scale, void* stream) {
  SG_REQUIRE(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, "bad world / rank");
  SG_REQUIRE(shard > 0 && (shard % 4) == 0 && shard * world >= n, "shard must be a multiple of 4 floats covering the arena");
  DpPtrs p;
  DpFlags f;
  for (int q = 0; q < DP_MAX_WORLD; ++q) {
    p.p[q] = q < world ? reinterpret_cast<float*>(peer_params[q]) : nullptr;
    f.p[q] = q < world ? reinterpret_cast<unsigned*>(peer_flags[q]) : nullptr;
  }
  int grid = sg_grid_for(shard / 4, 256 * 2, 2);
  if (grid > DP_MAX_BLOCKS) grid = DP_MAX_BLOCKS;
  sg_launch_pdl<SG_PDL_CHAIN>(dp_adam_allgather_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p, f, mc_params, (const unsigned*)epoch_dev, world, rank,
                              n, shard, gsum, m, v, lr, beta1, beta2, eps, step_dev, slots_local, clip_norm, grad_scale);
  int rc = subgnn_check_launch("dp_adam_allgather_kernel");
  if (rc || !epoch_dev) return rc;
  sg_launch_pdl<SG_PDL_CHAIN>(dp_finish_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (const unsigned*)f.p[rank], epoch_dev, world);
  return subgnn_check_launch("dp_finish_kernel");
}

}  // extern "C"
