// Optimizer step over the flat parameter arena: gradient-norm clipping + Adam.
//
// Replaces (reference): SubGNN.py:1156-1161 torch.optim.Adam(lr) (betas 0.9/0.999, eps 1e-8, no weight decay)
// and pytorch-lightning 0.7.1's gradient_clip_val -> torch.nn.utils.clip_grad_norm_ (train_config.py:109-158):
//   total_norm = ||g||_2 over all parameters; g *= min(1, clip / (total_norm + 1e-6)).
// All parameters live in ONE contiguous fp32 arena (same for grads / m / v), so the whole optimizer is two
// streaming kernels: a sum-of-squares reduction and a fused clip+Adam update (reads p, g, m, v; writes p, m, v:
// 28 B per parameter, HBM bound).  A parameter whose gradient is identically zero with zero state stays put,
// which matches torch skipping parameters whose .grad is None (P/S-channel `linear`, N-channel `linear_position`).
#include "common.cuh"
#include "../../include/subgnn_b200.h"

__global__ void fill_zero_kernel(float4* __restrict__ p4, long long n4, float* __restrict__ tail, int ntail) {
  sg_pdl_sync();
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) p4[i] = z;
  if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] = 0.f;
}

// Deterministic: every block writes its partial sum, the LAST block to finish (ticket) adds the partials in a fixed order.  The result
// feeds the clip coefficient of every parameter, so an order-dependent sum (float atomics) lets data-parallel ranks — which hold
// bit-identical all-reduced gradients — drift apart by a few ulps per step.
#define SUMSQ_MAX_BLOCKS 2048
__device__ float sg_sumsq_part[SUMSQ_MAX_BLOCKS];
__device__ unsigned sg_sumsq_ticket = 0;
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  sg_pdl_sync();
  float s = 0.f;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s = fmaf(g[i], g[i], s);
  __shared__ float ws[8];
  __shared__ bool last;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += ws[w];
    sg_sumsq_part[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(&sg_sumsq_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += 256) t += __ldcg(sg_sumsq_part + i);      // fixed assignment, fixed order
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += ws[w];
    *out += tot;
    sg_sumsq_ticket = 0;                                      // ready for the next launch (launches of this kernel are stream ordered)
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
            float beta1, float beta2, float eps, const int* __restrict__ step_dev, const float* __restrict__ sumsq, float clip_norm,
            float grad_scale) {
  sg_pdl_sync();
  const int t = *step_dev;
  float coef = grad_scale;
  if (clip_norm > 0.f && sumsq) {
    const float total = sqrtf(*sumsq) * grad_scale;           // norm of the (already averaged) gradient
    const float c = clip_norm / (total + 1e-6f);
    if (c < 1.f) coef *= c;
  }
  const float bc1 = 1.f - powf(beta1, (float)t);
  const float bc2 = 1.f - powf(beta2, (float)t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  auto upd = [&](float gi_raw, float& pi, float& mi_, float& vi_) {
    const float gi = gi_raw * coef;
    const float mi = beta1 * mi_ + (1.f - beta1) * gi;
    const float vi = beta2 * vi_ + (1.f - beta2) * gi * gi;
    mi_ = mi;
    vi_ = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi -= step_size * (mi / denom);
  };
  const long long n4 = n >> 2;                    // the arenas are 16-byte aligned (checked by the host wrapper)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    upd(g4.x, p4.x, m4.x, v4.x); upd(g4.y, p4.y, m4.y, v4.y); upd(g4.z, p4.z, m4.z, v4.z); upd(g4.w, p4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(p)[i] = p4; reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4;
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    upd(g[i], p[i], m[i], v[i]);
}

// ------------------------------------------------------------------------------------------------
// clip_grad_norm_ + Adam in ONE launch [r2]: one CTA per SM, each owning a contiguous slice of the arena.
//   phase 1  the slice's gradients are read ONCE into registers (CA_KEEP float4 per thread; longer slices re-read), their
//            sum of squares goes to sg_ca_part[cta]; the slice's p / m / v loads are issued BEFORE the barrier so their HBM
//            latency overlaps the reduction and the barrier
//   barrier  sense-reversing grid barrier (arrive counter + epoch word); all CTAs are co-resident (grid <= SM count, one CTA
//            fits every SM once the producer kernels have drained — nothing this kernel waits for waits for it)
//   phase 2  every CTA adds the partials in the same fixed order (bit-identical coefficient everywhere), then updates its slice.
// Against sumsq_kernel + adam_kernel this removes one launch boundary and the second pass over the gradients.
#define CA_THREADS 512
#define CA_KEEP 6
#define CA_MAX_BLOCKS 256
__device__ float sg_ca_part[CA_MAX_BLOCKS];
__device__ unsigned sg_ca_arrive = 0;
__device__ unsigned sg_ca_epoch = 0;

__global__ void __launch_bounds__(CA_THREADS, 1)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                 float beta1, float beta2, float eps, const int* __restrict__ step_dev, float* __restrict__ sumsq_out, float clip_norm,
                 float grad_scale) {
  sg_pdl_sync();
  const int tid = threadIdx.x;
  const unsigned epoch0 = *reinterpret_cast<volatile unsigned*>(&sg_ca_epoch);     // cannot advance before this CTA arrives
  const long long n4 = n >> 2;
  const long long per = (n4 + gridDim.x - 1) / gridDim.x;
  const long long lo = blockIdx.x * per, hi = lo + per < n4 ? lo + per : n4;
  const float4* g4p = reinterpret_cast<const float4*>(g);
  float4* p4p = reinterpret_cast<float4*>(p);
  float4* m4p = reinterpret_cast<float4*>(m);
  float4* v4p = reinterpret_cast<float4*>(v);
  float4 gk[CA_KEEP], pk[CA_KEEP], mk[CA_KEEP], vk[CA_KEEP];
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < CA_KEEP; ++u) {
    const long long i = lo + tid + (long long)u * CA_THREADS;
    gk[u] = i < hi ? g4p[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < CA_KEEP; ++u) {
    const long long i = lo + tid + (long long)u * CA_THREADS;
    if (i < hi) { pk[u] = p4p[i]; mk[u] = m4p[i]; vk[u] = v4p[i]; }
  }
#pragma unroll
  for (int u = 0; u < CA_KEEP; ++u) {
    s = fmaf(gk[u].x, gk[u].x, s); s = fmaf(gk[u].y, gk[u].y, s); s = fmaf(gk[u].z, gk[u].z, s); s = fmaf(gk[u].w, gk[u].w, s);
  }
  for (long long i = lo + tid + (long long)CA_KEEP * CA_THREADS; i < hi; i += CA_THREADS) {     // slices longer than the register window
    const float4 x = g4p[i];
    s = fmaf(x.x, x.x, s); s = fmaf(x.y, x.y, s); s = fmaf(x.z, x.z, s); s = fmaf(x.w, x.w, s);
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + tid; i < n; i += CA_THREADS) s = fmaf(g[i], g[i], s);
  __shared__ float ws[CA_THREADS / 32];
  __shared__ float s_tot;
  s = warp_sum(s);
  if ((tid & 31) == 0) ws[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < CA_THREADS / 32; ++w) t += ws[w];
    sg_ca_part[blockIdx.x] = t;
    __threadfence();
    if (atomicAdd(&sg_ca_arrive, 1u) == gridDim.x - 1) {
      sg_ca_arrive = 0;                                        // every CTA of this launch has arrived: none touches it again
      __threadfence();
      atomicExch(&sg_ca_epoch, epoch0 + 1u);
    } else {
      const long long t0 = clock64();
      while (*reinterpret_cast<volatile unsigned*>(&sg_ca_epoch) == epoch0)
        if (clock64() - t0 > 20000000000LL) __trap();          // ~10 s: a CTA of the grid never became resident
    }
    __threadfence();
  }
  __syncthreads();
  if (tid < 32) {
    float t = 0.f;
    for (unsigned i = tid; i < gridDim.x; i += 32) t += __ldcg(sg_ca_part + i);        // same assignment and order in every CTA
    t = warp_sum(t);
    if (tid == 0) s_tot = t;
  }
  __syncthreads();
  const float total_sq = s_tot;
  if (blockIdx.x == 0 && tid == 0 && sumsq_out) *sumsq_out += total_sq;
  const int t = *step_dev;
  float coef = grad_scale;
  if (clip_norm > 0.f) {
    const float total = sqrtf(total_sq) * grad_scale;
    const float c = clip_norm / (total + 1e-6f);
    if (c < 1.f) coef *= c;
  }
  const float bc1 = 1.f - powf(beta1, (float)t);
  const float bc2 = 1.f - powf(beta2, (float)t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  auto upd = [&](float gi_raw, float& pi, float& mi_, float& vi_) {
    const float gi = gi_raw * coef;
    const float mi = beta1 * mi_ + (1.f - beta1) * gi;
    const float vi = beta2 * vi_ + (1.f - beta2) * gi * gi;
    mi_ = mi;
    vi_ = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi -= step_size * (mi / denom);
  };
#pragma unroll
  for (int u = 0; u < CA_KEEP; ++u) {
    const long long i = lo + tid + (long long)u * CA_THREADS;
    if (i < hi) {
      upd(gk[u].x, pk[u].x, mk[u].x, vk[u].x); upd(gk[u].y, pk[u].y, mk[u].y, vk[u].y);
      upd(gk[u].z, pk[u].z, mk[u].z, vk[u].z); upd(gk[u].w, pk[u].w, mk[u].w, vk[u].w);
      p4p[i] = pk[u]; m4p[i] = mk[u]; v4p[i] = vk[u];
    }
  }
  for (long long i = lo + tid + (long long)CA_KEEP * CA_THREADS; i < hi; i += CA_THREADS) {
    const float4 x = g4p[i];
    float4 a = p4p[i], b = m4p[i], c = v4p[i];
    upd(x.x, a.x, b.x, c.x); upd(x.y, a.y, b.y, c.y); upd(x.z, a.z, b.z, c.z); upd(x.w, a.w, b.w, c.w);
    p4p[i] = a; m4p[i] = b; v4p[i] = c;
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + tid; i < n; i += CA_THREADS) upd(g[i], p[i], m[i], v[i]);
}

__global__ void sum_to_scalar_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
  sg_pdl_sync();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  __shared__ float ws[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) *out = t;
  }
}

__global__ void inc_step_kernel(int* step) { if (threadIdx.x == 0 && blockIdx.x == 0) *step += 1; }

extern "C" {

int subgnn_fill_zero(float* p, long long n, void* stream) {
  if (n <= 0) return SUBGNN_OK;
  SG_REQUIRE(((size_t)p & 15) == 0, "buffer must be 16-byte aligned");
  const long long n4 = n / 4;
  sg_launch_pdl(fill_zero_kernel, dim3(sg_grid_for(n4 > 0 ? n4 : 1, 256, 8)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<float4*>(p), n4, p + n4 * 4, (int)(n - n4 * 4));
  return subgnn_check_launch("fill_zero_kernel");
}

int subgnn_grad_sumsq(const float* g, long long n, float* out_sumsq, void* stream) {
  if (n <= 0) return SUBGNN_OK;
  SG_REQUIRE(((size_t)g & 15) == 0, "buffer must be 16-byte aligned");
  int grid = sg_grid_for(n / 4 + 1, 256, 4);
  if (grid > SUMSQ_MAX_BLOCKS) grid = SUMSQ_MAX_BLOCKS;
  sg_launch_pdl<SG_PDL_CHAIN>(sumsq_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, g, n, out_sumsq);
  return subgnn_check_launch("sumsq_kernel");
}

int subgnn_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                     const int* step_dev, const float* sumsq_dev, float clip_norm, float grad_scale, void* stream) {
  if (n <= 0) return SUBGNN_OK;
  SG_REQUIRE((((size_t)p | (size_t)g | (size_t)m | (size_t)v) & 15) == 0, "buffers must be 16-byte aligned");
  sg_launch_pdl<SG_PDL_CHAIN>(adam_kernel, dim3(sg_grid_for((n + 3) / 4, 256, 8)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, lr, beta1, beta2, eps, step_dev,
                sumsq_dev, clip_norm, grad_scale);
  return subgnn_check_launch("adam_kernel");
}

int subgnn_clip_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                          const int* step_dev, float* sumsq_dev, float clip_norm, float grad_scale, void* stream) {
  if (n <= 0) return SUBGNN_OK;
  SG_REQUIRE((((size_t)p | (size_t)g | (size_t)m | (size_t)v) & 15) == 0, "buffers must be 16-byte aligned");
  int grid = subgnn_sm_count();                                 // one CTA per SM: the grid barrier needs every CTA resident
  if (grid > CA_MAX_BLOCKS) grid = CA_MAX_BLOCKS;
  const long long n4 = n >> 2;
  if (n4 < (long long)grid * CA_THREADS) grid = (int)((n4 + CA_THREADS - 1) / CA_THREADS);
  if (grid < 1) grid = 1;
  sg_launch_pdl<SG_PDL_CHAIN>(clip_adam_kernel, dim3(grid), dim3(CA_THREADS), 0, (cudaStream_t)stream, p, g, m, v, n, lr, beta1, beta2, eps,
                              step_dev, sumsq_dev, clip_norm, grad_scale);
  return subgnn_check_launch("clip_adam_kernel");
}

int subgnn_sum_to_scalar(const float* x, int n, float* out, void* stream) {
  sg_launch_pdl<SG_PDL_CHAIN>(sum_to_scalar_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, x, n, out);
  return subgnn_check_launch("sum_to_scalar_kernel");
}

int subgnn_inc_step(int* step_dev, void* stream) {
  inc_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_dev);
  return subgnn_check_launch("inc_step_kernel");
}

}  // extern "C"
