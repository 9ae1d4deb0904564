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

int subgnn_sum_to_scalar(const float* x, int n, float* out, void* stream) {
  sg_launch_pdl<SG_PDL_CHAIN>(sum_to_scalar_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, x, n, out);
  return subgnn_check_launch("sum_to_scalar_kernel");
}

int subgnn_inc_step(int* step_dev, void* stream) {
  inc_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_dev);
  return subgnn_check_launch("inc_step_kernel");
}

}  // extern "C"
