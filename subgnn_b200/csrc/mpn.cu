// Stand-alone SG_MPN message passing (module-level drop-in): fused similarity lookup + gather-scale-reduce.
//
// Replaces (reference) subgraph_mpn.py:36-131,176-231 for one SG_MPN.forward call with the reference's own
// (materialised) inputs:
//   create_patch_embedding_matrix / create_edge_index (:36-71)  -> implicit: edge (r, a) exists iff mask[r][a]
//   get_similarities (:73-103)                                  -> s = sims[r][anchor_id - 1] (N/P) | sims[r][sim_index[a]] (S)
//   message + scatter-add aggregate (:227-231, torch_scatter)   -> agg[r] = sum_a m s x[r][a]
//   generate_pos_struc_embeddings (:105-123)                    -> pos_lin[r][a] = w_p . (m s x[r][a]) + b_p
// The update relu(W [x ; agg] + b) (:233-241) runs through subgnn_linear_fwd on the [cc | agg] buffer this kernel
// fills.  Rows are ALL B*C rows of the padded layout (padded components have an all-false mask, so agg = 0 and
// pos_lin = b_p exactly as in the reference).  One warp per row; lanes own contiguous slices of D (vector loads).
#include "common.cuh"
#include "../../include/subgnn_b200.h"

#define MPN_WARPS 8

// cat[r] = [cc[r] | agg[r]] (R x 2D), pos_lin (R x A), s_eff (R x A) = m * s saved for the backward pass
__global__ void __launch_bounds__(MPN_WARPS * 32)
mpn_fwd_kernel(const float* __restrict__ cc, const float* __restrict__ x /*[R][A][D]*/, const float* __restrict__ sims, int n_opt,
               const int* __restrict__ anchor_ids /*[R][A] or null*/, const int* __restrict__ sim_index /*[A] or null*/,
               const unsigned char* __restrict__ mask /*[R][A]*/, const float* __restrict__ wp, const float* __restrict__ bp,
               float* __restrict__ cat, float* __restrict__ pos_lin, float* __restrict__ s_eff, int R, int A, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float b = bp[0];
  for (int r = blockIdx.x * MPN_WARPS + warp; r < R; r += gridDim.x * MPN_WARPS) {
    for (int k = lane; k < D; k += 32) cat[(size_t)r * 2 * D + k] = cc[(size_t)r * D + k];
    for (int k0 = 0; k0 < D; k0 += 32) {
      // each pass owns a 32-wide slice of D; anchors are streamed once per slice (D <= 128 => <= 4 passes)
      const int k = k0 + lane;
      float agg = 0.f;
      const float w = k < D ? wp[k] : 0.f;
      for (int a = 0; a < A; ++a) {
        const bool m = mask[(size_t)r * A + a] != 0;
        float s = 0.f;
        if (m) {
          const int col = sim_index ? sim_index[a] : anchor_ids[(size_t)r * A + a] - 1;
          s = sims[(size_t)r * n_opt + col];
        }
        const float v = (m && k < D) ? s * x[((size_t)r * A + a) * D + k] : 0.f;
        agg += v;
        float dot = warp_sum(v * w);
        if (lane == 0) {
          if (k0 == 0) { pos_lin[(size_t)r * A + a] = dot + b; s_eff[(size_t)r * A + a] = s; }
          else pos_lin[(size_t)r * A + a] += dot;
        }
      }
      if (k < D) cat[(size_t)r * 2 * D + D + k] = agg;
    }
  }
}

// dx[r][a] = s_eff (dagg[r] + dpos[r][a] w_p);  dwp += sum dpos s_eff x;  dbp += sum dpos
__global__ void __launch_bounds__(MPN_WARPS * 32)
mpn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ s_eff, const float* __restrict__ dcat /*[R][2D], dagg at +D*/,
               const float* __restrict__ dpos, const float* __restrict__ wp, float* __restrict__ dx, float* __restrict__ dwp,
               float* __restrict__ dbp, int R, int A, int D) {
  __shared__ float s_dwp[MPN_WARPS][256];
  __shared__ float s_dbp[MPN_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc_w[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc_w[q] = 0.f;
  float acc_b = 0.f;
  for (int r = blockIdx.x * MPN_WARPS + warp; r < R; r += gridDim.x * MPN_WARPS) {
    for (int a = 0; a < A; ++a) {
      const float s = s_eff[(size_t)r * A + a];
      const float g = dpos[(size_t)r * A + a];
      if (lane == 0) acc_b += g;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = lane + 32 * q;
        if (k < D) {
          const size_t xi = ((size_t)r * A + a) * D + k;
          dx[xi] = s * (dcat[(size_t)r * 2 * D + D + k] + g * wp[k]);
          acc_w[q] = fmaf(g * s, x[xi], acc_w[q]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) s_dwp[warp][lane + 32 * q] = acc_w[q];
  if (lane == 0) s_dbp[warp] = acc_b;
  __syncthreads();
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < MPN_WARPS; ++w) s += s_dwp[w][k];
    if (s != 0.f) atomicAdd(dwp + k, s);
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < MPN_WARPS; ++w) s += s_dbp[w];
    if (s != 0.f) atomicAdd(dbp, s);
  }
}

extern "C" {

int subgnn_mpn_fwd(const float* cc, const float* anchor_embeds, const float* sims, int n_opt, const int* anchor_ids, const int* sim_index,
                   const unsigned char* mask, const float* wp, const float* bp, float* cat, float* pos_lin, float* s_eff, int R, int A,
                   int D, void* stream) {
  SG_REQUIRE(R >= 0 && A >= 1 && D >= 1 && D <= 256, "bad sizes");
  SG_REQUIRE((anchor_ids != nullptr) != (sim_index != nullptr), "exactly one of anchor_ids / sim_index");
  if (R == 0) return SUBGNN_OK;
  mpn_fwd_kernel<<<sg_grid_for(R, MPN_WARPS, 8), MPN_WARPS * 32, 0, (cudaStream_t)stream>>>(cc, anchor_embeds, sims, n_opt, anchor_ids, sim_index,
                                                                                          mask, wp, bp, cat, pos_lin, s_eff, R, A, D);
  return subgnn_check_launch("mpn_fwd_kernel");
}

int subgnn_mpn_bwd(const float* anchor_embeds, const float* s_eff, const float* dcat, const float* dpos, const float* wp, float* dx, float* dwp,
                   float* dbp, int R, int A, int D, void* stream) {
  SG_REQUIRE(R >= 0 && A >= 1 && D >= 1 && D <= 256, "bad sizes");
  if (R == 0) return SUBGNN_OK;
  mpn_bwd_kernel<<<sg_grid_for(R, MPN_WARPS, 4), MPN_WARPS * 32, 0, (cudaStream_t)stream>>>(anchor_embeds, s_eff, dcat, dpos, wp, dx, dwp, dbp, R, A, D);
  return subgnn_check_launch("mpn_bwd_kernel");
}

}  // extern "C"
