// Anchor-patch samplers for the neighbourhood and position channels + k-hop border sets.
//
// Replaces (reference):
//   anchor_patch_samplers.py:163-198  sample_neighborhood_anchor_patch   -> subgnn_sample_rows
//   anchor_patch_samplers.py:200-208  sample_position_anchor_patches     -> subgnn_sample_rows (pad_rule = 0)
//   subgraph_utils.py:146-176 / SubGNN.py:673-700  k-hop border sets     -> subgnn_border_khop_bitmap / _expand
//
// The reference's randn/argmax construction (rand[pad] = 0; argmax) returns PAD with probability 2^-k for a
// row with k valid entries and at least one pad column; the same law is sampled directly here:
// one Philox block per (row, anchor): word1's top k bits all zero -> PAD, else word0 -> uniform index.
#include "common.cuh"
#include "../../include/subgnn_b200.h"

__global__ void sample_rows_kernel(const int* __restrict__ ptr, const int* __restrict__ items, int n_rows, int width,
                                   int n_anchors, int pad_rule, unsigned long long seed, unsigned step, unsigned tag,
                                   int* __restrict__ out) {
  const long long total = (long long)n_rows * n_anchors;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(t / n_anchors);
    const int beg = ptr[row], k = ptr[row + 1] - beg;
    int v = 0;
    if (k > 0) {
      const Philox4 r = sg_draw(seed, (uint64_t)t, step, tag);
      const bool pad = pad_rule && k < width && k < 32 && (r.y >> (32 - k)) == 0;
      if (!pad) v = items[beg + sg_index(r.x, (uint32_t)k)];
    }
    out[t] = v;
  }
}

// ---- k-hop border: one CTA per component, visited/frontier bitmaps in shared memory ------------------
__global__ void __launch_bounds__(256)
border_khop_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int n_nodes, const int* __restrict__ cc_ptr,
                   const int* __restrict__ cc_nodes, int n_cc, int k, unsigned* __restrict__ bitmaps, int* __restrict__ counts) {
  extern __shared__ unsigned sm[];
  const int words = (n_nodes + 31) / 32;
  unsigned* visited = sm;
  unsigned* frontier = sm + words;
  unsigned* next = sm + 2 * words;
  __shared__ int s_count;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int c = blockIdx.x; c < n_cc; c += gridDim.x) {
    __syncthreads();
    for (int w = threadIdx.x; w < words; w += blockDim.x) { visited[w] = 0; frontier[w] = 0; next[w] = 0; }
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const int beg = cc_ptr[c], end = cc_ptr[c + 1];
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const int u = cc_nodes[i] - 1;
      atomicOr(&visited[u >> 5], 1u << (u & 31));
      atomicOr(&frontier[u >> 5], 1u << (u & 31));
    }
    __syncthreads();
    for (int hop = 0; hop < k; ++hop) {
      // warps take frontier words; lanes stride over the neighbours of each frontier node
      for (int w = warp; w < words; w += nwarps) {
        unsigned bits = frontier[w];
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const int u = w * 32 + b;
          const int off = rowptr[u], deg = rowptr[u + 1] - off;
          for (int i = lane; i < deg; i += 32) {
            const int v = col[off + i];
            const unsigned m = 1u << (v & 31);
            if (!(visited[v >> 5] & m)) atomicOr(&next[v >> 5], m);
          }
        }
      }
      __syncthreads();
      for (int w = threadIdx.x; w < words; w += blockDim.x) {
        const unsigned nw = next[w] & ~visited[w];
        visited[w] |= nw;
        frontier[w] = nw;
        next[w] = 0;
      }
      __syncthreads();
    }
    // border = reached minus the component itself
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const int u = cc_nodes[i] - 1;
      atomicAnd(&visited[u >> 5], ~(1u << (u & 31)));
    }
    __syncthreads();
    int local = 0;
    for (int w = threadIdx.x; w < words; w += blockDim.x) {
      const unsigned v = visited[w];
      bitmaps[(size_t)c * words + w] = v;
      local += __popc(v);
    }
    local = warp_sum_i(local);
    if (lane == 0) atomicAdd(&s_count, local);
    __syncthreads();
    if (threadIdx.x == 0) counts[c] = s_count;
  }
}

// bitmap -> ascending 1-indexed id list at out[out_ptr[c] ...]
__global__ void border_expand_kernel(const unsigned* __restrict__ bitmaps, int n_nodes, int n_cc, const long long* __restrict__ out_ptr,
                                     int* __restrict__ out) {
  const int words = (n_nodes + 31) / 32;
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  for (int c = warp_global; c < n_cc; c += warps_total) {
    long long base = out_ptr[c];
    for (int w0 = 0; w0 < words; w0 += 32) {
      const int w = w0 + lane;
      const unsigned v = w < words ? bitmaps[(size_t)c * words + w] : 0u;
      const int cnt = __popc(v);
      int pre = cnt;                                             // inclusive scan over lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += t;
      }
      long long pos = base + pre - cnt;
      unsigned bits = v;
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        out[pos++] = w * 32 + b + 1;
      }
      base += __shfl_sync(0xffffffffu, pre, 31);
    }
  }
}

extern "C" {

int subgnn_sample_rows(const int* ptr, const int* items, int n_rows, int width, int n_anchors, int pad_rule,
                       unsigned long long seed, unsigned step, int position_stream, int* out, void* stream) {
  SG_REQUIRE(n_rows >= 0 && n_anchors >= 1, "bad sizes");
  if (n_rows == 0) return SUBGNN_OK;
  sample_rows_kernel<<<sg_grid_for((long long)n_rows * n_anchors, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      ptr, items, n_rows, width, n_anchors, pad_rule, seed, step, position_stream ? SG_TAG_POS : SG_TAG_NEIGH, out);
  return subgnn_check_launch("sample_rows_kernel");
}

int subgnn_border_khop_bitmap(const int* rowptr, const int* col, int n_nodes, const int* cc_ptr, const int* cc_nodes, int n_cc, int k,
                              unsigned* bitmaps, int* counts, void* stream) {
  SG_REQUIRE(n_nodes > 0 && n_cc >= 0 && k >= 0, "bad sizes");
  if (n_cc == 0) return SUBGNN_OK;
  const size_t smem = (size_t)3 * ((n_nodes + 31) / 32) * sizeof(unsigned);
  SG_REQUIRE(smem <= 200 * 1024, "graph too large for shared-memory bitmaps (> 546k nodes)");
  cudaFuncSetAttribute(border_khop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  border_khop_kernel<<<sg_grid_for(n_cc, 1, 4), 256, smem, (cudaStream_t)stream>>>(rowptr, col, n_nodes, cc_ptr, cc_nodes, n_cc, k, bitmaps, counts);
  return subgnn_check_launch("border_khop_kernel");
}

int subgnn_border_khop_expand(const unsigned* bitmaps, int n_nodes, int n_cc, const long long* out_ptr, int* out, void* stream) {
  if (n_cc == 0) return SUBGNN_OK;
  border_expand_kernel<<<sg_grid_for(n_cc, 8, 8), 256, 0, (cudaStream_t)stream>>>(bitmaps, n_nodes, n_cc, out_ptr, out);
  return subgnn_check_launch("border_expand_kernel");
}

}  // extern "C"
