// All-pairs hop table on the GPU: one CTA per source node, level-synchronous BFS with shared-memory
// visited / frontier bitmaps, uint8 output row.
//
// Replaces (reference, offline tooling one step before the hot path — SURVEY.md §8(f) rank 3):
//   prepare_dataset/precompute_graph_metrics.py:20-26  get_shortest_path (SNAP BFS per source, Pool) and the
//   fp64 N x N shortest_path_matrix.npy consumed at SubGNN.py:848.
// Conventions kept: hop[s][t] = BFS distance, 0 for t == s and for unreachable t; distances are
// clamped to 255 (never reached on the benchmark shapes: diameter <= 8).
#include "common.cuh"
#include "../../include/subgnn_b200.h"

__global__ void __launch_bounds__(256)
hop_table_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int n_nodes, int src_begin, int src_end,
                 unsigned char* __restrict__ hop, size_t hop_stride) {
  extern __shared__ unsigned sm[];
  const int words = (n_nodes + 31) / 32;
  unsigned* visited = sm;
  unsigned* frontier = sm + words;
  unsigned* next = sm + 2 * words;
  __shared__ int s_any;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int src = src_begin + blockIdx.x; src < src_end; src += gridDim.x) {
    unsigned char* out = hop + (size_t)src * hop_stride;
    __syncthreads();
    for (int w = threadIdx.x; w < words; w += blockDim.x) { visited[w] = 0; frontier[w] = 0; next[w] = 0; }
    for (int i = threadIdx.x; i < n_nodes; i += blockDim.x) out[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
      visited[src >> 5] = 1u << (src & 31);
      frontier[src >> 5] = 1u << (src & 31);
    }
    __syncthreads();
    for (int level = 1;; ++level) {
      const unsigned char lv = (unsigned char)(level > 255 ? 255 : level);
      if (threadIdx.x == 0) s_any = 0;
      __syncthreads();
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
            if (!(visited[v >> 5] & m)) {
              const unsigned old = atomicOr(&next[v >> 5], m);
              if (!(old & m)) out[v] = lv;               // first discovery in this level
            }
          }
        }
      }
      __syncthreads();
      int any = 0;
      for (int w = threadIdx.x; w < words; w += blockDim.x) {
        const unsigned nw = next[w];
        visited[w] |= nw;
        frontier[w] = nw;
        next[w] = 0;
        any |= (nw != 0);
      }
      if (any) s_any = 1;
      __syncthreads();
      if (!s_any) break;
    }
  }
}

extern "C" int subgnn_hop_table(const int* rowptr, const int* col, int n_nodes, int src_begin, int src_end, unsigned char* hop,
                                long long hop_stride, void* stream) {
  SG_REQUIRE(n_nodes > 0 && src_begin >= 0 && src_end <= n_nodes && hop_stride >= n_nodes, "bad sizes");
  if (src_end <= src_begin) return SUBGNN_OK;
  const size_t smem = (size_t)3 * ((n_nodes + 31) / 32) * sizeof(unsigned);
  SG_REQUIRE(smem <= 200 * 1024, "graph too large for shared-memory bitmaps");
  cudaFuncSetAttribute(hop_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  hop_table_kernel<<<sg_grid_for(src_end - src_begin, 1, 8), 256, smem, (cudaStream_t)stream>>>(rowptr, col, n_nodes, src_begin, src_end, hop,
                                                                                               (size_t)hop_stride);
  return subgnn_check_launch("hop_table_kernel");
}
