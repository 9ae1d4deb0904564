// Triangular random walks over the CSR base graph — warp-per-walk, counter-based Philox.
//
// Replaces (reference, all Python/networkx):
//   anchor_patch_samplers.py:49-113   triangular_random_walk (+ :20-47 is_triangle / get_neighbors)
//   anchor_patch_samplers.py:210-243  sample_structure_anchor_patches      -> subgnn_walk_full
//   anchor_patch_samplers.py:118-158  perform_random_walks(inside=True)    -> subgnn_walk_patch(border=0)
//   anchor_patch_samplers.py:118-158  perform_random_walks(inside=False)   -> subgnn_walk_patch(border=1)
//   subgraph_utils.py:126-144         get_border_nodes (no dense N x N adjacency: a patch node is a border
//                                     node iff deg_G(p) > |N(p) ∩ patch|)
//
// Layout: rowptr[N+1], col[2E] int32, 0-indexed, neighbour lists sorted ascending.  Node ids in patches
// and in the output are the reference's 1-indexed ids, 0 = PAD.  One warp owns one walk; the previous
// node's neighbour list is staged in shared memory (coalesced copy) so the per-neighbour triangle test
// (n in N(prev)) is a shared-memory binary search; lanes stride over N(curr) with coalesced loads.
// Random draws: Philox4x32-10, key = seed, counter = (walk id, step, tag): step 0 = start node,
// step 1 = second node, step k+2 = k-th move (word0 -> uniform choice, word1 -> beta coin).
#include "common.cuh"
#include "../../include/subgnn_b200.h"

#define WALK_WARPS 8
#define PREV_CAP 2048   // ints of N(prev) staged per warp (8 KB); longer rows are searched in global/L2

struct PrevRow {
  const int* ptr;   // smem copy if staged, else global row
  int n;
};

__device__ __forceinline__ PrevRow stage_prev(const int* __restrict__ rowptr, const int* __restrict__ col, int prev,
                                              int* sbuf, int lane) {
  int off = rowptr[prev], n = rowptr[prev + 1] - off;
  PrevRow r;
  r.n = n;
  if (n <= PREV_CAP) {
    for (int i = lane; i < n; i += 32) sbuf[i] = col[off + i];
    __syncwarp();
    r.ptr = sbuf;
  } else {
    r.ptr = col + off;
  }
  return r;
}

// select the target-th (0-based) flagged element among items i = 0..n-1 (flag evaluated by f(i, &value))
template <typename F>
__device__ __forceinline__ int warp_select(int n, int target, int lane, F f) {
  int base = 0;
  for (int off = 0; off < n; off += 32) {
    int i = off + lane;
    int value = 0;
    bool flag = (i < n) && f(i, value);
    unsigned bits = __ballot_sync(0xffffffffu, flag);
    int c = __popc(bits);
    if (target < base + c) {
      int src = __fns(bits, 0, target - base + 1);
      return __shfl_sync(0xffffffffu, value, src);
    }
    base += c;
  }
  return -1;
}

// ------------------------------------------------------------------------------------------------
// full-graph walks (structure anchor patch sampling)
__global__ void __launch_bounds__(WALK_WARPS * 32)
walk_full_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int n_nodes, int n_walks, int walk_len,
                 float beta, unsigned long long seed, int* __restrict__ out) {
  extern __shared__ int smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* sprev = smem + warp * PREV_CAP;
  const int warps_total = gridDim.x * WALK_WARPS;
  for (int wid = blockIdx.x * WALK_WARPS + warp; wid < n_walks; wid += warps_total) {
    int* o = out + (size_t)wid * walk_len;
    for (int i = lane; i < walk_len; i += 32) o[i] = 0;
    __syncwarp();
    Philox4 r = sg_draw(seed, (uint64_t)wid, 0, SG_TAG_STRUC);
    int prev = (int)sg_index(r.x, (uint32_t)n_nodes);                 // anchor_patch_samplers.py:70
    if (lane == 0) o[0] = prev + 1;
    int off = rowptr[prev], deg = rowptr[prev + 1] - off;
    if (deg == 0 || walk_len < 2) continue;                           // :74,:83-84
    r = sg_draw(seed, (uint64_t)wid, 1, SG_TAG_STRUC);
    int curr = col[off + sg_index(r.x, (uint32_t)deg)];
    if (lane == 0) o[1] = curr + 1;
    for (int k = 0; k < walk_len - 2; ++k) {                          // :88
      PrevRow pr = stage_prev(rowptr, col, prev, sprev, lane);
      const int coff = rowptr[curr], cdeg = rowptr[curr + 1] - coff;
      if (cdeg == 0) break;                                           // :94
      int n_tri = 0;
      for (int i = lane; i < cdeg; i += 32) n_tri += sorted_contains(pr.ptr, pr.n, col[coff + i]) ? 1 : 0;
      n_tri = warp_sum_i(n_tri);
      const int n_non = cdeg - n_tri;
      r = sg_draw(seed, (uint64_t)wid, (uint32_t)(k + 2), SG_TAG_STRUC);
      bool take_tri;                                                  // :96-106
      if (n_tri == 0) take_tri = false;
      else if (n_non == 0) take_tri = true;
      else take_tri = sg_unit(r.y) <= beta;
      const int target = (int)sg_index(r.x, (uint32_t)(take_tri ? n_tri : n_non));
      int nxt = warp_select(cdeg, target, lane, [&](int i, int& v) {
        v = col[coff + i];
        return sorted_contains(pr.ptr, pr.n, v) == take_tri;
      });
      __syncwarp();
      prev = curr;
      curr = nxt;
      if (lane == 0) o[k + 2] = curr + 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// walks over one anchor patch: inside the induced subgraph, or over its border
// (in-border nodes ∪ nodes outside the patch).  One CTA per patch, one warp per walk.
template <bool BORDER>
__global__ void __launch_bounds__(WALK_WARPS * 32)
walk_patch_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int n_nodes,
                  const int* __restrict__ patches, const int* __restrict__ patch_len, int patch_stride, int n_patches,
                  int n_walks, int walk_len, float beta, unsigned long long seed, int* __restrict__ out) {
  extern __shared__ int smem[];
  int* spatch = smem;                              // [patch_stride] sorted unique 0-indexed nodes
  int* sflag = smem + patch_stride;                // [patch_stride] in-border flag
  int* sprev_all = smem + 2 * patch_stride;        // [WALK_WARPS][PREV_CAP]  (BORDER only)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int p = blockIdx.x; p < n_patches; p += gridDim.x) {
    const int pl = patch_len[p];
    __syncthreads();
    for (int j = threadIdx.x; j < pl; j += blockDim.x) spatch[j] = patches[(size_t)p * patch_stride + j] - 1;
    __syncthreads();
    if (BORDER) {
      // in_border[j] <=> some neighbour of patch[j] lies outside the patch (subgraph_utils.py:141-143)
      for (int j = warp; j < pl; j += WALK_WARPS) {
        const int node = spatch[j];
        const int off = rowptr[node], deg = rowptr[node + 1] - off;
        int inside = 0;
        for (int q = lane; q < pl; q += 32) inside += sorted_contains(col + off, deg, spatch[q]) ? 1 : 0;
        inside = warp_sum_i(inside);
        if (lane == 0) sflag[j] = (inside < deg) ? 1 : 0;
      }
      __syncthreads();
    }
    for (int w = warp; w < n_walks; w += WALK_WARPS) {
      const uint64_t wid = (uint64_t)p * n_walks + w;
      int* o = out + wid * walk_len;
      for (int i = lane; i < walk_len; i += 32) o[i] = 0;
      __syncwarp();
      if (pl == 0) continue;                                          // :134-135 all-PAD patch
      Philox4 r = sg_draw(seed, wid, 0, SG_TAG_WALK);
      int prev, curr;
      if (!BORDER) {
        // ---- inside: every candidate is a patch node adjacent to the current node -------------
        prev = spatch[sg_index(r.x, (uint32_t)pl)];                   // :70
        if (lane == 0) o[0] = prev + 1;
        if (walk_len < 2) continue;
        int off = rowptr[prev], deg = rowptr[prev + 1] - off;
        int cnt = 0;
        for (int q = lane; q < pl; q += 32) cnt += sorted_contains(col + off, deg, spatch[q]) ? 1 : 0;
        cnt = warp_sum_i(cnt);
        if (cnt == 0) continue;                                       // :74
        r = sg_draw(seed, wid, 1, SG_TAG_WALK);
        curr = warp_select(pl, (int)sg_index(r.x, (uint32_t)cnt), lane, [&](int q, int& v) {
          v = spatch[q];
          return sorted_contains(col + off, deg, v);
        });
        if (lane == 0) o[1] = curr + 1;
        for (int k = 0; k < walk_len - 2; ++k) {
          const int coff = rowptr[curr], cdeg = rowptr[curr + 1] - coff;
          const int poff = rowptr[prev], pdeg = rowptr[prev + 1] - poff;
          int n_all = 0, n_tri = 0;
          for (int q = lane; q < pl; q += 32) {
            const int v = spatch[q];
            if (sorted_contains(col + coff, cdeg, v)) {
              ++n_all;
              n_tri += sorted_contains(col + poff, pdeg, v) ? 1 : 0;
            }
          }
          n_all = warp_sum_i(n_all);
          n_tri = warp_sum_i(n_tri);
          if (n_all == 0) break;                                      // :94
          const int n_non = n_all - n_tri;
          r = sg_draw(seed, wid, (uint32_t)(k + 2), SG_TAG_WALK);
          bool take_tri;
          if (n_tri == 0) take_tri = false;
          else if (n_non == 0) take_tri = true;
          else take_tri = sg_unit(r.y) <= beta;
          const int target = (int)sg_index(r.x, (uint32_t)(take_tri ? n_tri : n_non));
          int nxt = warp_select(pl, target, lane, [&](int q, int& v) {
            v = spatch[q];
            return sorted_contains(col + coff, cdeg, v) && (sorted_contains(col + poff, pdeg, v) == take_tri);
          });
          prev = curr;
          curr = nxt;
          if (lane == 0) o[k + 2] = curr + 1;
        }
      } else {
        // ---- border: start on an in-border node, move over N_G(curr) ∩ (in_border ∪ V\patch) ----
        int* sprev = sprev_all + warp * PREV_CAP;
        int nb = 0;
        for (int q = lane; q < pl; q += 32) nb += sflag[q];
        nb = warp_sum_i(nb);
        if (nb == 0) continue;                                        // no border node: empty walk
        prev = warp_select(pl, (int)sg_index(r.x, (uint32_t)nb), lane, [&](int q, int& v) {
          v = spatch[q];
          return sflag[q] != 0;
        });                                                           // :78
        if (lane == 0) o[0] = prev + 1;
        if (walk_len < 2) continue;
        auto valid = [&](int v) {                                     // :79 / :38  n in all_valid_nodes
          int j = sorted_find(spatch, pl, v);
          return j < 0 || sflag[j] != 0;
        };
        int off = rowptr[prev], deg = rowptr[prev + 1] - off;
        int cnt = 0;
        for (int i = lane; i < deg; i += 32) cnt += valid(col[off + i]) ? 1 : 0;
        cnt = warp_sum_i(cnt);
        if (cnt == 0) continue;                                       // :80
        r = sg_draw(seed, wid, 1, SG_TAG_WALK);
        curr = warp_select(deg, (int)sg_index(r.x, (uint32_t)cnt), lane, [&](int i, int& v) {
          v = col[off + i];
          return valid(v);
        });
        if (lane == 0) o[1] = curr + 1;
        for (int k = 0; k < walk_len - 2; ++k) {
          PrevRow pr = stage_prev(rowptr, col, prev, sprev, lane);
          const int coff = rowptr[curr], cdeg = rowptr[curr + 1] - coff;
          int n_all = 0, n_tri = 0;
          for (int i = lane; i < cdeg; i += 32) {
            const int v = col[coff + i];
            if (valid(v)) {
              ++n_all;
              n_tri += sorted_contains(pr.ptr, pr.n, v) ? 1 : 0;
            }
          }
          n_all = warp_sum_i(n_all);
          n_tri = warp_sum_i(n_tri);
          if (n_all == 0) break;
          const int n_non = n_all - n_tri;
          r = sg_draw(seed, wid, (uint32_t)(k + 2), SG_TAG_WALK);
          bool take_tri;
          if (n_tri == 0) take_tri = false;
          else if (n_non == 0) take_tri = true;
          else take_tri = sg_unit(r.y) <= beta;
          const int target = (int)sg_index(r.x, (uint32_t)(take_tri ? n_tri : n_non));
          int nxt = warp_select(cdeg, target, lane, [&](int i, int& v) {
            v = col[coff + i];
            return valid(v) && (sorted_contains(pr.ptr, pr.n, v) == take_tri);
          });
          __syncwarp();
          prev = curr;
          curr = nxt;
          if (lane == 0) o[k + 2] = curr + 1;
        }
      }
    }
  }
}

extern "C" {

int subgnn_walk_full(const int* rowptr, const int* col, int n_nodes, int n_walks, int walk_len, float beta,
                     unsigned long long seed, int* out_walks, void* stream) {
  SG_REQUIRE(n_nodes > 0 && n_walks >= 0 && walk_len >= 1, "bad sizes");
  if (n_walks == 0) return SUBGNN_OK;
  const size_t smem = (size_t)WALK_WARPS * PREV_CAP * sizeof(int);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(walk_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  int grid = sg_grid_for(n_walks, WALK_WARPS, 2);
  walk_full_kernel<<<grid, WALK_WARPS * 32, smem, (cudaStream_t)stream>>>(rowptr, col, n_nodes, n_walks, walk_len, beta, seed, out_walks);
  return subgnn_check_launch("walk_full_kernel");
}

int subgnn_walk_patch(const int* rowptr, const int* col, int n_nodes, const int* patches, const int* patch_len,
                      int patch_stride, int n_patches, int n_walks, int walk_len, float beta, int border,
                      unsigned long long seed, int* out_walks, void* stream) {
  SG_REQUIRE(n_nodes > 0 && n_patches >= 0 && n_walks >= 1 && walk_len >= 1 && patch_stride >= 1, "bad sizes");
  if (n_patches == 0) return SUBGNN_OK;
  const size_t smem = ((size_t)2 * patch_stride + (border ? (size_t)WALK_WARPS * PREV_CAP : 0)) * sizeof(int);
  SG_REQUIRE(smem <= 200 * 1024, "patch too long for shared memory staging");
  int grid = sg_grid_for(n_patches, 1, 2);
  if (border) {
    cudaFuncSetAttribute(walk_patch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    walk_patch_kernel<true><<<grid, WALK_WARPS * 32, smem, (cudaStream_t)stream>>>(rowptr, col, n_nodes, patches, patch_len, patch_stride,
                                                                                 n_patches, n_walks, walk_len, beta, seed, out_walks);
  } else {
    cudaFuncSetAttribute(walk_patch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    walk_patch_kernel<false><<<grid, WALK_WARPS * 32, smem, (cudaStream_t)stream>>>(rowptr, col, n_nodes, patches, patch_len, patch_stride,
                                                                                  n_patches, n_walks, walk_len, beta, seed, out_walks);
  }
  return subgnn_check_launch("walk_patch_kernel");
}

}  // extern "C"
