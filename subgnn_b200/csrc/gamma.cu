// gamma: position/neighbourhood similarity (shortest-path min) and structure similarity (degree
// sequences + DTW).
//
// Replaces (reference):
//   SubGNN.py:752-781   compute_shortest_path_similarities  -> subgnn_sp_min_dense / subgnn_sp_min_gather
//   gamma.py:21-49      get_degree_sequence                 -> subgnn_degree_seq        (integer, bit-exact)
//   gamma.py:51-59      calc_dist / calc_dtw (fastdtw, radius 1, fp64) -> subgnn_dtw_batch
//   SubGNN.py:783-833   compute_structure_patch_similarities (all (cc, patch) pairs)    -> subgnn_dtw_batch
//
// DTW arithmetic is fp64 on purpose: the reference computes in Python floats and the fastdtw window of
// every level depends on first-minimum tie-breaks, which are pervasive on integer degree sequences; the
// same operation order in fp64 reproduces the oracle bit for bit, the fp32 cast happens once at the end
// (SubGNN.py:822).  fastdtw windows use the closed form of fastdtw's __expand_window for radius 1: fine row r
// covers columns [2(jmin-1), 2 jmax+3] over the coarse path cells of coarse rows (r>>1)-1 .. (r>>1)+1.
#include "common.cuh"
#include "../../include/subgnn_b200.h"

// ------------------------------------------------------------------------------------------------
// SP-min, streaming form: out[row][:] = min_{u in cc(row)} hop[u][:]  (uint8 table -> fp32 row)
__global__ void sp_min_dense_kernel(const uint8_t* __restrict__ hop, int n_nodes, size_t hop_stride,
                                    const int* __restrict__ cc_ptr, const int* __restrict__ cc_nodes, int n_rows,
                                    float* __restrict__ out) {
  // each thread owns 16 consecutive columns (one 128-bit load per component node)
  const int chunks = (n_nodes + 15) / 16;
  const long long total = (long long)n_rows * chunks;
  const bool vec_ok = (hop_stride % 16) == 0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(t / chunks), ch = (int)(t % chunks);
    const int c0 = ch * 16;
    const int beg = cc_ptr[row], end = cc_ptr[row + 1];
    float* o = out + (size_t)row * n_nodes + c0;
    if (beg == end) {
      for (int k = 0; k < 16 && c0 + k < n_nodes; ++k) o[k] = 0.f;   // padded component (SubGNN.py:778)
      continue;
    }
    if (vec_ok && c0 + 16 <= n_nodes) {
      uint4 m = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      for (int i = beg; i < end; ++i) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(hop + (size_t)(cc_nodes[i] - 1) * hop_stride + c0));
        m.x = __vminu4(m.x, v.x); m.y = __vminu4(m.y, v.y); m.z = __vminu4(m.z, v.z); m.w = __vminu4(m.w, v.w);
      }
      const unsigned w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 f = make_float4((float)(w[q] & 0xff), (float)((w[q] >> 8) & 0xff), (float)((w[q] >> 16) & 0xff), (float)(w[q] >> 24));
        if (((size_t)o & 15) == 0) *reinterpret_cast<float4*>(o + 4 * q) = f;
        else { o[4 * q] = f.x; o[4 * q + 1] = f.y; o[4 * q + 2] = f.z; o[4 * q + 3] = f.w; }
      }
    } else {
      for (int k = 0; k < 16 && c0 + k < n_nodes; ++k) {
        unsigned m = 255;
        for (int i = beg; i < end; ++i) m = min(m, (unsigned)hop[(size_t)(cc_nodes[i] - 1) * hop_stride + c0 + k]);
        o[k] = (float)m;
      }
    }
  }
}

// SP-min, gather form: out[row][a] = min_{u in cc(row)} hop[u][anchor(row,a)-1], 0 for PAD anchors.
// anchor(row,a) = anchors[anchor_row(row)*A + a]; anchor_row maps a component to its anchor list
// (identity for N anchors, the owning subgraph for P-internal, 0 for the shared P-border list).
__global__ void sp_min_gather_kernel(const uint8_t* __restrict__ hop, size_t hop_stride, const int* __restrict__ cc_ptr,
                                     const int* __restrict__ cc_nodes, int n_rows, const int* __restrict__ anchors,
                                     const int* __restrict__ anchor_row, int A, float* __restrict__ out) {
  const long long total = (long long)n_rows * A;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(t / A), a = (int)(t % A);
    const int ar = anchor_row ? anchor_row[row] : row;
    const int id = anchors[(size_t)ar * A + a];
    const int beg = cc_ptr[row], end = cc_ptr[row + 1];
    float r = 0.f;
    if (id != 0 && beg < end) {
      unsigned m = 255;
      for (int i = beg; i < end; ++i) m = min(m, (unsigned)__ldg(hop + (size_t)(cc_nodes[i] - 1) * hop_stride + (id - 1)));
      r = (float)m;
    }
    out[t] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// degree sequences: one warp per row of node ids (PAD = 0 anywhere, duplicates kept — SURVEY F9)
#define DEG_WARPS 4
__global__ void __launch_bounds__(DEG_WARPS * 32)
degree_seq_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ rows, int n_rows,
                  int stride, int internal, int* __restrict__ out_seq, int* __restrict__ out_len) {
  extern __shared__ int smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* nodes = smem + warp * 3 * stride;
  int* first = nodes + stride;
  int* vals = first + stride;
  for (int row = blockIdx.x * DEG_WARPS + warp; row < n_rows; row += gridDim.x * DEG_WARPS) {
    const int* src = rows + (size_t)row * stride;
    int n = 0;
    for (int off = 0; off < stride; off += 32) {                 // gamma.py:27 drop PAD, keep order
      const int i = off + lane;
      const int id = i < stride ? src[i] : 0;
      const unsigned bits = __ballot_sync(0xffffffffu, id != 0);
      if (id != 0) nodes[n + __popc(bits & ((1u << lane) - 1))] = id - 1;
      n += __popc(bits);
    }
    __syncwarp();
    for (int j = lane; j < n; j += 32) {                         // gamma.py:29 subgraph on the unique ids
      int f = 1;
      for (int q = 0; q < j; ++q) f &= (nodes[q] != nodes[j]);
      first[j] = f;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {                         // gamma.py:30 degree inside the patch
      const int node = nodes[i];
      const int off = rowptr[node], deg = rowptr[node + 1] - off;
      int cnt = 0;
      for (int j = 0; j < n; ++j)
        if (first[j] && sorted_contains(col + off, deg, nodes[j])) cnt += (nodes[j] == node) ? 2 : 1;  // self loop counts twice (networkx)
      int full = deg + (sorted_contains(col + off, deg, node) ? 1 : 0);
      vals[i] = internal ? cnt : full - cnt;                     // gamma.py:33-49
    }
    __syncwarp();
    int* dst = out_seq + (size_t)row * stride;
    for (int i = lane; i < stride; i += 32) dst[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {                         // stable rank sort (ascending)
      const int v = vals[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += (vals[j] < v) || (vals[j] == v && j < i);
      dst[rank] = v;
    }
    if (lane == 0) out_len[row] = n;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// DTW: one thread per (a, b) pair, per-thread workspace in shared memory
struct DtwWs {
  int seq_x, seq_y, rows, par, jmn, jmx, total;   // word offsets (4-byte words) inside a thread's workspace
};

__host__ __device__ inline DtwWs dtw_layout(int LA, int LB) {
  DtwWs w;
  int o = 0;
  w.rows = o; o += 2 * 2 * (LB + 1);                     // two rolling rows of doubles (first => 8-byte aligned)
  w.seq_x = o; o += 2 * LA + 2;                          // all coarsening levels of x as floats (exact)
  w.seq_y = o; o += 2 * LB + 2;
  w.par = o; o += (((LA / 2 + 1) * (LB / 2 + 1) + 3) / 4 + 3) / 4 + 1;   // 2-bit parents for levels >= 1
  w.jmn = o; o += (LA / 2 + 2 + 1) / 2;                  // int16 per coarse row
  w.jmx = o; o += (LA / 2 + 2 + 1) / 2;
  o += (o & 1);                                          // keep every thread's base 8-byte aligned ...
  if (((o / 2) & 1) == 0) o += 2;                        // ... and the 8-byte stride odd (bank spread)
  w.total = o;
  return w;
}

__device__ __forceinline__ double dtw_cost(float a, float b) {      // gamma.py:51-52 in fp64
  const double x = (double)a, y = (double)b;
  const double mx = x > y ? x : y, mn = x > y ? y : x;
  return (mx + 1.0) / (mn + 1.0) - 1.0;
}

__global__ void dtw_batch_kernel(const int* __restrict__ seqA, const int* __restrict__ lenA, const int* __restrict__ rowsA, int nA, int strideA,
                                 const int* __restrict__ seqB, const int* __restrict__ lenB, int nB, int strideB,
                                 int LA, int LB, int fast, float* __restrict__ out) {
  extern __shared__ __align__(16) int smem[];
  const DtwWs L = dtw_layout(LA, LB);
  int* ws = smem + (size_t)threadIdx.x * L.total;
  double* rows = reinterpret_cast<double*>(ws + L.rows);
  float* xs = reinterpret_cast<float*>(ws + L.seq_x);
  float* ys = reinterpret_cast<float*>(ws + L.seq_y);
  unsigned char* par = reinterpret_cast<unsigned char*>(ws + L.par);
  short* jmn = reinterpret_cast<short*>(ws + L.jmn);
  short* jmx = reinterpret_cast<short*>(ws + L.jmx);
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  const long long total = (long long)nA * nB;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int ai = (int)(p / nB), b = (int)(p % nB);
    const int a = rowsA ? rowsA[ai] : ai;                           // optional row subset (length buckets)
    float* outp = out + (size_t)a * nB + b;
    const int n0 = lenA[a], m0 = lenB[b];
    if (n0 == 0 || m0 == 0) { *outp = 0.f; continue; }            // SubGNN.py:831
    for (int i = 0; i < n0; ++i) xs[i] = (float)seqA[(size_t)a * strideA + i];
    for (int j = 0; j < m0; ++j) ys[j] = (float)seqB[(size_t)b * strideB + j];
    // coarsening levels (fastdtw __reduce_by_half); level l of x starts at xoff[l]
    int nlev = 0, xoff[17], yoff[17], xn[17], yn[17];
    xoff[0] = 0; yoff[0] = 0; xn[0] = n0; yn[0] = m0;
    if (fast) {
      while (xn[nlev] >= 3 && yn[nlev] >= 3 && nlev < 16) {
        const int n = xn[nlev], m = yn[nlev];
        const float* sx = xs + xoff[nlev];
        const float* sy = ys + yoff[nlev];
        float* dx = xs + xoff[nlev] + n;
        float* dy = ys + yoff[nlev] + m;
        for (int i = 0; i < n / 2; ++i) dx[i] = (sx[2 * i] + sx[2 * i + 1]) * 0.5f;   // exact in fp32 for degree data
        for (int j = 0; j < m / 2; ++j) dy[j] = (sy[2 * j] + sy[2 * j + 1]) * 0.5f;
        xoff[nlev + 1] = xoff[nlev] + n; yoff[nlev + 1] = yoff[nlev] + m;
        xn[nlev + 1] = n / 2; yn[nlev + 1] = m / 2;
        ++nlev;
      }
    }
    double dist = 0.0;
    for (int lev = nlev; lev >= 0; --lev) {
      const int n = xn[lev], m = yn[lev];
      const float* sx = xs + xoff[lev];
      const float* sy = ys + yoff[lev];
      const bool windowed = lev < nlev;
      const int nc = windowed ? xn[lev + 1] : 0;
      double* prev = rows;
      double* cur = rows + (LB + 1);
      int plo = 0, phi = 0;            // valid index range of prev (D column index = j+1)
      prev[0] = 0.0;
      for (int r = 0; r < n; ++r) {
        int lo = 0, hi = m - 1;
        if (windowed) {
          const int ci = r >> 1;
          const int c0 = ci - 1 < 0 ? 0 : ci - 1, c1 = ci + 1 > nc - 1 ? nc - 1 : ci + 1;
          int mn = jmn[c0], mx = jmx[c0];
          for (int c = c0 + 1; c <= c1; ++c) { mn = min(mn, (int)jmn[c]); mx = max(mx, (int)jmx[c]); }
          lo = max(0, 2 * (mn - 1));
          hi = min(m - 1, 2 * mx + 3);
        }
        const float xv = sx[r];
        double left = INF;             // D[r][lo-1] is outside the window
        for (int j = lo; j <= hi; ++j) {
          const double dt = dtw_cost(xv, sy[j]);
          const double up = (j + 1 >= plo && j + 1 <= phi) ? prev[j + 1] : INF;
          const double dg = (j >= plo && j <= phi) ? prev[j] : INF;
          double best = up + dt; int pp = 0;                        // first minimum in the order up, left, diag
          double c = left + dt;
          if (c < best) { best = c; pp = 1; }
          c = dg + dt;
          if (c < best) { best = c; pp = 2; }
          cur[j + 1] = best;
          left = best;
          if (lev > 0) {
            const int cell = r * m + j;
            const unsigned char sh = (cell & 3) * 2;
            par[cell >> 2] = (par[cell >> 2] & ~(3 << sh)) | (pp << sh);
          }
        }
        plo = lo + 1; phi = hi + 1;
        double* t = prev; prev = cur; cur = t;
      }
      dist = (m >= plo && m <= phi) ? prev[m] : INF;
      if (lev > 0) {                   // backtrace -> per-row column range of the warp path
        for (int i = 0; i < n; ++i) { jmn[i] = 32767; jmx[i] = -1; }
        int i = n - 1, j = m - 1;
        while (i >= 0 && j >= 0) {
          if (j < jmn[i]) jmn[i] = (short)j;
          if (j > jmx[i]) jmx[i] = (short)j;
          const int cell = i * m + j;
          const int pp = (par[cell >> 2] >> ((cell & 3) * 2)) & 3;
          if (pp == 0) --i; else if (pp == 1) --j; else { --i; --j; }
        }
      }
    }
    *outp = (float)(1.0 / (dist + 1.0));                           // gamma.py:59, cast SubGNN.py:822
  }
}

// ------------------------------------------------------------------------------------------------
// Exact DTW, wavefront mapping: a group of G lanes (G = 4 ... 32, 32/G pairs per warp) owns one (a, b) pair.  Lane l of the
// group holds rows l*R ... l*R+R-1 of the DP (the component sequence, the short side) in registers and walks the columns of
// the patch sequence one per iteration, one anti-diagonal step behind lane l-1: D[i-1][j] arrives by __shfl_up_sync, D[i-1][j-1]
// is the value received one iteration earlier, D[i][j-1] is the lane's own previous result.  No DP state in shared memory (only
// the patch sequence is staged there).  Values are identical to the row-by-row evaluation: every cell is cost + min of the same
// three fp64 numbers, and min(a+c, b+c, d+c) == min(a, b, d) + c in floating point (rounding is monotonic).
template <int G, int R>
__global__ void __launch_bounds__(128) dtw_exact_wave_kernel(const int* __restrict__ seqA, const int* __restrict__ lenA, const int* __restrict__ rowsA, int nA, int strideA,
                                                             const int* __restrict__ seqB, const int* __restrict__ lenB, int nB, int strideB,
                                                             int LB, float* __restrict__ out) {
  extern __shared__ __align__(16) int smem[];
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;                                  // lane inside the group
  const int gw = lane / G;                                  // group inside the warp
  const int groups_per_cta = blockDim.x / G;
  int* ys = smem + (threadIdx.x / G) * LB;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  const long long total = (long long)nA * nB;
  const long long stride = (long long)gridDim.x * groups_per_cta;
  // all groups of a warp run the same number of rounds (shuffles use the full mask)
  for (long long base = (long long)blockIdx.x * groups_per_cta + threadIdx.x / G - gw; base < total; base += stride) {
    const long long p = base + gw;
    const bool valid = p < total;
    int n = 0, m = 0, a = 0, b = 0;
    if (valid) {
      a = (int)(p / nB); b = (int)(p % nB);
      if (rowsA) a = rowsA[a];                              // optional row subset (length buckets)
      n = lenA[a]; m = lenB[b];
      if (n == 0 || m == 0) {                               // SubGNN.py:831
        if (gl == 0) out[(size_t)a * nB + b] = 0.f;
        n = 0; m = 0;
      }
    }
    for (int j = gl; j < m; j += G) ys[j] = seqB[(size_t)b * strideB + j];
    double x[R], col[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = gl * R + r;
      x[r] = row < n ? (double)seqA[(size_t)a * strideA + row] : 0.0;
      col[r] = INF;                                          // D[row][-1]
    }
    __syncwarp();
    const int iters = __reduce_max_sync(0xffffffffu, m > 0 ? m + (n - 1) / R : 0);
    double last = INF, diag_in = INF;
    for (int t = 0; t < iters; ++t) {
      const double recv = __shfl_up_sync(0xffffffffu, last, 1, G);      // D[l*R-1][t-l], produced one iteration ago
      const int j = t - gl;
      if (j >= 0 && j < m) {
        double up = gl == 0 ? INF : recv;
        double dg = gl == 0 ? (j == 0 ? 0.0 : INF) : diag_in;
        const double y = (double)ys[j];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const double left = col[r];
          const double mx = x[r] > y ? x[r] : y, mn = x[r] > y ? y : x[r];
          const double best = fmin(fmin(up, left), dg) + ((mx + 1.0) / (mn + 1.0) - 1.0);    // gamma.py:51-52
          dg = left; up = best; col[r] = best;
        }
        last = up;
      }
      diag_in = recv;
    }
    if (n > 0 && gl == (n - 1) / R) {
      double d = col[0];
#pragma unroll
      for (int r = 1; r < R; ++r) if (r == (n - 1) % R) d = col[r];
      out[(size_t)a * nB + b] = (float)(1.0 / (d + 1.0));    // gamma.py:59, cast SubGNN.py:822
    }
    __syncwarp();
  }
}

template <int G, int R>
static int launch_dtw_wave(const int* seqA, const int* lenA, const int* rowsA, int nA, int strideA, const int* seqB, const int* lenB, int nB,
                           int strideB, int LB, float* out, cudaStream_t st) {
  const int threads = 128, groups = threads / G;
  const size_t smem = (size_t)groups * LB * sizeof(int);
  if (smem > 200 * 1024) return -1;
  cudaFuncSetAttribute(dtw_exact_wave_kernel<G, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const long long total = (long long)nA * nB;
  const long long want = (total + groups - 1) / groups;
  const int grid = (int)(want < 148LL * 16 ? want : 148LL * 16);
  dtw_exact_wave_kernel<G, R><<<grid, threads, smem, st>>>(seqA, lenA, rowsA, nA, strideA, seqB, lenB, nB, strideB, LB, out);
  return 0;
}

extern "C" {

int subgnn_sp_min_dense(const unsigned char* hop, int n_nodes, long long hop_stride, const int* cc_ptr, const int* cc_nodes,
                        int n_rows, float* out, void* stream) {
  SG_REQUIRE(n_nodes > 0 && n_rows >= 0 && hop_stride >= n_nodes, "bad sizes");
  if (n_rows == 0) return SUBGNN_OK;
  const long long work = (long long)n_rows * ((n_nodes + 15) / 16);
  sp_min_dense_kernel<<<sg_grid_for(work, 256, 8), 256, 0, (cudaStream_t)stream>>>(hop, n_nodes, (size_t)hop_stride, cc_ptr, cc_nodes, n_rows, out);
  return subgnn_check_launch("sp_min_dense_kernel");
}

int subgnn_sp_min_gather(const unsigned char* hop, long long hop_stride, const int* cc_ptr, const int* cc_nodes, int n_rows,
                         const int* anchors, const int* anchor_row, int n_anchors, float* out, void* stream) {
  SG_REQUIRE(n_rows >= 0 && n_anchors >= 1, "bad sizes");
  if (n_rows == 0) return SUBGNN_OK;
  sp_min_gather_kernel<<<sg_grid_for((long long)n_rows * n_anchors, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      hop, (size_t)hop_stride, cc_ptr, cc_nodes, n_rows, anchors, anchor_row, n_anchors, out);
  return subgnn_check_launch("sp_min_gather_kernel");
}

int subgnn_degree_seq(const int* rowptr, const int* col, const int* rows, int n_rows, int stride, int internal, int* out_seq,
                      int* out_len, void* stream) {
  SG_REQUIRE(n_rows >= 0 && stride >= 1, "bad sizes");
  if (n_rows == 0) return SUBGNN_OK;
  const size_t smem = (size_t)DEG_WARPS * 3 * stride * sizeof(int);
  SG_REQUIRE(smem <= 200 * 1024, "row too long");
  cudaFuncSetAttribute(degree_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  degree_seq_kernel<<<sg_grid_for(n_rows, DEG_WARPS, 8), DEG_WARPS * 32, smem, (cudaStream_t)stream>>>(rowptr, col, rows, n_rows, stride,
                                                                                                  internal, out_seq, out_len);
  return subgnn_check_launch("degree_seq_kernel");
}

int subgnn_dtw_batch_rows(const int* seqA, const int* lenA, const int* rowsA, int n_rows, int strideA, const int* seqB, const int* lenB,
                          int nB, int strideB, int max_len_a, int max_len_b, int mode, float* out, void* stream) {
  const int nA = n_rows;                       // pairs (rowsA[i], b), i < n_rows; out is indexed by the ORIGINAL row
  SG_REQUIRE(nA >= 0 && nB >= 0 && max_len_a >= 1 && max_len_b >= 1, "bad sizes");
  SG_REQUIRE(mode == SUBGNN_DTW_EXACT || mode == SUBGNN_DTW_FASTDTW_R1 || mode == SUBGNN_DTW_EXACT_THREAD, "unknown DTW mode");
  SG_REQUIRE(max_len_a <= strideA && max_len_b <= strideB && max_len_a < 32000 && max_len_b < 32000, "bad max lengths");
  if ((long long)nA * nB == 0) return SUBGNN_OK;
  if (mode == SUBGNN_DTW_EXACT && max_len_a <= 256) {         // lanes own the rows of the component sequence
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
#define SG_WAVE(G, R) rc = launch_dtw_wave<G, R>(seqA, lenA, rowsA, nA, strideA, seqB, lenB, nB, strideB, max_len_b, out, st)
    if (max_len_a <= 4) SG_WAVE(4, 1);
    else if (max_len_a <= 8) SG_WAVE(8, 1);
    else if (max_len_a <= 16) SG_WAVE(16, 1);
    else if (max_len_a <= 32) SG_WAVE(32, 1);
    else if (max_len_a <= 64) SG_WAVE(32, 2);
    else if (max_len_a <= 128) SG_WAVE(32, 4);
    else SG_WAVE(32, 8);
#undef SG_WAVE
    if (rc == 0) return subgnn_check_launch("dtw_exact_wave_kernel");
  }
  const DtwWs L = dtw_layout(max_len_a, max_len_b);
  const size_t per_thread = (size_t)L.total * 4;
  int threads = (int)((200 * 1024) / per_thread) / 32 * 32;
  if (threads > 128) threads = 128;
  SG_REQUIRE(threads >= 32, "sequences too long for the shared-memory DTW workspace");
  const size_t smem = per_thread * threads;
  cudaFuncSetAttribute(dtw_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = sg_grid_for((long long)nA * nB, threads, 1);
  dtw_batch_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(seqA, lenA, rowsA, nA, strideA, seqB, lenB, nB, strideB, max_len_a, max_len_b,
                                                                  mode == SUBGNN_DTW_FASTDTW_R1, out);
  return subgnn_check_launch("dtw_batch_kernel");
}

int subgnn_dtw_batch(const int* seqA, const int* lenA, int nA, int strideA, const int* seqB, const int* lenB, int nB, int strideB,
                     int max_len_a, int max_len_b, int mode, float* out, void* stream) {
  return subgnn_dtw_batch_rows(seqA, lenA, nullptr, nA, strideA, seqB, lenB, nB, strideB, max_len_a, max_len_b, mode, out, stream);
}

}  // extern "C"
