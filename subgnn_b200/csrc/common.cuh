// Common device helpers for libsubgnn_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define SUBGNN_OK 0
#define SUBGNN_ERR_ARG -1
#define SUBGNN_ERR_CUDA -2

void subgnn_set_error(const char* fmt, ...);
int subgnn_check_launch(const char* what);
void subgnn_note_variant(const char* fmt, ...);

#define SG_REQUIRE(cond, msg)                                  \
  do {                                                         \
    if (!(cond)) {                                             \
      subgnn_set_error("%s: %s", __func__, msg);               \
      return SUBGNN_ERR_ARG;                                   \
    }                                                          \
  } while (0)

static inline int sg_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// grid sizing: persistent-style grid, a multiple of the SM count (148 on B200)
int subgnn_sm_count();
static inline int sg_grid_for(long long work_items, int per_block, int max_blocks_per_sm = 8) {
  long long need = (work_items + per_block - 1) / per_block;
  long long cap = (long long)subgnn_sm_count() * max_blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The training step is a chain of ~25 dependent launches of a few microseconds each, so
// the launch / scheduling gap between two of them is a visible share of the step.  Every kernel of the step starts with
// sg_pdl_sync(): griddepcontrol.wait returns once ALL prerequisite grids have completed and flushed (full dependency, so the
// kernel body is unchanged), then griddepcontrol.launch_dependents lets the next launch in stream order be scheduled while this
// grid runs.  Kernels are launched through sg_launch_pdl, which sets cudaLaunchAttributeProgrammaticStreamSerialization (also
// honoured by stream capture: the step graph gets programmatic edges).  Triggering only after the wait keeps completion
// transitive along a stream (C waits for B, B passed its wait only after A completed).  SUBGNN_B200_PDL=0 disables the
// attribute; the two instructions are no-ops for a normally launched kernel.
__device__ __forceinline__ void sg_pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// Producers of DERIVED WEIGHTS (permuted / transposed copies that later kernels stage before their own wait) never trigger
// early: their dependents are scheduled at grid completion, so a consumer's pre-wait prologue cannot overlap the producer.
__device__ __forceinline__ void sg_pdl_wait_only() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// kernel classes for the SUBGNN_B200_PDL bit mask (bit = class): small element-wise / per-patch kernels, tensor-core GEMMs,
// LSTM recurrences, row (component) kernels
enum { SG_PDL_SMALL = 0, SG_PDL_GEMM = 1, SG_PDL_RECUR = 2, SG_PDL_ROW = 3, SG_PDL_CHAIN = 4 };   // CHAIN: small kernels ON the step's critical chain
int subgnn_pdl_enabled(int kernel_class);
template <int CLASS = SG_PDL_SMALL, class... KArgs, class... Args>
static inline void sg_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = subgnn_pdl_enabled(CLASS);
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based); identical construction in oracle/rng.py
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

#define SG_TAG_WALK 0x57414C4Bu
#define SG_TAG_NEIGH 0x4E454947u
#define SG_TAG_POS 0x504F5331u
#define SG_TAG_STRUC 0x53545231u
#define SG_TAG_DROP 0x44524F50u

__host__ __device__ __forceinline__ Philox4 sg_draw(uint64_t seed, uint64_t item, uint32_t step, uint32_t tag) {
  return philox4x32_10((uint32_t)item, (uint32_t)(item >> 32), step, tag, (uint32_t)seed, (uint32_t)(seed >> 32));
}
__host__ __device__ __forceinline__ uint32_t sg_index(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }
__host__ __device__ __forceinline__ float sg_unit(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

// keep-mask for dropout: element idx of stream (seed, salt); returns scale (0 or 1/(1-p))
__device__ __forceinline__ float sg_dropout_scale(uint64_t seed, uint32_t salt, uint64_t idx, float p) {
  if (p <= 0.f) return 1.f;
  Philox4 r = sg_draw(seed, idx >> 2, salt, SG_TAG_DROP);
  uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return sg_unit(w) >= p ? 1.f / (1.f - p) : 0.f;
}

// cooperative global -> shared staging with U independent loads in flight per thread (a plain load->store loop serialises on
// the L2 latency: the store of iteration i needs its load before the load of iteration i+1 issues)
template <int U, class Src>
__device__ __forceinline__ void sg_stage(float* dst, int n, Src src) {
  for (int e0 = threadIdx.x; e0 < n; e0 += U * (int)blockDim.x) {
    float r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * (int)blockDim.x;
      r[u] = e < n ? src(e) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * (int)blockDim.x;
      if (e < n) dst[e] = r[u];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// thread-block clusters: rank, barrier, store into a peer CTA's shared memory (distributed shared memory)
__device__ __forceinline__ unsigned sg_cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void sg_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void sg_st_cluster(const float* local_addr, unsigned rank, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(local_addr)), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(remote), "f"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// lower_bound membership test in a sorted int list
__device__ __forceinline__ bool sorted_contains(const int* __restrict__ a, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    int v = a[mid];
    if (v < key) lo = mid + 1; else hi = mid;
  }
  return lo < n && a[lo] == key;
}
__device__ __forceinline__ int sorted_find(const int* __restrict__ a, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    int v = a[mid];
    if (v < key) lo = mid + 1; else hi = mid;
  }
  return (lo < n && a[lo] == key) ? lo : -1;
}
