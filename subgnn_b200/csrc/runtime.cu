// Error reporting / device queries for the C-ABI.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "../../include/subgnn_b200.h"

static thread_local char g_err[512] = "";

void subgnn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// default classes launched with programmatic stream serialization (see sg_pdl_sync in common.cuh).  Measured on B200, ms/step
// (ppi_bp / hpo_metab / em_user shapes, same box): mask 0: 0.445 / 0.860 / 0.712; 4 (recurrences): 0.442 / 0.843 / 0.727;
// 14 (GEMMs + recurrences + row kernels): 0.431 / 0.800 / 0.739; 15 (everything): 0.426 / 0.879 / 0.827 — early-scheduled
// small kernels take SM slots from the other graph branches at the two H = 128 shapes.  Bit 4 (SG_PDL_CHAIN: the small kernels ON
// the critical chain — readout MLP, structure q, LSTM head, optimizer) added later, same three shapes, mask 14 -> 30:
// 0.3896 -> 0.3833 / 0.7546 -> 0.7473 / 0.659 -> 0.653.
#define SUBGNN_PDL_DEFAULT_MASK 30

static unsigned long long g_launches = 0;

// kernel-instantiation log: the dispatchers note which template instantiation they launched ("row_fwd<2>", "lstm_fwd_tile<1,4,64>",
// "tc_gemm<fwd,128>" ...) so that the parity tests can assert that they exercised the instantiations the benchmark shapes run
#define SG_VARIANT_MAX 64
static char g_variants[SG_VARIANT_MAX][48];
static int g_n_variants = 0;

void subgnn_note_variant(const char* fmt, ...) {
  char buf[48];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  for (int i = 0; i < g_n_variants; ++i)
    if (strcmp(g_variants[i], buf) == 0) return;
  if (g_n_variants < SG_VARIANT_MAX) strcpy(g_variants[g_n_variants++], buf);
}

int subgnn_check_launch(const char* what) {
  ++g_launches;   // one call per kernel launch (also while a CUDA graph is being captured)
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    subgnn_set_error("%s: %s", what, cudaGetErrorString(e));
    return SUBGNN_ERR_CUDA;
  }
  return SUBGNN_OK;
}

int subgnn_pdl_enabled(int kernel_class) {
  static int mask = -1;
  if (mask < 0) {
    const char* e = getenv("SUBGNN_B200_PDL");
    mask = e ? atoi(e) : SUBGNN_PDL_DEFAULT_MASK;
  }
  return (mask >> kernel_class) & 1;
}

int subgnn_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;  // B200
  }
  return sms;
}

extern "C" {
const char* subgnn_last_error(void) { return g_err; }
int subgnn_abi_version(void) { return SUBGNN_ABI_VERSION; }
int subgnn_device_sm_count(void) { return subgnn_sm_count(); }
unsigned long long subgnn_launch_count(void) { return g_launches; }
int subgnn_variant_log(char* buf, int cap) {
  int n = 0;
  if (cap > 0) buf[0] = 0;
  for (int i = 0; i < g_n_variants; ++i) {
    const int len = (int)strlen(g_variants[i]);
    if (n + len + 2 > cap) break;
    memcpy(buf + n, g_variants[i], len);
    n += len;
    buf[n++] = ';';
    buf[n] = 0;
  }
  return g_n_variants;
}
void subgnn_variant_log_reset(void) { g_n_variants = 0; }
}
