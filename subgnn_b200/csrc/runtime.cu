// Error reporting / device queries for the C-ABI.
#include <stdarg.h>
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

static unsigned long long g_launches = 0;

int subgnn_check_launch(const char* what) {
  ++g_launches;   // one call per kernel launch (also while a CUDA graph is being captured)
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    subgnn_set_error("%s: %s", what, cudaGetErrorString(e));
    return SUBGNN_ERR_CUDA;
  }
  return SUBGNN_OK;
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
}
