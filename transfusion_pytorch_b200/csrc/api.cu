// Library-level plumbing of the C ABI: error string, device query, version.
#include "common.cuh"
#include "../../include/tfx_b200.h"
#include <stdarg.h>
#include <stdio.h>

namespace tfx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return -3;
}

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace tfx

extern "C" {

const char* tfx_last_error(void) { return tfx::g_err; }

int tfx_version(void) { return TFX_B200_VERSION; }

int tfx_init(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { tfx::set_error("tfx_init: no CUDA device visible"); return -1; }
  if (device < 0 || device >= n) { tfx::set_error("tfx_init: device %d out of range (%d visible)", device, n); return -1; }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (major != 10) { tfx::set_error("tfx_init: device %d is sm_%d%d; this library contains sm_100a code only", device, major, minor); return -1; }
  return 0;
}

}
