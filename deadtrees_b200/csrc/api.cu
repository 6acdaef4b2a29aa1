// Error plumbing, device check and version for the deadtrees_b200 C-ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_last_error[512] = "";

void dt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int dt_check_device() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = DT_ERR_CUDA;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    dt_set_error("no CUDA device: %s (deadtrees_b200 has no CPU fallback)", cudaGetErrorString(e));
    cudaGetLastError();
    return DT_ERR_CUDA;
  }
  if (dev == cached_dev) return cached_rc;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cached_dev = dev;
  if (major != 10) {
    dt_set_error("device %d is sm_%d%d; deadtrees_b200 kernels are built for sm_100a only", dev, major, minor);
    cached_rc = DT_ERR_UNSUPPORTED_ARCH;
  } else {
    cached_rc = DT_OK;
  }
  return cached_rc;
}

extern "C" {

int dt_version(void) { return DT_VERSION; }

int dt_last_error(char* buf, size_t n) {
  if (buf && n) {
    strncpy(buf, g_last_error, n - 1);
    buf[n - 1] = 0;
  }
  return static_cast<int>(strlen(g_last_error));
}

int dt_device_check(void) { return dt_check_device(); }

}  // extern "C"
