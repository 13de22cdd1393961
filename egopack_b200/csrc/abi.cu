// Library-level entry points: version, per-thread error string, device info.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace egp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return dev;
}

int sm_count() {
  static int cached[kMaxDevices] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("EGP_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

}  // namespace egp

extern "C" {

int egp_version(void) { return 2; }

int egp_last_error(char* buf, size_t len) {
  if (!buf || len == 0) return EGP_ERR_INVALID;
  strncpy(buf, egp::g_err, len - 1);
  buf[len - 1] = '\0';
  return EGP_OK;
}

int egp_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  EGP_CUDA(cudaGetDevice(&dev));
  int sms = 0, maj = 0, min = 0;
  EGP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  EGP_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  EGP_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return EGP_OK;
}

}  // extern "C"
