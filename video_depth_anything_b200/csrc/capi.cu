// Library-level entry points of libvda (error string, version, device query).
#include <stdarg.h>
#include <stdio.h>

#include "../../include/vda.h"
#include "common.cuh"

namespace vda {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace vda

extern "C" int vda_version(void) { return 100; }
extern "C" const char* vda_last_error(void) { return vda::g_err; }

extern "C" int vda_device_query(int device, int* sm_count, int* cc_major, int* cc_minor) {
  cudaDeviceProp prop;
  VDA_CUDA(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  VDA_CHECK(prop.major == 10, "libvda is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  return 0;
}
