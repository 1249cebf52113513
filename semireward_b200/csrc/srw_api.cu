// srw_api — library-level C ABI: version, error string, device check, launch counter.
#include <stdarg.h>

#include <atomic>

#include "../../include/srw.h"
#include "srw_common.cuh"

namespace srw {

std::atomic<int64_t> g_launches{0};
static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_last_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return SRW_ERR_CUDA;
}

}  // namespace srw

extern "C" int srw_version(void) { return 1; }

extern "C" const char* srw_last_error(void) { return srw::g_err; }

extern "C" int64_t srw_kernel_launches(void) { return srw::g_launches.load(); }

extern "C" int srw_device_check(int* sm_major, int* sm_minor, int* sm_count) {
  int dev = 0;
  SRW_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SRW_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_major) *sm_major = prop.major;
  if (sm_minor) *sm_minor = prop.minor;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {
    srw::set_last_error("semireward_b200 is built for sm_100a only; current device is sm_%d%d (%s)", prop.major, prop.minor, prop.name);
    return SRW_ERR_UNSUPPORTED;
  }
  return SRW_OK;
}
