// srw_api — library-level C ABI: version, error string, device check, launch counter.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <vector>

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

// ---- profiling -------------------------------------------------------------------------------------------------
struct ProfRec { int cls; cudaEvent_t e0, e1; double flops, bytes; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
bool g_prof_on = false;

void* prof_begin(int cls, double flops, double bytes, cudaStream_t s) {
  if (!g_prof_on) return nullptr;
  ProfRec* r = new ProfRec{cls, nullptr, nullptr, flops, bytes};
  cudaEventCreate(&r->e0);
  cudaEventCreate(&r->e1);
  cudaEventRecord(r->e0, s);
  return r;
}
void prof_end(void* h, cudaStream_t s) {
  if (!h) return;
  ProfRec* r = reinterpret_cast<ProfRec*>(h);
  cudaEventRecord(r->e1, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(*r);
  delete r;
}

static int g_pdl_mode = -1;   // -1: read SRW_PDL on first use
bool pdl_enabled() {
  if (g_pdl_mode < 0) {
    const char* e = getenv("SRW_PDL");
    g_pdl_mode = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl_mode != 0;
}

}  // namespace srw

extern "C" int srw_set_pdl_mode(int on) {
  srw::g_pdl_mode = on ? 1 : 0;
  return SRW_OK;
}

extern "C" int srw_profile_enable(int on) {
  srw::g_prof_on = on != 0;
  return SRW_OK;
}

extern "C" int srw_profile_collect(srw_profile_stats* out) {
  SRW_REQUIRE(out, "srw_profile_collect: null pointer");
  SRW_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < SRW_PROF_NUM; ++i) out[i] = srw_profile_stats{0, 0.0, 0.0, 0.0};
  std::lock_guard<std::mutex> lk(srw::g_prof_mu);
  for (auto& r : srw::g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    if (r.cls >= 0 && r.cls < SRW_PROF_NUM) {
      out[r.cls].launches += 1; out[r.cls].total_ms += ms; out[r.cls].flops += r.flops; out[r.cls].bytes += r.bytes;
    }
  }
  srw::g_prof.clear();
  return SRW_OK;
}

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
