// Error reporting, launch accounting and version for the C ABI (include/vglamm.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.h"

namespace vg {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(VG_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return VG_OK;
}
}  // namespace vg

extern "C" const char* vg_last_error(void) { return vg::g_err; }
extern "C" int vg_version(void) { return 1; }
extern "C" long long vg_launch_count(void) { return vg::g_launches.load(); }
extern "C" void vg_launch_count_reset(void) { vg::g_launches.store(0); }
extern "C" void vg_launch_count_add(long long n) { vg::g_launches.fetch_add(n); }
