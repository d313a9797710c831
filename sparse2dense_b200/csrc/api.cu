// api.cu -- version and error reporting of libs2d_b200.so.
#include <stdarg.h>

#include <atomic>
#include <string.h>

#include "common.cuh"

namespace s2d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
  return S2D_ERR_CUDA;
}

}  // namespace s2d

namespace s2d {
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace s2d

extern "C" unsigned long long s2d_kernel_launches(void) { return s2d::g_launches.load(); }
extern "C" int s2d_version(void) { return 100; }
extern "C" const char* s2d_last_error(void) { return s2d::g_err; }
