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

namespace s2d {
__global__ void export_i32_kernel(const int* __restrict__ src, int n, volatile int* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace s2d

// Small control values (row counts) go to the host through MAPPED pinned memory written by a kernel instead of a
// cudaMemcpy: a DMA copy would queue behind any bulk device->host transfer in flight on the copy engine (the
// pipelined result read-back) and stall the next step's launch sequence for milliseconds.
extern "C" int s2d_export_i32(const int* src, int n, int* dst_host_mapped, void* stream) {
  S2D_REQUIRE(n >= 0 && (n == 0 || (src && dst_host_mapped)), "s2d_export_i32: bad argument");
  if (n == 0) return S2D_OK;
  s2d::export_i32_kernel<<<1, n < 256 ? 32 * ((n + 31) / 32) : 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, n, dst_host_mapped);
  S2D_LAUNCH_CHECK();
  s2d::count_launches(1);
  return S2D_OK;
}

extern "C" unsigned long long s2d_kernel_launches(void) { return s2d::g_launches.load(); }
extern "C" int s2d_version(void) { return 100; }
extern "C" const char* s2d_last_error(void) { return s2d::g_err; }

// fp32 FFMA microbenchmark (tools/measure_peaks.py; not part of the public header): every thread runs `iters` rounds of 16
// independent FMAs, 2 * 16 * iters * threads flops in total.  The denominator for the kernels that still run on CUDA cores.
namespace s2d {
__global__ void __launch_bounds__(256) ffma_peak_kernel(int iters, float* __restrict__ sink) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
  const float m = 1.0000001f, c = 1e-9f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678f) sink[0] = s;            // never true: keeps the loop alive
}
}  // namespace s2d
extern "C" long long s2d_debug_ffma(int iters, float* sink, void* stream) {
  const int blocks = s2d::kNumSMs * 16;
  s2d::ffma_peak_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(iters, sink);
  return 2LL * 16 * iters * blocks * 256;       // flops of the launch
}
