// l2_probe.cu -- how fast can all SMs pull L2-resident data into shared memory?  (the ceiling of a gather-GEMM whose
// operands are re-read from L2 once per kernel offset)
//   mode 0: cp.async (LDGSTS) 16 B per lane, 8 warps x 32 copies in flight per SM, rows of 128 B at pseudo-random positions
//           (the gather of conv_bf2.cu);  mode 1: the same, consecutive rows;  mode 2: cp.async.bulk 16 KB pieces, ring of 4;
//   mode 3: LDG.128 into registers, 32 warps.
// Build on the GPU box: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/l2_probe tools/l2_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(1024) probe(const uint8_t* __restrict__ src, size_t bytes, int mode, int iters, float* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t rows = bytes / 128;
  if (mode <= 1) {
    // warp w fills its own 16 KB stage (nw warps, stages reused modulo 8): 32 copies per lane = 128 rows x 128 B per iteration;
    // working sets are powers of two, rows advance by an odd stride (mode 0: scattered 4-row groups) or by 4 (mode 1)
    const int nw = blockDim.x >> 5;
    const uint32_t dst0 = smem_u32(smem) + (uint32_t)(warp & 7) * 16384u + (uint32_t)(lane >> 3) * 128u + (uint32_t)(lane & 7) * 16u;
    const size_t mask = rows - 1;
    size_t r = ((size_t)blockIdx.x * nw + warp) * 128 * 977;
    const size_t step = mode == 0 ? 4 * 7919 : 4;
    const uint8_t* lane_src = src + (lane & 7) * 16;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const size_t row = ((r & mask) & ~size_t(3)) + (lane >> 3);
        r += step;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(dst0 + 512u * q), "l"(lane_src + row * 128) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");       // two tiles in flight per warp
      if (mode == 1) r += (size_t)(148 * nw - 1) * 128;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (mode == 2) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)));
      asm volatile("fence.mbarrier_init.release.cluster;");
      const size_t pieces = bytes / 16384;
      for (int it = 0; it < iters + 4; ++it) {
        const int s = it & 3;
        if (it >= 4) mbar_wait(smem_u32(bars + s), ((it >> 2) - 1) & 1);
        if (it < iters) {
          const size_t piece = ((size_t)blockIdx.x + (size_t)it * 148) % pieces;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(16384) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(smem) + (uint32_t)s * 16384u),
                       "l"(src + piece * 16384), "r"(16384), "r"(smem_u32(bars + s))
                       : "memory");
        }
      }
    }
  } else {
    float acc = 0.f;
    const size_t n16 = bytes / 16;
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) % n16;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src) + i);
        acc += v.x + v.y + v.z + v.w;
        i += (size_t)gridDim.x * blockDim.x;
        if (i >= n16) i -= n16;
      }
    }
    if (acc == 12345.f) *sink = acc;
  }
}

int main() {
  uint8_t* buf;
  float* sink;
  const size_t cap = 512ull << 20;
  cudaMalloc(&buf, cap);
  cudaMemset(buf, 0, cap);
  cudaMalloc(&sink, 4);
  const int smem = (128 << 10) + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const char* names[4] = {"cp.async 16 B, scattered 128 B rows, 2 x 16 KB in flight / warp", "cp.async 16 B, consecutive rows", "cp.async.bulk 16 KB, ring of 4",
                          "LDG.128, 32 warps x 8 in flight"};
  for (int mode = 0; mode < 4; ++mode)
    for (int nwarps : {8, 16, 32})
    for (size_t mb : {32, 64, 256}) {
      if (mode >= 2 && nwarps != 8) continue;
      const size_t bytes = mb << 20;
      const int threads = mode == 3 ? 1024 : 32 * nwarps;
      const int iters = mode == 3 ? 64 : 400;
      const double moved = mode == 3 ? 148.0 * threads * iters * 8 * 16 : (mode == 2 ? 148.0 * iters * 16384 : 148.0 * nwarps * iters * 16384);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      probe<<<148, threads, smem>>>(buf, bytes, mode, iters, sink);      // warm the L2
      cudaEventRecord(e0);
      probe<<<148, threads, smem>>>(buf, bytes, mode, iters, sink);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("%-60s %2d warps, working set %3zu MB: %7.2f TB/s = %5.1f B/clk/SM at %d MHz\n", names[mode], threads / 32, mb, moved / ms * 1e-9,
             moved / (ms * 1e-3) / 148.0 / (clk_khz * 1e3), clk_khz / 1000);
    }
  return 0;
}
