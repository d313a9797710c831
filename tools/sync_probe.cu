// sync_probe.cu -- latencies of the hand-offs the warp-specialised kernels are built from (one CTA, clock64):
//   (a) mbarrier ping-pong between two warps (arrive -> try_wait wake-up), (b) tcgen05.commit with nothing outstanding ->
//   waiter wakes, (c) one MMA (N = 128) + commit -> waiter wakes, (d) cp.async.bulk 16 KB from L2 -> waiter wakes,
//   (e) 32 x cp.async 16 B per lane (one gathered tile) -> wait_group 0.
// Build on the GPU box: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/sync_probe tools/sync_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(128) probe(const uint8_t* gsrc, long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[4];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (64 << 10) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  const uint32_t b0 = smem_u32(bars), b1 = smem_u32(bars + 1), b2 = smem_u32(bars + 2), b3 = smem_u32(bars + 3);
  // (a) ping-pong: warp 0 arrives on b0, warp 1 waits b0 and arrives on b1, warp 0 waits b1
  if (warp == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (lane == 0) mbar_arrive(b0);
      mbar_wait(b1, i & 1);
    }
    if (lane == 0) out[0] = (clock64() - t0) / iters;
  } else if (warp == 1) {
    for (int i = 0; i < iters; ++i) {
      mbar_wait(b0, i & 1);
      if (lane == 0) mbar_arrive(b1);
    }
  }
  __syncthreads();
  // (b) empty commit -> own wait ; (c) one MMA + commit -> own wait
  if (warp == 0) {
    const bool leader = elect_one();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (leader) commit(b2);
      __syncwarp();
      mbar_wait(b2, i & 1);
    }
    if (lane == 0) out[1] = (clock64() - t0) / iters;
    const uint64_t hi = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;
    const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3FFF) | (1u << 16);
    const uint32_t w_lo = ((smem_u32(smem + (32 << 10)) >> 4) & 0x3FFF) | (1u << 16);
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (leader) { mma(tmem, hi | a_lo, hi | w_lo, idesc, 0u); commit(b3); }
      __syncwarp();
      mbar_wait(b3, i & 1);
    }
    if (lane == 0) out[2] = (clock64() - t0) / iters;
    // six MMAs + commit
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (leader) {
        mma(tmem, hi | a_lo, hi | w_lo, idesc, 0u); mma(tmem, hi | (a_lo + 2), hi | (w_lo + 2), idesc, 1u);
        mma(tmem, hi | (a_lo + 4), hi | (w_lo + 4), idesc, 1u); mma(tmem, hi | (a_lo + 6), hi | (w_lo + 6), idesc, 1u);
        mma(tmem, hi | a_lo, hi | (w_lo + 4), idesc, 1u); mma(tmem, hi | (a_lo + 2), hi | (w_lo + 6), idesc, 1u);
        commit(b3);
      }
      __syncwarp();
      mbar_wait(b3, (iters + i) & 1);
    }
    if (lane == 0) out[3] = (clock64() - t0) / iters;
  }
  __syncthreads();
  // (d) bulk copy 16 KB (L2 resident after the first) -> wait
  if (warp == 0) {
    uint32_t par = (uint32_t)(iters & 1);                   // phases b0 went through in (a)
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b0), "r"(16384) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
                     "l"(gsrc + (size_t)(i & 15) * 16384), "r"(16384), "r"(b0) : "memory");
      }
      mbar_wait(b0, par);
      par ^= 1;
    }
    if (lane == 0) out[4] = (clock64() - t0) / iters;
    // (e) a gathered tile: 32 cp.async of 16 B per lane, rows 2 KB apart
    const long long t1 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const uint32_t dst = smem_u32(smem) + (uint32_t)(q * 512 + lane * 16);
        const uint8_t* src = gsrc + (size_t)((i * 37 + q * 4 + (lane >> 3)) & 1023) * 2048 + (lane & 7) * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(dst), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (lane == 0) out[5] = (clock64() - t1) / iters;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main() {
  uint8_t* g;
  long long* d;
  cudaMalloc(&g, 4 << 20);
  cudaMemset(g, 0, 4 << 20);
  cudaMalloc(&d, 64 * sizeof(long long));
  const int smem = (64 << 10) + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    probe<<<1, 128, smem>>>(g, d, 200);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  }
  long long h[8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("mbarrier ping-pong round trip (2 hand-offs): %lld clk\n", h[0]);
  printf("tcgen05.commit (nothing outstanding) -> wake: %lld clk\n", h[1]);
  printf("1 MMA (N=128) + commit -> wake: %lld clk\n", h[2]);
  printf("6 MMAs (N=128) + commit -> wake: %lld clk\n", h[3]);
  printf("cp.async.bulk 16 KB from L2 -> wake: %lld clk\n", h[4]);
  printf("32 x cp.async 16 B per lane (16 KB tile) + wait_group 0: %lld clk\n", h[5]);
  return 0;
}
