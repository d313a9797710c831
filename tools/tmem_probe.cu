// tmem_probe.cu -- empirical probe of tcgen05.st fragment layouts and of TS-mode (A in TMEM) tcgen05.mma.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tmem_probe tools/tmem_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_zero16(uint32_t taddr) {
  uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// shape: 0 = 16x64b.x1 (1 reg), 1 = 16x128b.x1 (2 regs), 2 = 16x256b.x1 (4 regs), 3 = 16x128b.x2 (4 regs),
//        4 = 16x256b.x2 (8 regs), 5 = 32x32b.x2 (2 regs)
__global__ void probe_st(int shape, uint32_t* out /*[128][16]*/) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = s_tmem;
  const uint32_t mine = base + ((uint32_t)(warp * 32) << 16);
  tmem_st_zero16(mine);
  __syncthreads();
  if (warp == 0) {
    uint32_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = 0x10000u + (uint32_t)lane * 256u + (uint32_t)i;
    if (shape == 0) asm volatile("tcgen05.st.sync.aligned.16x64b.x1.b32 [%0], {%1};" ::"r"(mine), "r"(r[0]) : "memory");
    if (shape == 1) asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1,%2};" ::"r"(mine), "r"(r[0]), "r"(r[1]) : "memory");
    if (shape == 2) asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(mine), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
    if (shape == 3) asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" ::"r"(mine), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
    if (shape == 4) asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(mine), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    if (shape == 5) asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(mine), "r"(r[0]), "r"(r[1]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[16];
  tmem_ld16(mine, v);
  for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 16 + i] = v[i];
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32));
}

// D[128 x 32] = A[128 x 32](TMEM, lane = m, column = k) * B[32 x 32]^T (smem, K-major SW128), TF32
__global__ void probe_ts_mma(const float* A, const float* B, float* D) {
  __shared__ __align__(1024) uint8_t s_b[32 * 128];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // B tile: row n (0..31), 32 floats, 128B-swizzled 16 B chunks
  for (int i = tid; i < 32 * 32; i += 128) {
    const int n = i / 32, k = i % 32;
    const uint32_t off = n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4;
    *reinterpret_cast<float*>(s_b + off) = B[n * 32 + k];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = s_tmem;
  const uint32_t mine = base + ((uint32_t)(warp * 32) << 16);
  // A: lane m = tid, columns 32..63 hold A[m][0..31]
  {
    const float* a = A + tid * 32;
    for (int c = 0; c < 32; c += 8) {
      uint32_t r[8];
      for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(a[c + i]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(mine + 32 + c), "r"(r[0]),
                   "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t desc_hi = 64u | (1u << 14) | (2u << 29);
    for (int q = 0; q < 4; ++q) {
      const uint64_t db = ((uint64_t)desc_hi << 32) | ((((smem_u32(s_b) >> 4) & 0x3FFF) | (1u << 16)) + 2 * q);
      const uint32_t acc = q != 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(base),
          "r"(base + 32 + 8 * q), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[16];
  tmem_ld16(mine, v);
  for (int i = 0; i < 16; ++i) D[tid * 32 + i] = __uint_as_float(v[i]);
  tmem_ld16(mine + 16, v);
  for (int i = 0; i < 16; ++i) D[tid * 32 + 16 + i] = __uint_as_float(v[i]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64));
}

int main() {
  const char* names[] = {"16x64b.x1", "16x128b.x1", "16x256b.x1", "16x128b.x2", "16x256b.x2", "32x32b.x2"};
  uint32_t* d_out;
  cudaMalloc(&d_out, 128 * 16 * 4);
  static uint32_t h[128 * 16];
  for (int shape = 0; shape < 6; ++shape) {
    probe_st<<<1, 128>>>(shape, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shape %s: CUDA error %s\n", names[shape], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("== tcgen05.st.%s : (tmem lane, column) <- (thread, reg)\n", names[shape]);
    for (int l = 0; l < 128; ++l)
      for (int c = 0; c < 16; ++c)
        if (h[l * 16 + c]) printf("  lane %3d col %2d <- thread %2u reg %u\n", l, c, (h[l * 16 + c] >> 8) & 0xff, h[l * 16 + c] & 0xff);
  }
  // TS-mode MMA check
  static float A[128 * 32], B[32 * 32], D[128 * 32];
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) A[m * 32 + k] = (float)((m * 3 + k) % 7 - 3) + 0.5f * (float)(k % 3);
  for (int n = 0; n < 32; ++n) for (int k = 0; k < 32; ++k) B[n * 32 + k] = (float)((n + 2 * k) % 5 - 2);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof(A)); cudaMalloc(&dB, sizeof(B)); cudaMalloc(&dD, sizeof(D));
  cudaMemcpy(dA, A, sizeof(A), cudaMemcpyHostToDevice); cudaMemcpy(dB, B, sizeof(B), cudaMemcpyHostToDevice);
  probe_ts_mma<<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("TS mma: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(D, dD, sizeof(D), cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
    double ref = 0; for (int k = 0; k < 32; ++k) ref += (double)A[m * 32 + k] * B[n * 32 + k];
    double er = fabs(ref - D[m * 32 + n]); if (er > maxerr) maxerr = er;
  }
  printf("== TS-mode tcgen05.mma (A in TMEM lane=m col=k): max abs err %g -> %s\n", maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
  return 0;
}
