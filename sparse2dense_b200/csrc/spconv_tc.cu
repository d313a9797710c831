// spconv_tc.cu -- sparse convolution on the 5th-gen tensor cores (tcgen05, TF32 in / fp32 accumulate
// in TMEM), output-stationary gather-GEMM with the same fused epilogue as spconv_simt.cu.
//
//   out[o,:] = act( (sum_k in[tbl[k][o],:] @ W[k]) * scale + shift (+ residual[o,:]) )
//
// One CTA owns BM = 128 output rows and all Cout channels; the accumulator D[128 x Cout] lives in
// TMEM for the whole tile.  The contraction runs over (kernel offset k) x (32-channel chunk c):
//
//   warps 0-3  A producers: gather the 128 neighbour rows of offset k (LDG.128, 8 lanes per 128 B
//              row), split every value into TF32 hi + lo parts (cvt.rna) and store them into the
//              128B-swizzled K-major UMMA layout (STS.128, conflict free); fence.proxy.async +
//              mbarrier arrive.  After the main loop the same warps run the epilogue
//              (tcgen05.ld -> BN affine / residual / ReLU -> global).
//   warp 4     B loader: one cp.async.bulk per step copies the pre-swizzled, pre-split weight tile
//              (hi + lo, packed once per layer by s2d_spconv_pack_weights) and completes its bytes
//              on the stage's mbarrier; also owns the TMEM allocation.
//   warp 5     MMA issuer: one thread issues tcgen05.mma.kind::tf32 (M=128, N=Cout, K=8) x 4 per
//              step -- three per K-slice in the split-precision mode:  Ahi*Bhi + Alo*Bhi + Ahi*Blo --
//              and tcgen05.commit releases the stage / publishes the accumulator.
//
// Precision modes: TF32X3 (error-compensated, ~fp32 accuracy: this is what meets the 1e-3 parity
// bar through 21 layers) and TF32 (single pass, ~5e-4 relative error per layer).
//
// Offsets k for which no row of the tile has a neighbour are skipped by all three roles.
#include "common.cuh"

namespace s2d {

constexpr int kTcMaxK = 27;
constexpr int kBM = 128;
constexpr int kBK = 32;  // fp32 elements per row of a stage = 128 B = one swizzle row

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (ignored for swizzled K-major) | [32,46) SBO >> 4 =
//   1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format TF32 (2<<7, 2<<10), K-major A and B,
// n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of the 16 B chunk `c16` (0..7) of row `r` inside a [rows x 32 fp32] SW128 tile
__host__ __device__ __forceinline__ uint32_t sw128_chunk_offset(int r, int c16) {
  return (uint32_t)r * 128u + (uint32_t)((c16 ^ (r & 7)) << 4);
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int CIN, int COUT, int PASSES>
struct TcCfg {
  static constexpr int NCHUNK = CIN / kBK;
  static constexpr int NPART = PASSES == 3 ? 2 : 1;
  static constexpr int A_TILE = kBM * kBK * 4;   // 16 KB
  static constexpr int B_TILE = COUT * kBK * 4;  // 4 / 8 / 16 KB
  static constexpr int STAGE_BYTES = NPART * (A_TILE + B_TILE);
  static constexpr int NBR_BYTES = kTcMaxK * kBM * 4;
  static constexpr int BUDGET = 232448 - 1280;   // 227 KB opt-in limit minus barriers/alignment slack
  static constexpr int STAGES_RAW = (BUDGET - NBR_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NBR_BYTES + 256 + 1024;  // + barriers + alignment slack
  static constexpr int TMEM_COLS = COUT < 32 ? 32 : COUT;
  static_assert(CIN % kBK == 0, "Cin must be a multiple of 32");
  static_assert(COUT % 16 == 0 && COUT >= 16 && COUT <= 256, "UMMA N constraint for M=128");
  static_assert((COUT & (COUT - 1)) == 0, "TMEM allocation needs a power of two");
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
};

template <int CIN, int COUT, int PASSES>
__global__ void __launch_bounds__(192, 1)
spconv_tc_kernel(const float* __restrict__ in, const float* __restrict__ packed, const int* __restrict__ tbl,
                 int tbl_stride, int n_out, int K, const float* __restrict__ scale, const float* __restrict__ shift,
                 const float* __restrict__ residual, int relu, float* __restrict__ out) {
  using Cfg = TcCfg<CIN, COUT, PASSES>;
  constexpr int STAGES = Cfg::STAGES, NCHUNK = Cfg::NCHUNK, NPART = Cfg::NPART;
  constexpr int A_TILE = Cfg::A_TILE, B_TILE = Cfg::B_TILE;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;                                          // STAGES x [A_hi | A_lo | B_hi | B_lo]
  int* s_nbr = reinterpret_cast<int*>(smem + STAGES * Cfg::STAGE_BYTES);   // [K][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_nbr) + Cfg::NBR_BYTES);
  // bars[0..STAGES) full, [STAGES..2*STAGES) empty, [2*STAGES] accumulator ready
  uint32_t* s_misc = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);  // [0] tmem base, [1] kmask

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile0 = blockIdx.x * kBM;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum_bar = smem_u32(bars + 2 * STAGES);

  if (tid == 0) s_misc[1] = 0;
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 128 + 1);  // 128 A-producer arrives + the B loader's arrive.expect_tx
      mbar_init(empty0 + 8 * s, 1);       // one tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(s_misc), Cfg::TMEM_COLS);
  __syncthreads();
  for (int idx = tid; idx < K * kBM; idx += 192) {
    const int k = idx >> 7, r = idx & 127;
    const int row = tile0 + r;
    const int j = row < n_out ? __ldg(tbl + (size_t)k * tbl_stride + row) : -1;
    s_nbr[idx] = j;
    if (j >= 0) atomicOr(&s_misc[1], 1u << k);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_misc[0];
  const uint32_t kmask = s_misc[1];

  if (warp < 4) {
    // ===================== A producers =====================
    const int c16 = tid & 7;      // 16 B chunk inside the 128 B row
    const int r0 = tid >> 3;      // 16 rows per pass, 8 passes
    int it = 0;
    for (int k = 0; k < K; ++k) {
      if (!((kmask >> k) & 1)) continue;
      const int* nbr = s_nbr + k * kBM;
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c, ++it) {
        const int stage = it % STAGES;
        const uint32_t phase = (it / STAGES) & 1;
        float4 v[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const int j = nbr[p * 16 + r0];
          v[p] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j >= 0) v[p] = __ldg(reinterpret_cast<const float4*>(in + (size_t)j * CIN + c * kBK) + c16);
        }
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        uint8_t* a_hi = stage_base + (size_t)stage * Cfg::STAGE_BYTES;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const int r = p * 16 + r0;
          const uint32_t off = sw128_chunk_offset(r, c16);
          float4 hi;
          hi.x = tf32_rna(v[p].x); hi.y = tf32_rna(v[p].y); hi.z = tf32_rna(v[p].z); hi.w = tf32_rna(v[p].w);
          *reinterpret_cast<float4*>(a_hi + off) = hi;
          if constexpr (PASSES == 3) {
            float4 lo;
            lo.x = v[p].x - hi.x; lo.y = v[p].y - hi.y; lo.z = v[p].z - hi.z; lo.w = v[p].w - hi.w;
            *reinterpret_cast<float4*>(a_hi + A_TILE + off) = lo;
          }
        }
        fence_proxy_async();               // generic-proxy stores -> visible to the tensor core (async proxy)
        mbar_arrive(full0 + 8 * stage);
      }
    }
    // ===================== epilogue =====================
    const int row = tile0 + tid;           // TMEM lane == tile row; warp w may touch lanes [32w, 32w+32)
    if (kmask != 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < COUT; c0 += 32) {
      uint32_t acc[32];
      if (kmask != 0) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0u;
      }
      if (row < n_out) {
        float* dst = out + (size_t)row * COUT + c0;
        const float* res = residual ? residual + (size_t)row * COUT + c0 : nullptr;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 y;
          y.x = __uint_as_float(acc[4 * q + 0]); y.y = __uint_as_float(acc[4 * q + 1]);
          y.z = __uint_as_float(acc[4 * q + 2]); y.w = __uint_as_float(acc[4 * q + 3]);
          if (scale) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c0) + q);
            y.x *= sc.x; y.y *= sc.y; y.z *= sc.z; y.w *= sc.w;
          }
          if (shift) {
            const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c0) + q);
            y.x += sh.x; y.y += sh.y; y.z += sh.z; y.w += sh.w;
          }
          if (res) {
            const float4 rr = __ldg(reinterpret_cast<const float4*>(res) + q);
            y.x += rr.x; y.y += rr.y; y.z += rr.z; y.w += rr.w;
          }
          if (relu) {
            y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
          }
          reinterpret_cast<float4*>(dst)[q] = y;
        }
      }
    }
  } else if (warp == 4) {
    // ===================== B loader =====================
    if (lane == 0) {
      int it = 0;
      for (int k = 0; k < K; ++k) {
        if (!((kmask >> k) & 1)) continue;
        for (int c = 0; c < NCHUNK; ++c, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t bar = full0 + 8 * stage;
          mbar_arrive_expect_tx(bar, NPART * B_TILE);
          const float* src = packed + (size_t)(k * NCHUNK + c) * 2 * (B_TILE / 4);
          bulk_copy_g2s(smem_u32(stage_base + (size_t)stage * Cfg::STAGE_BYTES + NPART * A_TILE), src, NPART * B_TILE,
                        bar);
        }
      }
    }
  } else {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, COUT);
      int it = 0;
      for (int k = 0; k < K; ++k) {
        if (!((kmask >> k) & 1)) continue;
        for (int c = 0; c < NCHUNK; ++c, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(stage_base + (size_t)stage * Cfg::STAGE_BYTES);
          const uint32_t b_hi = a_hi + NPART * A_TILE;
#pragma unroll
          for (int s = 0; s < kBK / 8; ++s) {   // UMMA K = 8 for TF32: 32 B steps inside the swizzled row
            const uint64_t da_hi = make_desc_sw128(a_hi + 32 * s);
            const uint64_t db_hi = make_desc_sw128(b_hi + 32 * s);
            if constexpr (PASSES == 3) {
              const uint64_t da_lo = make_desc_sw128(a_hi + A_TILE + 32 * s);
              const uint64_t db_lo = make_desc_sw128(b_hi + B_TILE + 32 * s);
              umma_tf32(tmem_base, da_lo, db_hi, idesc, (it | s) != 0);   // small terms first
              umma_tf32(tmem_base, da_hi, db_lo, idesc, 1);
              umma_tf32(tmem_base, da_hi, db_hi, idesc, 1);
            } else {
              umma_tf32(tmem_base, da_hi, db_hi, idesc, (it | s) != 0);
            }
          }
          umma_commit(empty0 + 8 * stage);   // stage reusable once these MMAs have read it
        }
      }
      if (it > 0) umma_commit(accum_bar);      // accumulator complete
    }
  }
  __syncwarp();

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// weight packing: W [K, Cin, Cout] -> per (k, chunk): [hi tile | lo tile], each Cout rows x 32 fp32 in the
// exact shared-memory image the kernel's B descriptor expects (K-major, 128B swizzle), TF32-rounded.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ W, int K, int Cin, int Cout,
                                                           float* __restrict__ packed) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)K * Cin * Cout;
  if (idx >= total) return;
  const int n = (int)(idx % Cout);
  const int ci = (int)((idx / Cout) % Cin);
  const int k = (int)(idx / ((long long)Cout * Cin));
  const float w = W[idx];
  const float hi = tf32_rna(w);
  const float lo = tf32_rna(w - hi);
  const int c = ci / kBK, kk = ci % kBK;
  const int nchunk = Cin / kBK;
  const size_t tile = (size_t)Cout * kBK;                       // floats per tile
  const size_t base = (size_t)(k * nchunk + c) * 2 * tile;
  const size_t pos = (sw128_chunk_offset(n, kk >> 2) >> 2) + (kk & 3);
  packed[base + pos] = hi;
  packed[base + tile + pos] = lo;
}

template <int CIN, int COUT, int PASSES>
static int launch_tc(const float* in, const float* packed, const int* tbl, int tbl_stride, int n_out, int K,
                     const float* scale, const float* shift, const float* residual, int relu, float* out,
                     cudaStream_t st) {
  using Cfg = TcCfg<CIN, COUT, PASSES>;
  static bool configured = false;
  if (!configured) {
    S2D_CUDA(cudaFuncSetAttribute(spconv_tc_kernel<CIN, COUT, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    configured = true;
  }
  spconv_tc_kernel<CIN, COUT, PASSES><<<div_up(n_out, kBM), 192, Cfg::SMEM_BYTES, st>>>(
      in, packed, tbl, tbl_stride, n_out, K, scale, shift, residual, relu, out);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

static bool tc_supported(int Cin, int Cout) {
  return (Cin == 32 && (Cout == 32 || Cout == 64)) || (Cin == 64 && (Cout == 64 || Cout == 128)) ||
         (Cin == 128 && Cout == 128);
}

int spconv_fwd_tf32(const float* in, int n_in, const float* packed, const int* tbl, int tbl_stride, int n_out,
                    int Cin, int Cout, int K, const float* scale, const float* shift, const float* residual, int relu,
                    float* out, int passes, cudaStream_t st) {
  (void)n_in;
  if (!tc_supported(Cin, Cout)) {
    set_error("s2d_spconv_fwd: no tcgen05 kernel for Cin=%d Cout=%d (supported: 32->32/64, 64->64/128, 128->128)", Cin,
              Cout);
    return S2D_ERR_UNSUPPORTED;
  }
#define S2D_TC_CASE(ci, co)                                                                                           \
  if (Cin == ci && Cout == co)                                                                                        \
    return passes == 3 ? launch_tc<ci, co, 3>(in, packed, tbl, tbl_stride, n_out, K, scale, shift, residual, relu, out, st) \
                       : launch_tc<ci, co, 1>(in, packed, tbl, tbl_stride, n_out, K, scale, shift, residual, relu, out, st)
  S2D_TC_CASE(32, 32);
  S2D_TC_CASE(32, 64);
  S2D_TC_CASE(64, 64);
  S2D_TC_CASE(64, 128);
  S2D_TC_CASE(128, 128);
#undef S2D_TC_CASE
  return S2D_ERR_UNSUPPORTED;
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_spconv_tf32_supported(int Cin, int Cout) { return tc_supported(Cin, Cout) ? 1 : 0; }

extern "C" size_t s2d_spconv_packed_bytes(int K, int Cin, int Cout) {
  if (K < 1 || Cin < 1 || Cout < 1 || Cin % kBK) return 0;
  return (size_t)K * Cin * Cout * 2 * sizeof(float);
}

extern "C" int s2d_spconv_pack_weights(const float* W, int K, int Cin, int Cout, float* packed, void* stream) {
  S2D_REQUIRE(W && packed && K >= 1 && K <= kTcMaxK, "s2d_spconv_pack_weights: bad argument");
  S2D_REQUIRE(tc_supported(Cin, Cout), "s2d_spconv_pack_weights: unsupported shape Cin=%d Cout=%d", Cin, Cout);
  pack_weights_kernel<<<div_up((long long)K * Cin * Cout, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      W, K, Cin, Cout, packed);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
