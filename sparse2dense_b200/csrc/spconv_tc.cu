// spconv_tc.cu -- sparse convolution on the 5th-gen tensor cores (tcgen05, TF32 in / fp32 accumulate
// in TMEM), output-stationary gather-GEMM with the same fused epilogue as spconv_simt.cu.
//
//   out[o,:] = act( (sum_k in[tbl[k][o],:] @ W[k]) * scale + shift (+ residual[o,:]) )
//
// One CTA owns T = 4 tiles of BM = 128 output rows and all Cout channels; the T accumulators D[128 x Cout]
// live in TMEM (T*Cout <= 512 columns) for the whole CTA.  The contraction runs over
// (kernel offset k) x (32-channel chunk c) x (tile t); the weight tile of (k,c) is fetched ONCE per CTA and
// reused by the T tiles, which divides the L2->SM weight traffic (the dominant stream at Cout=128) by T.
//
//   warps 0-7  A producers: gather the neighbour rows of (k, t) (LDG.128, 8 lanes per 128 B row, three
//              steps of loads in flight per warp), split every value into TF32 hi + lo parts (cvt.rna) and
//              store them into the 128B-swizzled K-major UMMA layout (STS.128, conflict free);
//              fence.proxy.async + mbarrier arrive.  After the main loop the same warps run the epilogue
//              (tcgen05.ld -> BN affine / residual / ReLU -> global).
//   warp 8     loader: neighbour-table slices of offset k+2 (coalesced copy into a 4-slot ring) and one
//              cp.async.bulk per (k,c) for the pre-swizzled, pre-split weight tile (hi + lo, packed once
//              per layer by s2d_spconv_pack_weights), completing on the slot's mbarrier; owns TMEM alloc.
//   warp 9     MMA issuer: one thread issues tcgen05.mma.kind::tf32 (M=128, N=Cout, K=8) x 4 per step --
//              three per K-slice in the split-precision mode:  Alo*Bhi + Ahi*Blo + Ahi*Bhi -- and
//              tcgen05.commit releases the rings / publishes the accumulators.
//
// Precision modes: TF32X3 (error-compensated, ~fp32 accuracy: this is what meets the 1e-3 parity
// bar through 21 layers) and TF32 (single pass).
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
// D[tmem] (+)= A[tmem] * B[smem]  (TS mode: the A operand is read from tensor memory, lane = row, column = k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 lanes x 32 columns: reg 4n+e -> (lane l/4, column 8n + 2(l%4) + e), reg 4n+2+e -> lane l/4 + 8 (probed on B200)
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// TF32 keeps 10 explicit mantissa bits: drop (truncate) or round-to-nearest (ties away) the low 13 bits
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float tf32_round(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
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
constexpr int kProducerWarps = 8;
constexpr int kLoaderWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kTcThreads = 320;
constexpr int kNbrSlots = 4;

// A tile element (row m, channel ch of the 32-channel chunk) lives in TMEM lane m, column a_col(ch).  The
// permutation makes the 8 columns a thread owns in a tcgen05.st.16x256b.x4 fragment
// (columns 8n + 2j + e, j = lane % 4) equal to 8 CONTIGUOUS channels [8j, 8j+8) of its row, so the gather is
// two LDG.128 per row.  The weight tiles are packed with the same K permutation (pack_weights_kernel).
__host__ __device__ constexpr int a_col_of_channel(int ch) { return 8 * ((ch % 8) / 2) + 2 * (ch / 8) + (ch % 2); }

template <int COUT, int PASSES>
struct TcCfg {
  static constexpr int NPART = PASSES == 3 ? 2 : 1;
  static constexpr int T = COUT >= 128 ? 2 : 4;        // 128-row tiles per CTA sharing every weight tile
  static constexpr int ROWS = T * kBM;
  static constexpr int B_TILE = COUT * kBK * 4;        // 4 / 8 / 16 KB
  static constexpr int B_STAGE = NPART * B_TILE;
  static constexpr int SB = 4;                         // weight-tile ring (shared memory)
  static constexpr int ACC_COLS = T * COUT;            // TMEM: accumulators ...
  static constexpr int A_COLS = NPART * kBK;           // ... and one gathered A stage (hi | lo), 32 columns each
  static constexpr int SA_RAW = (512 - ACC_COLS) / A_COLS;
  static constexpr int SA = SA_RAW > 8 ? 8 : SA_RAW;   // gathered-tile ring (TMEM)
  static constexpr int TMEM_COLS = 512;
  static constexpr int NBR_BYTES = kNbrSlots * ROWS * 4;
  static constexpr int SMEM_BYTES = SB * B_STAGE + NBR_BYTES + 1024 + 1024;
  static_assert(COUT % 16 == 0 && COUT >= 32 && COUT <= 128, "UMMA N constraint for M=128 / TMEM budget");
  static constexpr int BATCH = 1;                      // steps per producer hand-off (2 measured slower except 128->128 TF32)
  static_assert(SA >= 2 * BATCH && T % BATCH == 0, "gathered-tile ring must hold two producer batches");
};

// One CTA = T row tiles.  Loop nest: kernel offset k > channel chunk c > tile t; the weight tile of (k,c) is
// loaded once and feeds T MMA groups (one per tile accumulator).  The gathered A operand never touches
// shared memory: the producer warps write it straight into TMEM (tcgen05.st) and the MMAs run in TS mode,
// so shared-memory bandwidth only carries the small weight slices.
// Arguments of one launch.  Rows may be strided (in_ld / out_ld / res_ld, in floats) so that layers can read
// from and write into channel slices of wider buffers (concatenations); blockIdx.y selects a block of COUT
// output channels (weights are packed per block); out_rows optionally remaps output rows (sub-pixel
// transposed convolutions).
struct ConvArgs {
  const float* in;
  const float* packed;
  const int* tbl;
  const float* scale;
  const float* shift;
  const float* residual;
  float* out;
  const int* out_rows;
  int in_ld, out_ld, res_ld, tbl_stride, n_out, K, nchunk, act, res_after_act;
};

__device__ __forceinline__ float apply_act(float y, int act) {
  if (act == S2D_ACT_RELU) return fmaxf(y, 0.f);
  if (act == S2D_ACT_GELU) return 0.5f * y * (1.f + erff(y * 0.70710678118654752440f));   // exact GELU (torch default)
  return y;
}

template <int COUT, int PASSES>
__global__ void __launch_bounds__(kTcThreads, 1) spconv_tc_kernel(const __grid_constant__ ConvArgs A) {
  using Cfg = TcCfg<COUT, PASSES>;
  constexpr int SA = Cfg::SA, SB = Cfg::SB, T = Cfg::T;
  constexpr int B_TILE = Cfg::B_TILE, B_STAGE = Cfg::B_STAGE, A_COLS = Cfg::A_COLS;
  const int NCHUNK = A.nchunk, K = A.K, n_out = A.n_out;
  const size_t blk_tiles = (size_t)blockIdx.y * K * NCHUNK;            // weight tiles before this channel block
  const float* __restrict__ packed = A.packed + blk_tiles * 2 * (B_TILE / 4);
  const int* __restrict__ tbl = A.tbl;
  const int tbl_stride = A.tbl_stride;
  const int cblk = blockIdx.y * COUT;                                  // first output channel of this block

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_ring = smem;                                       // SB x [B_hi | B_lo]
  int* s_nbr = reinterpret_cast<int*>(b_ring + SB * B_STAGE);   // kNbrSlots x [T*128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_nbr) + Cfg::NBR_BYTES);
  uint64_t* bar_a_full = bars;
  uint64_t* bar_a_empty = bar_a_full + SA;
  uint64_t* bar_b_full = bar_a_empty + SA;
  uint64_t* bar_b_empty = bar_b_full + SB;
  uint64_t* bar_n_full = bar_b_empty + SB;
  uint64_t* bar_n_empty = bar_n_full + kNbrSlots;
  uint64_t* bar_accum = bar_n_empty + kNbrSlots;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_accum + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile0 = blockIdx.x * Cfg::ROWS;
  const int nsteps = K * NCHUNK * T;

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(smem_u32(bar_a_full + s), kProducerWarps);        // one arrival per producer warp
      mbar_init(smem_u32(bar_a_empty + s), 1);                    // one tcgen05.commit
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(smem_u32(bar_b_full + s), 1);                     // arrive.expect_tx of the loader
      mbar_init(smem_u32(bar_b_empty + s), 1);
    }
    for (int s = 0; s < kNbrSlots; ++s) {
      mbar_init(smem_u32(bar_n_full + s), 32);                    // every loader lane arrives
      mbar_init(smem_u32(bar_n_empty + s), kProducerWarps);
    }
    mbar_init(smem_u32(bar_accum), 1);
    fence_barrier_init();
  }
  if (warp == kLoaderWarp) tmem_alloc(smem_u32(s_tmem), Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem_a0 = tmem_base + Cfg::ACC_COLS;            // first column of the A ring

  if (warp < kProducerWarps) {
    // ===================== A producers (8 warps x 16 rows) =====================
    // warp w owns TMEM lanes (= tile rows) [32*(w%4) + 16*(w/4), +16); lane l: rows rA = base + l/4 and rA + 8,
    // channels [8*(l%4), +8) of the chunk (see a_col_of_channel).
    const int row16 = 32 * (warp & 3) + 16 * (warp >> 2);
    const int rA = row16 + (lane >> 2);
    const int j4 = (lane & 3) * 2;               // float4 index of this lane's 8 channels inside the chunk
    const float4* in4 = reinterpret_cast<const float4*>(A.in);
    const int in_ld4 = A.in_ld >> 2;
    const uint32_t bar_a_full0 = smem_u32(bar_a_full), bar_a_empty0 = smem_u32(bar_a_empty);
    const uint32_t bar_n_full0 = smem_u32(bar_n_full), bar_n_empty0 = smem_u32(bar_n_empty);
    const uint32_t tmem_mine = tmem_a0 + ((uint32_t)row16 << 16);

    // Steps are produced in batches of BATCH: one empty-wait / tcgen05.wait::st / fence / arrive sequence per
    // batch amortises the fixed latencies of the hand-off; the loads of the next batch are already in flight.
    constexpr int BATCH = Cfg::BATCH;
    int lk = 0, lc = 0, lt = 0;                  // load stream position (k, c, t)
    int sstage = 0;                              // store stream position
    uint32_t sphase = 1;                         // first pass over the ring: slots are free
    // v[b][0..1] = row rA channels 8j..8j+7, v[b][2..3] = row rA+8
    auto load_batch = [&](float4 (&v)[BATCH][4]) {
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const int slot = lk & (kNbrSlots - 1);
        if ((lc | lt) == 0) mbar_wait(bar_n_full0 + 8 * slot, (lk / kNbrSlots) & 1);
        const int* nbr = s_nbr + slot * Cfg::ROWS + lt * kBM + rA;
        const int ja = nbr[0], jb = nbr[8];
        v[b][0] = v[b][1] = v[b][2] = v[b][3] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ja >= 0) {
          const float4* p = in4 + ((size_t)ja * in_ld4 + lc * (kBK / 4) + j4);
          v[b][0] = __ldg(p); v[b][1] = __ldg(p + 1);
        }
        if (jb >= 0) {
          const float4* p = in4 + ((size_t)jb * in_ld4 + lc * (kBK / 4) + j4);
          v[b][2] = __ldg(p); v[b][3] = __ldg(p + 1);
        }
        if (++lt == T) {
          lt = 0;
          if (++lc == NCHUNK) {
            lc = 0;
            __syncwarp();                        // every lane has read this k's indices
            if (lane == 0) mbar_arrive(bar_n_empty0 + 8 * slot);
            ++lk;
          }
        }
      }
    };
    auto store_batch = [&](const float4 (&v)[BATCH][4]) {
      const int stage0 = sstage;
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        mbar_wait(bar_a_empty0 + 8 * sstage, sphase);
        if (b == BATCH - 1) tc_fence_after();
        if (++sstage == SA) { sstage = 0; sphase ^= 1; }
      }
      int st = stage0;
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        // fragment order of tcgen05.st.16x256b.x4: reg 4n+e = row rA, column 8n+2j+e; reg 4n+2+e = row rA+8
        const float xa[8] = {v[b][0].x, v[b][0].y, v[b][0].z, v[b][0].w, v[b][1].x, v[b][1].y, v[b][1].z, v[b][1].w};
        const float xb[8] = {v[b][2].x, v[b][2].y, v[b][2].z, v[b][2].w, v[b][3].x, v[b][3].y, v[b][3].z, v[b][3].w};
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float a = xa[2 * n + e], c = xb[2 * n + e];
            if constexpr (PASSES == 3) {
              const float ah = tf32_trunc(a), ch = tf32_trunc(c);   // x = hi + lo exactly
              hi[4 * n + e] = __float_as_uint(ah);      lo[4 * n + e] = __float_as_uint(a - ah);
              hi[4 * n + 2 + e] = __float_as_uint(ch);  lo[4 * n + 2 + e] = __float_as_uint(c - ch);
            } else {
              hi[4 * n + e] = __float_as_uint(tf32_round(a));
              hi[4 * n + 2 + e] = __float_as_uint(tf32_round(c));
            }
          }
        const uint32_t dst = tmem_mine + (uint32_t)(st * A_COLS);
        tmem_st_16x256b_x4(dst, hi);
        if constexpr (PASSES == 3) tmem_st_16x256b_x4(dst + kBK, lo);
        if (++st == SA) st = 0;
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                           // one arrival per producer warp and stage
        st = stage0;
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          mbar_arrive(bar_a_full0 + 8 * st);
          if (++st == SA) st = 0;
        }
      }
    };

    const int nbatches = nsteps / BATCH;         // nsteps = K * NCHUNK * T is a multiple of T >= BATCH
    float4 bufA[BATCH][4], bufB[BATCH][4];
    if (nbatches > 0) load_batch(bufA);
    for (int nb = 0; nb < nbatches; nb += 2) {
      if (nb + 1 < nbatches) load_batch(bufB);
      store_batch(bufA);
      if (nb + 2 < nbatches) load_batch(bufA);
      if (nb + 1 < nbatches) store_batch(bufB);
    }

    // ===================== epilogue (same 8 warps) =====================
    // TMEM lane == row inside a tile; warp w may touch lanes [32*(w%4), +32).  Warps 0-3 take the even tiles,
    // warps 4-7 the odd ones.
    mbar_wait(smem_u32(bar_accum), 0);
    tc_fence_after();
    const int g = warp & 3;
    const float* __restrict__ scale = A.scale ? A.scale + cblk : nullptr;
    const float* __restrict__ shift = A.shift ? A.shift + cblk : nullptr;
    const int act = A.act, res_after = A.res_after_act;
#pragma unroll 1
    for (int t = warp >> 2; t < T; t += 2) {
      const int row = tile0 + t * kBM + g * 32 + lane;
      const int orow = (row < n_out && A.out_rows) ? __ldg(A.out_rows + row) : row;
#pragma unroll 1
      for (int c0 = 0; c0 < COUT; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(t * COUT + c0), acc);
        if (row < n_out) {
          float* dst = A.out + (size_t)orow * A.out_ld + cblk + c0;
          const float* res = A.residual ? A.residual + (size_t)orow * A.res_ld + cblk + c0 : nullptr;
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
            float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
            if (res) rr = __ldg(reinterpret_cast<const float4*>(res) + q);
            if (!res_after) { y.x += rr.x; y.y += rr.y; y.z += rr.z; y.w += rr.w; }
            y.x = apply_act(y.x, act); y.y = apply_act(y.y, act); y.z = apply_act(y.z, act); y.w = apply_act(y.w, act);
            if (res_after) { y.x += rr.x; y.y += rr.y; y.z += rr.z; y.w += rr.w; }
            reinterpret_cast<float4*>(dst)[q] = y;
          }
        }
      }
    }
  } else if (warp == kLoaderWarp) {
    // ===================== loader: neighbour-table slices (all lanes) + weight tiles (lane 0) =====================
    auto load_nbr = [&](int k) {
      const int slot = k % kNbrSlots;
      mbar_wait(smem_u32(bar_n_empty + slot), ((k / kNbrSlots) & 1) ^ 1);
      int* dst = s_nbr + slot * Cfg::ROWS;
      const int* src = tbl + (size_t)k * tbl_stride + tile0;
      int v[Cfg::ROWS / 32];
#pragma unroll
      for (int i = 0; i < Cfg::ROWS / 32; ++i)     // all loads in flight before the first store
        v[i] = (tile0 + lane + 32 * i < n_out) ? __ldg(src + lane + 32 * i) : -1;
#pragma unroll
      for (int i = 0; i < Cfg::ROWS / 32; ++i) dst[lane + 32 * i] = v[i];
      mbar_arrive(smem_u32(bar_n_full + slot));
    };
    if (0 < K) load_nbr(0);
    if (1 < K) load_nbr(1);
    for (int k = 0; k < K; ++k) {
      if (k + 2 < K) load_nbr(k + 2);
      if (lane == 0) {
        for (int c = 0; c < NCHUNK; ++c) {
          const int bstep = k * NCHUNK + c;
          const int bs = bstep % SB;
          mbar_wait(smem_u32(bar_b_empty + bs), ((bstep / SB) & 1) ^ 1);
          const uint32_t bar = smem_u32(bar_b_full + bs);
          mbar_arrive_expect_tx(bar, B_STAGE);
          bulk_copy_g2s(smem_u32(b_ring + (size_t)bs * B_STAGE), packed + (size_t)bstep * 2 * (B_TILE / 4), B_STAGE,
                        bar);
        }
      }
      __syncwarp();
    }
  } else {
    // ===================== MMA issuer =====================
    // The whole warp runs the (warp-uniform) loop so that descriptors and addresses stay in uniform
    // registers; one elected lane issues the tcgen05 instructions.  A comes from TMEM (TS mode).
    constexpr uint32_t idesc = make_idesc_tf32(kBM, COUT);
    // constant high word of the SW128 K-major descriptor: SBO = 1024 B, version 1, layout SWIZZLE_128B
    constexpr uint32_t desc_hi = 64u | (1u << 14) | (2u << 29);
    const uint32_t b_lo0 = ((smem_u32(b_ring) >> 4) & 0x3FFF) | (1u << 16);
    const uint32_t bar_a_full0 = smem_u32(bar_a_full), bar_a_empty0 = smem_u32(bar_a_empty);
    const uint32_t bar_b_full0 = smem_u32(bar_b_full), bar_b_empty0 = smem_u32(bar_b_empty);
    int stage = 0, bs = 0, t = 0;
    uint32_t a_phase = 0, b_phase = 0, first = 1;
    for (int s = 0; s < nsteps; ++s) {
      if (t == 0) mbar_wait(bar_b_full0 + 8 * bs, b_phase);
      mbar_wait(bar_a_full0 + 8 * stage, a_phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_hi = tmem_a0 + (uint32_t)(stage * A_COLS);
        const uint32_t b_lo = b_lo0 + (uint32_t)bs * (B_STAGE >> 4);
        const uint32_t d = tmem_base + (uint32_t)(t * COUT);
#pragma unroll
        for (int q = 0; q < kBK / 8; ++q) {   // UMMA K = 8 for TF32: 8 TMEM columns of A, 32 B of every B row
          const uint64_t db_hi = ((uint64_t)desc_hi << 32) | (b_lo + 2 * q);
          if constexpr (PASSES == 3) {
            const uint64_t db_lo = ((uint64_t)desc_hi << 32) | (b_lo + (B_TILE >> 4) + 2 * q);
            umma_tf32_ts(d, a_hi + kBK + 8 * q, db_hi, idesc, q == 0 ? (first ^ 1u) : 1u);   // small terms first
            umma_tf32_ts(d, a_hi + 8 * q, db_lo, idesc, 1);
            umma_tf32_ts(d, a_hi + 8 * q, db_hi, idesc, 1);
          } else {
            umma_tf32_ts(d, a_hi + 8 * q, db_hi, idesc, q == 0 ? (first ^ 1u) : 1u);
          }
        }
        umma_commit(bar_a_empty0 + 8 * stage);                  // gathered tile reusable once read
        if (t == T - 1) umma_commit(bar_b_empty0 + 8 * bs);     // weight tile reusable
        if (s == nsteps - 1) umma_commit(smem_u32(bar_accum));  // all accumulators complete
      }
      __syncwarp();
      if (++stage == SA) { stage = 0; a_phase ^= 1; }
      if (++t == T) {
        t = 0;
        first = 0;                                              // every tile has seen its first (k=0,c=0) MMA
        if (++bs == SB) { bs = 0; b_phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kLoaderWarp) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// weight packing: W [K, Cin, Cout] -> per (cout block, k, chunk): [hi tile | lo tile], each CB rows x 32 fp32 in
// the exact shared-memory image the kernel's B descriptor expects (K-major, 128B swizzle, K permuted like
// the TMEM A layout), TF32-rounded.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ W, int K, int Cin, int Cout,
                                                           int CB, float* __restrict__ packed) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)K * Cin * Cout;
  if (idx >= total) return;
  const int co = (int)(idx % Cout);
  const int ci = (int)((idx / Cout) % Cin);
  const int k = (int)(idx / ((long long)Cout * Cin));
  const float w = W[idx];
  const float hi = tf32_rna(w);
  const float lo = tf32_rna(w - hi);
  const int c = ci / kBK, kk = ci % kBK;
  const int nchunk = Cin / kBK;
  const int blk = co / CB, n = co % CB;
  const size_t tile = (size_t)CB * kBK;                         // floats per tile
  const size_t base = ((size_t)(blk * K + k) * nchunk + c) * 2 * tile;
  const int col = a_col_of_channel(kk);                         // K position the MMA sees (matches the TMEM A layout)
  const size_t pos = (sw128_chunk_offset(n, col >> 2) >> 2) + (col & 3);
  packed[base + pos] = hi;
  packed[base + tile + pos] = lo;
}

static int cout_block(int Cout) { return Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : 0)); }

static bool tc_supported(int Cin, int Cout) { return Cin >= kBK && Cin % kBK == 0 && cout_block(Cout) != 0; }

template <int COUT, int PASSES>
static int launch_tc(const ConvArgs& a, int Cout, cudaStream_t st) {
  using Cfg = TcCfg<COUT, PASSES>;
  static bool configured = false;
  if (!configured) {
    S2D_CUDA(cudaFuncSetAttribute(spconv_tc_kernel<COUT, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    configured = true;
  }
  const dim3 grid(div_up(a.n_out, Cfg::ROWS), Cout / COUT);
  spconv_tc_kernel<COUT, PASSES><<<grid, kTcThreads, Cfg::SMEM_BYTES, st>>>(a);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

int conv_fwd_tf32(const s2d_conv_params& p, cudaStream_t st) {
  if (!tc_supported(p.Cin, p.Cout)) {
    set_error("s2d_conv_fwd: no tcgen05 kernel for Cin=%d Cout=%d (need Cin %% 32 == 0 and Cout %% 32 == 0)", p.Cin,
              p.Cout);
    return S2D_ERR_UNSUPPORTED;
  }
  S2D_REQUIRE(p.in_ld % 4 == 0 && p.out_ld % 4 == 0 && (!p.residual || p.res_ld % 4 == 0),
              "s2d_conv_fwd: row strides must be multiples of 4 floats");
  ConvArgs a;
  a.in = p.in; a.packed = p.weights; a.tbl = p.tbl; a.scale = p.scale; a.shift = p.shift; a.residual = p.residual;
  a.out = p.out; a.out_rows = p.out_rows; a.in_ld = p.in_ld; a.out_ld = p.out_ld; a.res_ld = p.res_ld;
  a.tbl_stride = p.tbl_stride; a.n_out = p.n_out; a.K = p.K; a.nchunk = p.Cin / kBK; a.act = p.act;
  a.res_after_act = p.res_after_act;
  const int cb = cout_block(p.Cout);
  const bool x3 = p.precision == S2D_PRECISION_TF32X3;
  if (cb == 128) return x3 ? launch_tc<128, 3>(a, p.Cout, st) : launch_tc<128, 1>(a, p.Cout, st);
  if (cb == 64) return x3 ? launch_tc<64, 3>(a, p.Cout, st) : launch_tc<64, 1>(a, p.Cout, st);
  return x3 ? launch_tc<32, 3>(a, p.Cout, st) : launch_tc<32, 1>(a, p.Cout, st);
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_spconv_tf32_supported(int Cin, int Cout) { return tc_supported(Cin, Cout) ? 1 : 0; }

extern "C" size_t s2d_spconv_packed_bytes(int K, int Cin, int Cout) {
  if (K < 1 || !tc_supported(Cin, Cout)) return 0;
  return (size_t)K * Cin * Cout * 2 * sizeof(float);
}

extern "C" int s2d_spconv_pack_weights(const float* W, int K, int Cin, int Cout, float* packed, void* stream) {
  S2D_REQUIRE(W && packed && K >= 1 && K <= kTcMaxK, "s2d_spconv_pack_weights: bad argument");
  S2D_REQUIRE(tc_supported(Cin, Cout), "s2d_spconv_pack_weights: unsupported shape Cin=%d Cout=%d", Cin, Cout);
  pack_weights_kernel<<<div_up((long long)K * Cin * Cout, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      W, K, Cin, Cout, cout_block(Cout), packed);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
