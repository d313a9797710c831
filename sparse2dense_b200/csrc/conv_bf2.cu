// conv_bf2.cu -- the gather-GEMM of every convolution on the 5th-gen tensor cores, operands PRE-SPLIT into BF16 pairs.
//
//   out[o,:] = act( (sum_k in[tbl[k][o],:] @ W[k]) * scale + shift (+ residual[o,:]) )
//
// Why a second tcgen05 kernel (spconv_tc.cu is the first): there the producer warps convert every gathered fp32 row
// into tensor-core operands in registers (LDS -> split -> tcgen05.st), once per (row, kernel offset) = 27 times per
// row, and that instruction stream -- not the tensor pipe, not HBM -- bounded every layer at 750-970 clk per 128-row
// step.  Here the split is done ONCE per activation, by the epilogue of the layer that produces it:
//
//   x = hi + lo,  hi = bf16_rn(x),  lo = bf16_rn(x - hi)            (16 mantissa bits, stored as 2 x 16 bit = 4 B)
//   x*w ~= hi*w1 + hi*w2 + lo*w1                                     (three BF16 MMAs, fp32 accumulate in TMEM;
//                                                                     4e-6 relative per layer, 170x better than TF32)
//
// A "split row" keeps, per 32-channel chunk, [16 words hi | 16 words lo] (two BF16 per word), i.e. exactly the 128 B
// K-major shared-memory row the MMA wants: the gather is four cp.async per lane straight into the 128B-swizzled
// operand tile, and the tensor core reads A and B from shared memory (SS mode).  No register ever holds an operand.
// BF16 MMAs run at twice the TF32 rate, so the three products cost 1.5 TF32 passes (TF32x3 cost 3, TF32+BF16C 2).
//
// Executed work is cut by skipping (tile, kernel offset) pairs in which no row of the 128-row tile has a neighbour:
// the rulebook builders emit one 27-bit liveness mask per tile (rows are ordered by their neighbour pattern, see
// rulebook.cu), and all warp roles walk the same compacted step list.
//
//   warps 0-7   producers: warp w gathers rows [16w, 16w+16) of every step (cp.async 16 B, zero-fill for a missing
//               neighbour); the stage's mbarrier is armed with cp.async.mbarrier.arrive.noinc, so it completes when the
//               copies of all 256 lanes have landed and the producers never wait for their own data (the CUTLASS sm100
//               cp.async mainloop synchronises cp.async -> UMMA the same way); afterwards the same warps run the epilogue (tcgen05.ld -> BN affine / residual / activation
//               -> fp32 row and, optionally, the split row the next layer gathers from).
//   warp 8      weight tiles: one cp.async.bulk per live (offset step, channel chunk); owns TMEM alloc and builds the
//               step list.
//   warp 9      MMA issuer: one thread, six tcgen05.mma.kind::f16 per step (M = 128, N = COUT, K = 16), tcgen05.commit
//               releases the rings / publishes the accumulators.  A single in-order issuer sees every mbarrier phase,
//               so the dynamic step list needs no phase-aliasing rules.
#include "tc_ptx.cuh"

namespace s2d {

constexpr int kB2ProducerWarps = 8;
constexpr int kB2LoaderWarp = 8;
constexpr int kB2MmaWarp = 9;
constexpr int kB2Threads = 32 * 10;
constexpr int kB2MaxSteps = 2048;          // (offset steps) x (chunks) x (tiles) of one CTA; host-checked
constexpr int kB2AStage = kBM * 128;       // 128 rows x 128 B

// D[tmem] (+)= A[smem] * B[smem], BF16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int COUT, int T_, int SA_, int SB_>
struct B2Cfg {
  static constexpr int T = T_, SA = SA_, SB = SB_;
  static constexpr int B_STAGE = COUT * 128;                   // COUT rows x [w1 (64 B) | w2 (64 B)]
  static constexpr int ACC_STRIDE = COUT < 32 ? 32 : COUT;
  static constexpr int ACC_COLS = T * ACC_STRIDE;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static constexpr int EPC = COUT < 32 ? COUT : 32;
  static constexpr int STEP_BYTES = kB2MaxSteps * 2;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = SA * kB2AStage + SB * B_STAGE + STEP_BYTES + BAR_BYTES + 1024;
  static constexpr int OCC = (2 * (SMEM_BYTES + 1024) <= 227 * 1024 + 1024 && 2 * TMEM_COLS <= 512) ? 2 : 1;
  static_assert(ACC_COLS <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(COUT % 16 == 0 && COUT >= 16 && COUT <= 128, "UMMA N constraint for M = 128");
  static_assert(2 * SA + 2 * SB + 1 <= BAR_BYTES / 8 - 2, "barrier block");
};

struct B2Args {
  const uint32_t* in;        // split rows [n_in, in_ld] (words)
  const uint32_t* packed;    // weight image: per (COUT block, offset step, chunk) one [COUT x 128 B] SW128 tile
  const int* tbl;            // [K, tbl_stride]
  const int* tile_masks;     // [n_tiles] live-offset masks (bit k: some row of the tile has a neighbour at offset k) or null
  const float* scale;
  const float* shift;
  const float* residual;
  float* out;                // fp32 rows or null
  uint32_t* out_split;       // split rows or null
  const int* out_rows;
  int in_ld, out_ld, res_ld, split_ld, tbl_stride, n_out, K, nchunk, kps, ksteps, act, res_after_act;
  int tile_unit, unit_base, unit_rem, n_tiles;
  int dbg;   // ablation switches (tools/microbench_bf2.py): 1 no gather, 2 no MMA, 4 no weight copies, 8 no stores, 16 all rows missing
};

__device__ __forceinline__ float b2_act(float y, int act) {
  if (act == S2D_ACT_RELU) return fmaxf(y, 0.f);
  if (act == S2D_ACT_GELU) return 0.5f * y * (1.f + erff(y * 0.70710678118654752440f));
  return y;
}

// x -> (bf16_rn(x), bf16_rn(x - hi)) for two values, packed [lo half = first value]
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ra = a - __uint_as_float(hi << 16);
  const float rb = b - __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int COUT, int T, int SA, int SB>
__global__ void __launch_bounds__(kB2Threads, B2Cfg<COUT, T, SA, SB>::OCC)
conv_bf2_kernel(const __grid_constant__ B2Args A) {
  using Cfg = B2Cfg<COUT, T, SA, SB>;
  constexpr int B_STAGE = Cfg::B_STAGE;
  const int NCHUNK = A.nchunk, K = A.K, KS = A.ksteps, kps = A.kps, n_out = A.n_out;
  const int cblk = blockIdx.y * COUT;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;                                        // SA x [128 rows x 128 B], 128B-swizzled
  uint8_t* b_ring = a_ring + SA * kB2AStage;                     // SB x [COUT rows x 128 B], 128B-swizzled
  uint16_t* steps = reinterpret_cast<uint16_t*>(b_ring + SB * B_STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(steps) + Cfg::STEP_BYTES);
  uint64_t* bar_a_full = bars;
  uint64_t* bar_a_empty = bar_a_full + SA;
  uint64_t* bar_b_full = bar_a_empty + SA;
  uint64_t* bar_b_empty = bar_b_full + SB;
  uint64_t* bar_accum = bar_b_empty + SB;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_accum + 1);
  uint32_t* s_nsteps = s_tmem + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bx = (int)blockIdx.x;
  const int t_alloc = (A.unit_base + (bx < A.unit_rem ? 1 : 0)) * A.tile_unit;
  const int tile_first = (bx * A.unit_base + min(bx, A.unit_rem)) * A.tile_unit;
  if (t_alloc == 0 || tile_first >= A.n_tiles) return;
  const int Tr = min(t_alloc, A.n_tiles - tile_first);
  const int tile0 = tile_first * kBM;
  const int row_end = min(n_out, tile0 + Tr * kBM);

  if (warp == kB2MmaWarp && lane == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(smem_u32(bar_a_full + s), kB2ProducerWarps * 32);   // every producer lane: cp.async.mbarrier.arrive.noinc
      mbar_init(smem_u32(bar_a_empty + s), 1);                    // one tcgen05.commit
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(smem_u32(bar_b_full + s), 1);                     // arrive.expect_tx of the loader
      mbar_init(smem_u32(bar_b_empty + s), 1);                    // one tcgen05.commit after the last tile of the stage
    }
    mbar_init(smem_u32(bar_accum), 1);
    fence_barrier_init();
  }
  if (warp == kB2LoaderWarp) {
    tmem_alloc(smem_u32(s_tmem), Cfg::TMEM_COLS);
    // ---- step list: lane = offset step kk; entries ordered (kk, chunk, tile), tiles without a neighbour at kk skipped ----
    // entry = kk | chunk << 5 | tile << 10 | first-of-(kk,chunk) << 12 | last-of-(kk,chunk) << 13
    uint32_t nib = 0;
    uint32_t tmask[T];
#pragma unroll
    for (int t = 0; t < T; ++t)                                   // all loads in flight together
      tmask[t] = (A.tile_masks && t < Tr) ? (uint32_t)__ldg(A.tile_masks + tile_first + t) : 0xffffffffu;
    if (lane < KS) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        if (t >= Tr) break;
        uint32_t m = tmask[t];
        m &= K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
        if (m == 0) m = 1;                                        // a tile always runs at least one step (zero rows)
        const uint32_t live = kps == 1 ? (m >> lane) & 1u : ((m >> (2 * lane)) & 3u) != 0;
        nib |= live << t;
      }
    }
    const int cnt = NCHUNK * __popc(nib);
    const int off = warp_inclusive_scan(cnt) - cnt;
    int w = off;
    if (nib) {
      const int t_first = __ffs(nib) - 1, t_last = 31 - __clz(nib);
      for (int c = 0; c < NCHUNK; ++c)
        for (int t = t_first; t <= t_last; ++t)
          if ((nib >> t) & 1u) {
            if (w < kB2MaxSteps)
              steps[w] = (uint16_t)(lane | (c << 5) | (t << 10) | ((t == t_first) << 12) | ((t == t_last) << 13));
            ++w;
          }
    }
    const int total = __shfl_sync(0xffffffffu, off + cnt, 31);
    if (lane == 0) *s_nsteps = (uint32_t)min(total, kB2MaxSteps);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int nsteps = (int)*s_nsteps;

  if (warp < kB2ProducerWarps) {
    // ===================== producers: warp w gathers rows [16w, 16w + 16) of every step =====================
    const int o = lane >> 3, c8 = lane & 7;                  // lane = (row quad, 16 B piece of the 128 B operand row)
    const int rbase = 16 * warp + 4 * o;                     // this lane's rows inside a tile: rbase .. rbase + 3
    const int sub = kps == 2 ? (c8 >> 2) : 0;                // Cin = 16: pieces 0-3 come from offset 2kk, 4-7 from 2kk + 1
    const uint32_t piece = (uint32_t)(kps == 2 ? (c8 & 3) : c8) * 16u;
    const uint32_t dst_lane = (uint32_t)rbase * 128u + (uint32_t)((c8 ^ (4 * (o & 1))) << 4);   // (rbase + i) & 7 = 4(o&1) + i
    const char* in_bytes = reinterpret_cast<const char*>(A.in);
    const uint32_t row_bytes = (uint32_t)A.in_ld * 4u;
    const int* tbl = A.tbl;
    const int tbl_stride = A.tbl_stride;
    const uint32_t a_ring0 = smem_u32(a_ring);
    const uint32_t bar_full0 = smem_u32(bar_a_full), bar_empty0 = smem_u32(bar_a_empty);

    // Step entries and neighbour indices run three steps ahead of the copies in registers.  The entries are read with
    // ld.shared (a generic load of the list would queue behind the global traffic), the indices raw: rows past the end
    // of the tensor are masked at use, so nothing depends on the index load until the step is issued.
    const uint32_t steps0 = smem_u32(steps);
    auto load_entry = [&](int i) -> uint32_t { return i < nsteps ? lds_u16(steps0 + 2u * (uint32_t)i) : 0xffffffffu; };
    auto load_idx = [&](uint32_t e) -> int4 {
      int4 v = make_int4(-1, -1, -1, -1);
      if (e != 0xffffffffu && !(A.dbg & 64)) {
        const int k = (int)(e & 31u) * kps + sub;
        const int row0 = tile0 + (int)((e >> 10) & 3u) * kBM + rbase;
        if (k < K && row0 < row_end) v = __ldg(reinterpret_cast<const int4*>(tbl + (size_t)k * tbl_stride + row0));
      }
      return v;
    };

    uint32_t e0 = load_entry(0), e1 = load_entry(1), e2 = load_entry(2);
    int4 q0 = load_idx(e0), q1 = load_idx(e1), q2 = load_idx(e2);
    int stage = 0;
    uint32_t phase = 1;                                      // first pass over the ring: stages are free
#pragma unroll 1
    for (int i = 0; i < nsteps; ++i) {
      int4 cur = q0;
      const uint32_t e = e0;
      e0 = e1; e1 = e2; e2 = load_entry(i + 3);
      q0 = q1; q1 = q2; q2 = load_idx(e2);
      const uint32_t cbytes = (kps == 2 ? 0u : ((e >> 5) & 31u) * 128u) + piece;
      const int lim = row_end - (tile0 + (int)((e >> 10) & 3u) * kBM + rbase);   // rows of this quad inside the tensor
      if (lim < 4) { if (lim < 2) cur.y = -1; if (lim < 3) cur.z = -1; cur.w = -1; }
      mbar_wait(bar_empty0 + 8 * stage, phase);
      const uint32_t dst0 = a_ring0 + (uint32_t)stage * kB2AStage + dst_lane;
      const int idx[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (A.dbg & 1) break;
        const uint32_t off = (uint32_t)max(idx[r], 0) * row_bytes + cbytes;
        cp_async16_zfill((dst0 + (uint32_t)r * 128u) ^ ((uint32_t)r << 4), in_bytes + off,
                         (idx[r] >= 0 && !(A.dbg & 16)) ? 16u : 0u);
      }
      // the stage's full barrier gets this thread's arrival when its copies have landed (no wait in the producer)
      cp_async_mbar_arrive_noinc(bar_full0 + 8 * stage);
      if (++stage == SA) { stage = 0; phase ^= 1; }
    }

    // ===================== epilogue (same 8 warps) =====================
    // TMEM lane == row inside a tile; warp w may touch lanes [32*(w%4), +32); warps 0-3 take even tiles, 4-7 odd ones
    mbar_wait(smem_u32(bar_accum), 0);
    tc_fence_after();
    const int g = warp & 3;
    const float* __restrict__ scale = A.scale ? A.scale + cblk : nullptr;
    const float* __restrict__ shift = A.shift ? A.shift + cblk : nullptr;
    const int act = A.act, res_after = A.res_after_act;
    constexpr int EPC = Cfg::EPC;
#pragma unroll 1
    for (int t = warp >> 2; t < Tr; t += 2) {
      if (A.dbg & 128) break;
      const int row = tile0 + t * kBM + g * 32 + lane;
      const int orow = (row < row_end && A.out_rows) ? __ldg(A.out_rows + row) : row;
#pragma unroll 1
      for (int c0 = 0; c0 < COUT; c0 += EPC) {
        uint32_t acc[EPC];
        tmem_ld<EPC>(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(t * Cfg::ACC_STRIDE + c0), acc);
        if (row < row_end && !(A.dbg & 8)) {
          const float* res = A.residual ? A.residual + (size_t)orow * A.res_ld + cblk + c0 : nullptr;
          float y[EPC];
#pragma unroll
          for (int q = 0; q < EPC / 4; ++q) {
            float4 v;
            v.x = __uint_as_float(acc[4 * q + 0]); v.y = __uint_as_float(acc[4 * q + 1]);
            v.z = __uint_as_float(acc[4 * q + 2]); v.w = __uint_as_float(acc[4 * q + 3]);
            if (scale) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c0) + q);
              v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
            }
            if (shift) {
              const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c0) + q);
              v.x += sh.x; v.y += sh.y; v.z += sh.z; v.w += sh.w;
            }
            float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
            if (res) rr = __ldg(reinterpret_cast<const float4*>(res) + q);
            if (!res_after) { v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w; }
            v.x = b2_act(v.x, act); v.y = b2_act(v.y, act); v.z = b2_act(v.z, act); v.w = b2_act(v.w, act);
            if (res_after) { v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w; }
            y[4 * q + 0] = v.x; y[4 * q + 1] = v.y; y[4 * q + 2] = v.z; y[4 * q + 3] = v.w;
          }
          if (A.out) {
            float4* dst = reinterpret_cast<float4*>(A.out + (size_t)orow * A.out_ld + cblk + c0);
#pragma unroll
            for (int q = 0; q < EPC / 4; ++q) dst[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
          }
          if (A.out_split) {
            // split row: per chunk of EPC channels [EPC/2 words hi | EPC/2 words lo]
            uint32_t hi[EPC / 2], lo[EPC / 2];
#pragma unroll
            for (int q = 0; q < EPC / 2; ++q) split_pair(y[2 * q], y[2 * q + 1], hi[q], lo[q]);
            uint4* dst = reinterpret_cast<uint4*>(A.out_split + (size_t)orow * A.split_ld + cblk + c0);
#pragma unroll
            for (int q = 0; q < EPC / 8; ++q) {
              dst[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
              dst[EPC / 8 + q] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
            }
          }
        }
      }
    }
  } else if (warp == kB2LoaderWarp) {
    // ===================== weight tiles: one bulk copy per live (offset step, chunk) =====================
    // All lanes: pull the neighbour-table lines of the steps kIdxAhead ahead into L2 (the table is 27 x 4 B per row,
    // larger than the feature maps, and streams from DRAM; the producers' index loads then hit L2 and their
    // three-step register prefetch covers the latency).  Lane 0 also feeds the weight ring.
    constexpr int kIdxAhead = 24;
    const uint8_t* packed = reinterpret_cast<const uint8_t*>(A.packed) + (size_t)blockIdx.y * KS * NCHUNK * B_STAGE;
    const uint32_t steps0 = smem_u32(steps);
    const int* tbl = A.tbl;
    const int tbl_stride = A.tbl_stride;
    auto prefetch_idx = [&](int i) {
      if (i >= nsteps || lane >= 4 * kps || (A.dbg & 64)) return;
      const uint32_t e = lds_u16(steps0 + 2u * (uint32_t)i);
      if (((e >> 5) & 31u) != 0u) return;                      // all chunks of an (offset, tile) use the same indices
      const int k = (int)(e & 31u) * kps + (lane >> 2);
      const int row = tile0 + (int)((e >> 10) & 3u) * kBM + 32 * (lane & 3);     // 128 rows x 4 B = four 128 B lines
      if (k < K && row < row_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(tbl + (size_t)k * tbl_stride + row));
    };
    for (int i = 0; i < kIdxAhead; ++i) prefetch_idx(i);
    int bs = 0;
    uint32_t phase = 1;
    for (int i = 0; i < nsteps; ++i) {
      prefetch_idx(i + kIdxAhead);
      const uint32_t e = lds_u16(steps0 + 2u * (uint32_t)i);
      if (!((e >> 12) & 1u)) continue;
      if (lane == 0) {
        const int bstep = (int)(e & 31u) * NCHUNK + (int)((e >> 5) & 31u);
        mbar_wait(smem_u32(bar_b_empty + bs), phase);
        const uint32_t bar = smem_u32(bar_b_full + bs);
        if (A.dbg & 4) {
          mbar_arrive(bar);
        } else {
          mbar_arrive_expect_tx(bar, B_STAGE);
          bulk_copy_g2s(smem_u32(b_ring + (size_t)bs * B_STAGE), packed + (size_t)bstep * B_STAGE, B_STAGE, bar);
        }
      }
      if (++bs == SB) { bs = 0; phase ^= 1; }
      __syncwarp();
    }
  } else if (warp == kB2MmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, COUT);
      constexpr uint32_t desc_hi = 64u | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
      const uint32_t a_lo0 = ((smem_u32(a_ring) >> 4) & 0x3FFF) | (1u << 16);
      const uint32_t b_lo0 = ((smem_u32(b_ring) >> 4) & 0x3FFF) | (1u << 16);
      const uint32_t bar_a_full0 = smem_u32(bar_a_full), bar_a_empty0 = smem_u32(bar_a_empty);
      const uint32_t bar_b_full0 = smem_u32(bar_b_full), bar_b_empty0 = smem_u32(bar_b_empty);
      // (A slice, B slice) of the six MMAs of a step, small terms first; a slice = 32 B = 16 BF16 k positions.
      //   Cin >= 32: A = [hi 0-15 | hi 16-31 | lo 0-15 | lo 16-31], B = [w1 0-15 | w1 16-31 | w2 0-15 | w2 16-31]
      //   Cin == 16: A = [hi k0 | lo k0 | hi k1 | lo k1],            B = [w1 k0 | w2 k0 | w1 k1 | w2 k1]
      const uint32_t pa = kps == 1 ? ((2u) | (3u << 2) | (0u << 4) | (1u << 6) | (0u << 8) | (1u << 10))
                                   : ((1u) | (0u << 2) | (3u << 4) | (2u << 6) | (0u << 8) | (2u << 10));
      const uint32_t pb = kps == 1 ? ((0u) | (1u << 2) | (2u << 4) | (3u << 6) | (0u << 8) | (1u << 10))
                                   : ((0u) | (1u << 2) | (2u << 4) | (3u << 6) | (0u << 8) | (2u << 10));
      int as = 0, bs = -1;
      uint32_t a_phase = 0, b_phase = 1;
      uint32_t started = 0;
      const uint32_t steps0 = smem_u32(steps);
      uint32_t e_next = lds_u16(steps0);
      for (int i = 0; i < nsteps; ++i) {
        const uint32_t e = e_next;
        e_next = lds_u16(steps0 + 2u * (uint32_t)min(i + 1, nsteps - 1));
        const int t = (int)((e >> 10) & 3u);
        if ((e >> 12) & 1u) {                                     // first tile of a new weight stage
          if (++bs == SB) bs = 0;
          if (bs == 0) b_phase ^= 1;
          mbar_wait(bar_b_full0 + 8 * bs, b_phase);
        }
        mbar_wait(bar_a_full0 + 8 * as, a_phase);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(t * Cfg::ACC_STRIDE);
        const uint32_t a_lo = a_lo0 + (uint32_t)as * (kB2AStage >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)bs * (B_STAGE >> 4);
        uint32_t acc = (started >> t) & 1u;
        started |= 1u << t;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          if (A.dbg & 2) break;
          const uint64_t da = ((uint64_t)desc_hi << 32) | (a_lo + 2u * ((pa >> (2 * q)) & 3u));
          const uint64_t db = ((uint64_t)desc_hi << 32) | (b_lo + 2u * ((pb >> (2 * q)) & 3u));
          umma_bf16_ss(d, da, db, idesc, acc);
          acc = 1u;
        }
        if (A.dbg & 32) {                                         // ablation (with 2): plain arrivals instead of commits
          mbar_arrive(bar_a_empty0 + 8 * as);
          if ((e >> 13) & 1u) mbar_arrive(bar_b_empty0 + 8 * bs);
        } else {
          umma_commit(bar_a_empty0 + 8 * as);                       // gathered tile reusable once read
          if ((e >> 13) & 1u) umma_commit(bar_b_empty0 + 8 * bs);   // weight tile: last tile of the stage
        }
        if (++as == SA) { as = 0; a_phase ^= 1; }
      }
      umma_commit(smem_u32(bar_accum));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kB2LoaderWarp) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// weight packing: W [K, Cin, Cout] -> per (cout block, contraction step) one [CB rows x 128 B] tile in the exact
// shared-memory image of the B descriptor (K-major, 128B swizzle): row n = [w1 (32 BF16) | w2 (32 BF16)] for a
// 32-channel chunk, or [w1 k0 | w2 k0 | w1 k1 | w2 k1] (16 BF16 each) when two offsets of a 16-channel input share a step.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_weights_bf2_kernel(const float* __restrict__ W, int K, int Cin, int Cout, int CB,
                                                               int kps, int nchunk, int ksteps,
                                                               __nv_bfloat16* __restrict__ packed) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nstep = ksteps * nchunk;
  const long long total = (long long)(Cout / CB) * nstep * CB * kBK;
  if (idx >= total) return;
  const int kk32 = (int)(idx % kBK);
  const int n = (int)((idx / kBK) % CB);
  const int step = (int)((idx / ((long long)kBK * CB)) % nstep);
  const int blk = (int)(idx / ((long long)kBK * CB * nstep));
  int k, ci, pos1, pos2;                                           // BF16 positions of w1 / w2 inside the 64-element row
  if (kps == 1) {
    k = step / nchunk;
    ci = (step % nchunk) * kBK + kk32;
    pos1 = kk32; pos2 = 32 + kk32;
  } else {
    const int s = kk32 / 16, c = kk32 % 16;
    k = step * 2 + s;
    ci = c;
    pos1 = 32 * s + c; pos2 = 32 * s + 16 + c;
  }
  const int co = blk * CB + n;
  const float w = k < K ? W[((size_t)k * Cin + ci) * Cout + co] : 0.f;
  const __nv_bfloat16 w1 = __float2bfloat16_rn(w);
  const __nv_bfloat16 w2 = __float2bfloat16_rn(w - __bfloat162float(w1));
  __nv_bfloat16* tile = packed + ((size_t)blk * nstep + step) * (size_t)CB * 64;
  tile[(sw128_chunk_offset(n, pos1 >> 3) >> 1) + (pos1 & 7)] = w1;
  tile[(sw128_chunk_offset(n, pos2 >> 3) >> 1) + (pos2 & 7)] = w2;
}

// fp32 rows -> split rows (same geometry: word offset = channel offset; per chunk [hi | lo]); chunk = 32 channels, or
// the whole row when it has 16 channels.  One thread converts 8 channels (two float4 in, two uint4 out).
__global__ void __launch_bounds__(256) rows_split_kernel(const float* __restrict__ in, long long n, int C, int in_ld,
                                                         uint32_t* __restrict__ out, int out_ld) {
  const int per_row = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * per_row) return;
  const long long row = idx / per_row;
  const int j = (int)(idx % per_row);
  const int chunk = C == 16 ? 16 : 32;
  const int c0 = j * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(in + row * in_ld + c0));
  const float4 b = __ldg(reinterpret_cast<const float4*>(in + row * in_ld + c0 + 4));
  uint32_t hi[4], lo[4];
  split_pair(a.x, a.y, hi[0], lo[0]); split_pair(a.z, a.w, hi[1], lo[1]);
  split_pair(b.x, b.y, hi[2], lo[2]); split_pair(b.z, b.w, hi[3], lo[3]);
  const int cb = c0 / chunk * chunk, within = (c0 % chunk) / 2;
  uint32_t* dst = out + row * out_ld + cb;
  *reinterpret_cast<uint4*>(dst + within) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(dst + chunk / 2 + within) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// tile_masks[t] bit k <=> some row of tile t has a neighbour at offset k.  One block (128 threads) per tile.
__global__ void __launch_bounds__(128) tile_masks_kernel(const int* __restrict__ tbl, int stride, int K, int n,
                                                         int* __restrict__ masks) {
  __shared__ unsigned s_mask;
  if (threadIdx.x == 0) s_mask = 0u;
  __syncthreads();
  const int row = blockIdx.x * 128 + threadIdx.x;
  unsigned m = 0;
  for (int k = 0; k < K; ++k) {
    const bool hit = row < n && __ldg(tbl + (size_t)k * stride + row) >= 0;
    if (__ballot_sync(0xffffffffu, hit)) m |= 1u << k;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicOr(&s_mask, m);
  __syncthreads();
  if (threadIdx.x == 0) masks[blockIdx.x] = (int)s_mask;
}

static int g_b2_variant = 0;
static int g_b2_dbg = 0;

static int b2_cout_block(int Cout) {
  return Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : (Cout % 16 == 0 ? 16 : 0)));
}
bool bf2_supported(int Cin, int Cout) { return (Cin == 16 || (Cin >= 32 && Cin % 32 == 0)) && b2_cout_block(Cout) != 0; }

template <int COUT, int T, int SA, int SB>
static int launch_b2(const B2Args& a, int Cout, cudaStream_t st) {
  using Cfg = B2Cfg<COUT, T, SA, SB>;
  static bool configured = false;
  if (!configured) {
    S2D_CUDA(cudaFuncSetAttribute(conv_bf2_kernel<COUT, T, SA, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    configured = true;
  }
  B2Args b = a;
  const int n_tiles = div_up(a.n_out, kBM);
  S2D_REQUIRE(a.ksteps * a.nchunk * T <= kB2MaxSteps, "s2d_conv_fwd(bf16x2): %d x %d x %d contraction steps per CTA exceed %d",
              a.ksteps, a.nchunk, T, kB2MaxSteps);
  // Deal the 128-row tiles over a grid that is a whole number of waves (OCC CTAs per SM), at most T per CTA
  const int wave = kNumSMs * Cfg::OCC;
  const int g_full = div_up(n_tiles, T);
  int gx = div_up(g_full, wave) * wave;
  if (gx > n_tiles) gx = n_tiles;
  b.tile_unit = 1;
  b.unit_base = n_tiles / gx;
  b.unit_rem = n_tiles % gx;
  b.n_tiles = n_tiles;
  const dim3 grid(gx, Cout / COUT);
  conv_bf2_kernel<COUT, T, SA, SB><<<grid, kB2Threads, Cfg::SMEM_BYTES, st>>>(b);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

int conv_fwd_bf2(const s2d_conv_params& p, cudaStream_t st) {
  if (!bf2_supported(p.Cin, p.Cout)) {
    set_error("s2d_conv_fwd: no bf16x2 kernel for Cin=%d Cout=%d (need Cin == 16 or Cin %% 32 == 0, Cout %% 16 == 0)", p.Cin,
              p.Cout);
    return S2D_ERR_UNSUPPORTED;
  }
  S2D_REQUIRE(p.in_split, "s2d_conv_fwd(bf16x2): in_split (split rows, s2d_rows_split) is required");
  S2D_REQUIRE(p.out || p.out_split, "s2d_conv_fwd(bf16x2): no output");
  S2D_REQUIRE(p.in_split_ld % 4 == 0 && p.in_split_ld >= p.Cin && (reinterpret_cast<uintptr_t>(p.in_split) & 15) == 0,
              "s2d_conv_fwd(bf16x2): split rows must be 16 B aligned with a row stride that is a multiple of 4 words");
  S2D_REQUIRE(!p.out || p.out_ld % 4 == 0, "s2d_conv_fwd: row strides must be multiples of 4 floats");
  S2D_REQUIRE(!p.residual || p.res_ld % 4 == 0, "s2d_conv_fwd: row strides must be multiples of 4 floats");
  S2D_REQUIRE(p.tbl_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(p.tbl) & 15) == 0,
              "s2d_conv_fwd(bf16x2): the neighbour table must be 16 B aligned with tbl_stride %% 4 == 0 (got %d)", p.tbl_stride);
  S2D_REQUIRE((unsigned long long)p.n_in * (unsigned long long)p.in_split_ld * 4ull < (1ull << 32),
              "s2d_conv_fwd: input tensor larger than 4 GiB (32-bit gather offsets)");
  const int cb = b2_cout_block(p.Cout);
  if (p.out_split) {
    S2D_REQUIRE(p.out_split_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(p.out_split) & 15) == 0 &&
                    (cb >= 32 || (p.Cout == 16 && p.out_split_ld == 16)),
                "s2d_conv_fwd(bf16x2): out_split needs 32-channel output blocks (or a 16-channel row)");
  }
  B2Args a;
  a.in = static_cast<const uint32_t*>(p.in_split); a.packed = reinterpret_cast<const uint32_t*>(p.weights); a.tbl = p.tbl;
  a.tile_masks = p.tile_masks; a.scale = p.scale; a.shift = p.shift; a.residual = p.residual; a.out = p.out;
  a.out_split = static_cast<uint32_t*>(p.out_split); a.out_rows = p.out_rows; a.in_ld = p.in_split_ld; a.out_ld = p.out_ld;
  a.res_ld = p.res_ld; a.split_ld = p.out_split_ld; a.tbl_stride = p.tbl_stride; a.n_out = p.n_out; a.K = p.K;
  a.kps = p.Cin == 16 ? 2 : 1; a.nchunk = a.kps == 1 ? p.Cin / 32 : 1; a.ksteps = div_up(p.K, a.kps); a.act = p.act;
  a.res_after_act = p.res_after_act;
  a.dbg = g_b2_dbg;
  const int v = g_b2_variant;
  if (cb == 128) {
    if (v == 1) return launch_b2<128, 2, 3, 3>(a, p.Cout, st);
    if (v == 2) return launch_b2<128, 4, 6, 4>(a, p.Cout, st);
    return launch_b2<128, 2, 4, 2>(a, p.Cout, st);
  }
  if (cb == 64) {
    if (v == 1) return launch_b2<64, 4, 5, 3>(a, p.Cout, st);
    if (v == 2) return launch_b2<64, 4, 8, 4>(a, p.Cout, st);
    return launch_b2<64, 4, 4, 4>(a, p.Cout, st);
  }
  if (cb == 32) {
    if (v == 1) return launch_b2<32, 4, 4, 4>(a, p.Cout, st);
    if (v == 2) return launch_b2<32, 4, 8, 4>(a, p.Cout, st);
    return launch_b2<32, 4, 5, 4>(a, p.Cout, st);
  }
  if (v == 1) return launch_b2<16, 4, 4, 4>(a, p.Cout, st);
  if (v == 2) return launch_b2<16, 4, 8, 4>(a, p.Cout, st);
  return launch_b2<16, 4, 6, 4>(a, p.Cout, st);
}

int pack_weights_bf2(const float* W, int K, int Cin, int Cout, void* packed, cudaStream_t st) {
  const int kps = Cin == 16 ? 2 : 1;
  const int nchunk = kps == 1 ? Cin / 32 : 1;
  const int ksteps = div_up(K, kps);
  const long long total = (long long)ksteps * nchunk * kBK * Cout;
  pack_weights_bf2_kernel<<<div_up(total, 256), 256, 0, st>>>(W, K, Cin, Cout, b2_cout_block(Cout), kps, nchunk, ksteps,
                                                              static_cast<__nv_bfloat16*>(packed));
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

}  // namespace s2d

using namespace s2d;

// which (tiles per CTA, A stages, B stages) instantiation conv_fwd_bf2 launches (tuning aid; not part of the public header)
extern "C" void s2d_debug_bf2_variant(int v) { g_b2_variant = v; }
extern "C" void s2d_debug_bf2_flags(int f) { g_b2_dbg = f; }

extern "C" int s2d_table_tile_masks(const int* tbl, int tbl_stride, int K, int n_rows, int* tile_masks, void* stream) {
  S2D_REQUIRE(K >= 1 && K <= 31 && n_rows >= 0 && tbl_stride >= n_rows, "s2d_table_tile_masks: bad argument");
  if (n_rows == 0) return S2D_OK;
  S2D_REQUIRE(tbl && tile_masks, "s2d_table_tile_masks: null argument");
  tile_masks_kernel<<<div_up(n_rows, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(tbl, tbl_stride, K, n_rows, tile_masks);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_rows_split(const float* in, long long n_rows, int C, int in_ld, void* out, int out_ld, void* stream) {
  S2D_REQUIRE(n_rows >= 0 && C >= 16 && (C == 16 || C % 32 == 0), "s2d_rows_split: C = %d must be 16 or a multiple of 32", C);
  S2D_REQUIRE(in_ld >= C && out_ld >= C && in_ld % 4 == 0 && out_ld % 4 == 0, "s2d_rows_split: bad row stride");
  if (n_rows == 0) return S2D_OK;
  S2D_REQUIRE(in && out && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "s2d_rows_split: null or unaligned argument");
  const long long total = n_rows * (C / 8);
  rows_split_kernel<<<div_up(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, n_rows, C, in_ld,
                                                                                       static_cast<uint32_t*>(out), out_ld);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
