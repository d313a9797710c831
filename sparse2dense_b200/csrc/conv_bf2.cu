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
// Executed work is cut by skipping (tile, kernel offset) pairs in which no row of the 128-row tile has a neighbour: the
// rulebook rows are GROUPED by neighbour pattern (see "Row grouping" below; the launch scatters through out_rows, results are
// bit-identical), the builders emit one 27-bit liveness mask per tile, and all warp roles walk the same compacted block list
// (chunk-major: the offsets of one 32-channel chunk back to back, so gathered pieces are re-read from L2).
//
//   warps 0-7   producers: warp w fills stage w (one whole 128-row x 128 B tile per block: cp.async 16 B per lane straight
//               into the 128B-swizzled operand tile, zero-fill for a missing neighbour), waits for its own copies, fences
//               them to the async proxy and arrives on the stage's mbarrier.
//               Dense-grid mode (template flag TMA, s2d_conv_fwd_grid): one thread issues a 4-D TMA box per (block, tile)
//               instead -- a regular BEV map needs no neighbour table, the padding is the TMA zero fill.
//   warp 8      utility: block lists (double buffered), one cp.async.bulk per weight tile, L2 prefetch of table lines; owns
//               the TMEM allocation.
//   warps 9..   MMA issuers.  Cout <= 64 (and 1x1 layers): one per tile of the group, six tcgen05.mma.kind::f16 per block
//               (M = 128, N = COUT, K = 16).  Cout = 128: operands SWAPPED -- the weight tile is the M = 128 operand, the
//               gathered tiles of both tiles of the group one N = 256 operand, the accumulator is transposed -- and two
//               warps that issue ALTERNATE blocks (see there).  tcgen05.commit releases the rings / publishes the accumulators.
//   last 8      epilogue warps (two per TMEM lane quarter) on the second accumulator buffer: tcgen05.ld -> BN affine /
//               residual / ReLU / GELU -> fp32 row and split row, staged through shared memory for whole-row stores.
//
// Tile groups are handed out dynamically for masked launches (heaviest neighbour patterns first, grouping.cuh), statically
// otherwise.
//
// What bounds it (measured with tools/mma_probe.cu / sync_probe.cu / l2_probe.cu): the tensor pipe queues only
// ~1.5 MMAs, so the per-block overhead of the issuing warp used to idle it (fixed for Cout = 128 by the alternating
// issuers); what remains is the stream of gathered rows and weight tiles from L2 into the SM (~50 B/clk/SM for maps of
// 100-240 MB), 27 re-reads of every input row.
#include <string.h>

#include <atomic>

#include "grouping.cuh"
#include "tc_ptx.cuh"

namespace s2d {

// S gathered-tile stages, one producer warp each (warps 0..S-1); warp S = utility; warps S+1..S+T = MMA (one per tile of the
// group); then eight epilogue warps.  S = 8: one CTA per SM; S = 4: two CTAs per SM (half the ring each), whose independent
// pipelines hide each other's barrier round trips.
constexpr int kB2ListCap = 1024;           // (offset steps) x (chunks) of one tile group; host-checked
constexpr int kB2AStage = kBM * 128;       // 128 rows x 128 B
constexpr int kB2EpiRow = 144;             // staged accumulator row: 128 B + 16 B pad (conflict-free 16 B accesses)

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
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, int v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// A tile GROUP = T consecutive 128-row tiles (the unit a CTA works on); a BLOCK = one (offset step, channel chunk) of a
// group: one weight tile and up to T gathered tiles (one per tile of the group that has a neighbour at that offset).
// The eight gathered-tile stages form NB = 8 / T block slots; stage (slot, t) = slot * T + t always belongs to tile t, so
// producer warp w = stage w and MMA warp t = tile t see every phase of "their" barriers whatever the masks skip.
template <int COUT, int T_, int S_, bool SWAP_ = false, int PW_ = 1>
struct B2Cfg {
  static constexpr int PW = PW_;                               // producer warps per gathered-tile stage (each fills 128 / PW rows)
  static constexpr int T = T_;
  static constexpr int S = S_;
  static constexpr bool SWAP = SWAP_;                          // operands swapped: D[cout, rows of both tiles] (see the MMA warps)
  static constexpr int MW = SWAP ? 2 : T;                      // MMA warps (swapped: two issuers that alternate blocks)
  static constexpr int CTAS_PER_SM = S == 4 ? 2 : 1;
  static constexpr int NB = S / T;                             // gathered-tile slots per tile (stage = slot * T + t)
  // weight-tile ring.  Swapped operands: FOUR stages -- the issuing warps waited 17 % of their time for weight tiles with
  // three (two blocks of look-ahead do not always cover a 16 KB bulk copy under load); the epilogue stages 16 instead of 32
  // rows per item there, which pays for the extra stage
  static constexpr int SB = S == 4 ? (COUT >= 64 ? 2 : 4) : (COUT >= 128 ? (SWAP ? 4 : 3) : 4);
  static constexpr int EPI_ROWS = SWAP ? 16 : 32;              // rows staged at a time per epilogue warp
  static constexpr int UTIL_WARP = S * PW;
  static constexpr int MMA_WARP0 = S * PW + 1;
  static constexpr int EPI_WARPS = 8;                          // two per TMEM lane quarter: they split the (tile, column chunk) items
  static constexpr int WARPS = MMA_WARP0 + MW + EPI_WARPS;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int EPI_WARP0 = MMA_WARP0 + MW;
  static constexpr int B_STAGE = COUT * 128;                   // COUT rows x [w1 (64 B) | w2 (64 B)]
  static constexpr int ACC_STRIDE = COUT < 32 ? 32 : COUT;
  static constexpr int ACC_BUF = T * ACC_STRIDE;               // one accumulator set (T tiles)
  static constexpr int ACC_COLS = 2 * ACC_BUF;                 // double buffered: the epilogue of a group runs under the next
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static constexpr int EPC = COUT < 32 ? COUT : 32;
  static constexpr int LIST_BYTES = 2 * kB2ListCap * 2;
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_ROWS * kB2EpiRow;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = S * kB2AStage + SB * B_STAGE + LIST_BYTES + EPI_BYTES + S * 1024 + BAR_BYTES + 1024;
  static_assert(T == 2 || T == 4, "tiles per group");
  static_assert(PW == 1 || PW == 2, "producer warps per stage");
  static_assert(!SWAP || (COUT == 128 && T == 2), "swapped operands: M = COUT = 128, N = 2 tiles x 128 rows");
  static_assert((S == 4 || S == 8) && NB >= 1 && (NB & (NB - 1)) == 0, "stage ring");
  static_assert(ACC_COLS * CTAS_PER_SM <= 512, "TMEM budget");
  static_assert(TMEM_COLS * CTAS_PER_SM <= 512, "TMEM budget");
  static_assert((SMEM_BYTES + 1024) * CTAS_PER_SM <= 227 * 1024, "shared memory budget");
  static_assert(COUT % 16 == 0 && COUT >= 16 && COUT <= 128, "UMMA N constraint for M = 128");
  static_assert(2 * S + 2 * SB + 10 <= BAR_BYTES / 8 - 2, "barrier block");
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
  int n_tiles, n_groups;     // 128-row tiles; groups of T consecutive tiles
  // dense-grid (TMA) mode: the input is a regular [B, H, W] map, a tile is kGridTH x kGridTW pixels, the gathered tile of a
  // kernel offset is ONE 4-D tensor-map box (zero fill outside the map = the convolution's padding)
  int grid_tiles_x, grid_tiles_y, grid_kw, grid_pad;
  int g4;                // gather through TMA (tile::gather4, four 128 B rows per instruction) instead of cp.async; kps == 1 only
  unsigned int* sched;   // dynamic group scheduler: [gridDim.y] next-group counters + CTA exit counter at [15]; null = static
  long long* prof;   // optional [gridDim.x][16] cycle counters (tools/microbench_bf2.py --prof), null in production
  int dbg;   // ablation switches (tools/microbench_bf2.py): 1 no gather, 2 no MMA, 4 no weight copies, 8 no stores, 16 all rows
             // missing, 64 no index loads, 128 no epilogue, 256 no evict-last hint on the gathers, 512 plain (not evict-first)
             // table loads, 1024 no L2 prefetch of the next block's rows
};

constexpr int kGridTW = 16, kGridTH = 8;   // pixels of a dense-grid tile (16 x 8 = 128 rows)

// 4-D tiled TMA load: box (32 words, kGridTW, kGridTH, 1) at (c, x, y, b) -> 128 swizzled 128 B rows at dst, completing on bar
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c, int x, int y, int b, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(map), "r"(c), "r"(x), "r"(y), "r"(b), "r"(bar)
      : "memory");
}

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

// Persistent CTA (one per SM): group j of this CTA is tile group blockIdx.x + j * gridDim.x.  Per group the utility warp
// writes the block list (double buffered); every role walks ALL blocks of the list (so every waiter sees every phase of
// the per-slot barriers); the epilogue warps drain a group's accumulators from one TMEM buffer while the next group
// is gathered and multiplied into the other.
//   block entry = kk | live-tile nibble << 5 | chunk << 9
template <int COUT, int T, int S, bool TMA, bool SWAP, int PW>
__global__ void __launch_bounds__(B2Cfg<COUT, T, S, SWAP, PW>::THREADS, B2Cfg<COUT, T, S, SWAP, PW>::CTAS_PER_SM)
conv_bf2_kernel(const __grid_constant__ B2Args A, const __grid_constant__ CUtensorMap in_map) {
  using Cfg = B2Cfg<COUT, T, S, SWAP, PW>;
  constexpr int MW = Cfg::MW;
  constexpr int kB2Stages = S, kB2ProducerWarps = S * PW, kB2UtilWarp = Cfg::UTIL_WARP, kB2MmaWarp0 = Cfg::MMA_WARP0;
  constexpr int B_STAGE = Cfg::B_STAGE, NB = Cfg::NB, SB = Cfg::SB;
  const int NCHUNK = A.nchunk, K = A.K, KS = A.ksteps, kps = A.kps, n_out = A.n_out;
  const int cblk = blockIdx.y * COUT;
  const int n_groups = A.n_groups, gstep = (int)gridDim.x;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;                                        // 8 x [128 rows x 128 B], 128B-swizzled
  uint8_t* b_ring = a_ring + kB2Stages * kB2AStage;              // NB x [COUT rows x 128 B], 128B-swizzled
  uint8_t* lists = b_ring + SB * B_STAGE;                        // 2 x kB2ListCap u16 block entries
  uint8_t* epi = lists + Cfg::LIST_BYTES;                        // 8 warps x [32 rows x 144 B]
  uint8_t* idx_scratch = epi + Cfg::EPI_BYTES;                   // 8 producer warps x 1 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(idx_scratch + kB2Stages * 1024);
  uint64_t* bar_a_full = bars;                                   // [8]  producer warp -> MMA warp of the stage's tile
  uint64_t* bar_a_empty = bar_a_full + kB2Stages;                // [8]  MMA warp of the tile -> producer warp of the stage
  uint64_t* bar_b_full = bar_a_empty + kB2Stages;                // [SB] weight tile landed
  uint64_t* bar_b_empty = bar_b_full + SB;                       // [SB] T MMA warps -> weight loader
  uint64_t* bar_list_full = bar_b_empty + SB;                    // [2]
  uint64_t* bar_list_empty = bar_list_full + 2;                  // [2]
  uint64_t* bar_acc_full = bar_list_empty + 2;                   // [2]
  uint64_t* bar_acc_empty = bar_acc_full + 2;                    // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);
  uint32_t* s_nblocks = s_tmem + 1;                              // [2]
  volatile int* s_group = reinterpret_cast<volatile int*>(s_nblocks + 2);   // [2] tile group of the list buffer, -1 = no more work
  volatile uint32_t* s_turn = reinterpret_cast<volatile uint32_t*>(s_nblocks + 4);   // swapped operands: next block that may issue

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((int)blockIdx.x >= n_groups) return;

  if (warp == kB2MmaWarp0 && lane == 0) {
    for (int s = 0; s < kB2Stages; ++s) {
      mbar_init(smem_u32(bar_a_full + s), TMA ? 1 : PW);          // the producer warp(s) that own the stage
      mbar_init(smem_u32(bar_a_empty + s), 1);                    // tcgen05.commit of the MMA warp that owns the tile
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(smem_u32(bar_b_full + s), 1);                     // arrive.expect_tx of the loader
      mbar_init(smem_u32(bar_b_empty + s), SWAP ? 1 : T);         // every MMA warp: commit (or plain arrival if its tile is dead)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(bar_list_full + s), 1);                  // utility warp, after writing the list
      mbar_init(smem_u32(bar_list_empty + s), (TMA ? 1 : kB2ProducerWarps) + MW + Cfg::EPI_WARPS);   // every reader of the list / group id
      mbar_init(smem_u32(bar_acc_full + s), MW);                  // every MMA warp after its last MMA of the group
      mbar_init(smem_u32(bar_acc_empty + s), Cfg::EPI_WARPS);     // the epilogue warps
    }
    *s_turn = 0u;
    fence_barrier_init();
  }
  if (warp == kB2UtilWarp) tmem_alloc(smem_u32(s_tmem), Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t lists0 = smem_u32(lists);
  const bool prof = A.prof != nullptr;

  if (TMA && warp < kB2ProducerWarps) {
    // ===================== dense-grid producer: ONE thread issues a TMA box per (block, live tile) =====================
    // The regular structure of a BEV map needs no neighbour table: the 128 rows a tile gathers for kernel offset (ky, kx)
    // are the kGridTH x kGridTW box shifted by the offset, and everything outside the map is zero-filled by the TMA unit.
    if (warp == 0 && lane == 0) {
      uint32_t ph = (1u << kB2Stages) - 1u;                  // phase bit per stage; first use of every stage: free
      const int tiles_img = A.grid_tiles_x * A.grid_tiles_y;
      int gblk = 0, j = 0;
#pragma unroll 1
      for (;; ++j) {
        const int buf = j & 1;
        mbar_wait(smem_u32(bar_list_full + buf), (uint32_t)(j >> 1) & 1u);
        const int g = s_group[buf];
        if (g < 0) break;
        const int nblocks = (int)s_nblocks[buf];
        const uint32_t list0 = lists0 + (uint32_t)buf * (kB2ListCap * 2);
        int bx[T], by[T], bb[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int tile = g * T + t;
          bb[t] = tile / tiles_img;
          const int r = tile - bb[t] * tiles_img;
          by[t] = (r / A.grid_tiles_x) * kGridTH - A.grid_pad;
          bx[t] = (r % A.grid_tiles_x) * kGridTW - A.grid_pad;
        }
#pragma unroll 1
        for (int ib = 0; ib < nblocks; ++ib) {
          const uint32_t e = lds_u16(list0 + 2u * (uint32_t)ib);
          const int kk = (int)(e & 31u), chunk = (int)(e >> 9);
          const int ky = kk / A.grid_kw, kx = kk - ky * A.grid_kw;
          const int slot = (gblk + ib) & (NB - 1);
#pragma unroll
          for (int t = 0; t < T; ++t) {
            if (!((e >> (5 + t)) & 1u)) continue;
            const int stage = slot * T + t;
            mbar_wait(smem_u32(bar_a_empty + stage), (ph >> stage) & 1u);
            ph ^= 1u << stage;
            const uint32_t bar = smem_u32(bar_a_full + stage);
            mbar_arrive_expect_tx(bar, kB2AStage);
            tma_load_4d(smem_u32(a_ring) + (uint32_t)stage * kB2AStage, &in_map, chunk * 32, bx[t] + kx, by[t] + ky, bb[t], bar);
          }
        }
        gblk += nblocks;
        mbar_arrive(smem_u32(bar_list_empty + buf));
      }
    }
  } else if (warp < kB2ProducerWarps) {
    // ===================== producers: warp w owns stage w = (slot w / T, tile w % T) =====================
    // A warp's iteration is a serial chain of ~100 dependent instructions plus the latency of its own copies, so whole
    // 128-row tiles are dealt over the warps: eight gathers are in flight and one mbarrier arrival publishes a stage.
    //   lane = (o = lane >> 3, c8 = lane & 7): copy r (0..31) moves piece c8 of row 4r + o.
    // With PW = 2 two warps share a stage, each gathering 64 of its rows: the fill of a stage (issue of the copies + their
    // latency) is the long leg of the stage's cycle fill -> MMA -> empty, and that cycle, not a bandwidth, set the pace.
    constexpr int R = 32 / PW;                               // copies per lane: rows hbase + 4r + o, r = 0..R-1
    const int my_stage = warp / PW, hbase = (warp % PW) * (kBM / PW);
    const int my_slot = my_stage / T, my_t = my_stage % T;
    const int o = lane >> 3, c8 = lane & 7;
    const int sub = kps == 2 ? (c8 >> 2) : 0;                // Cin = 16: pieces 0-3 come from offset 2kk, 4-7 from 2kk + 1
    const uint32_t piece = (uint32_t)(kps == 2 ? (c8 & 3) : c8) * 16u;
    const char* in_bytes = reinterpret_cast<const char*>(A.in);
    const uint32_t row_units = (uint32_t)A.in_ld >> 2;       // row stride in 16 B units (32-bit offsets reach 64 GiB)
    const int* tbl = A.tbl;
    const int tbl_stride = A.tbl_stride;
    const uint32_t a_stage = smem_u32(a_ring) + (uint32_t)my_stage * kB2AStage + (uint32_t)hbase * 128u;
    const uint32_t bar_full = smem_u32(bar_a_full + my_stage), bar_empty = smem_u32(bar_a_empty + my_stage);
    uint32_t phase = 1;                                      // first use of the stage: free
    // the gathered map is re-read once per kernel offset: ask L2 to keep it (the read-once table and residual and the
    // written-once fp32 output are streamed with evict-first)
    uint64_t keep_policy;
    if (A.dbg & 256) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(keep_policy));
    else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));
    // index scratch of this warp: [2 offsets][4 row residues o][R copies r] -> idx of row hbase + 4r + o, read back as int4 over r
    const uint32_t scr = smem_u32(idx_scratch) + (uint32_t)warp * (1024u / PW);
    const uint32_t scr_rd = scr + (uint32_t)sub * (16u * R) + (uint32_t)o * (4u * R);
    const uint32_t dst_lane = (uint32_t)o * 128u;            // row 4r + o: byte (4r + o) * 128, swizzle ((4r + o) & 7) ^ c8
    int gblk = 0;                                            // blocks of this CTA before the current group
    int j = 0;
    long long pw_empty = 0, pw_data = 0, pw_list = 0, p_steps = 0;
    const long long p_t0 = prof ? clock64() : 0;
#pragma unroll 1
    for (;; ++j) {
      const int buf = j & 1;
      long long c0 = prof ? clock64() : 0;
      mbar_wait(smem_u32(bar_list_full + buf), (uint32_t)(j >> 1) & 1u);
      if (prof) pw_list += clock64() - c0;
      const int g = s_group[buf];
      if (g < 0) break;
      const int nblocks = (int)s_nblocks[buf];
      const uint32_t list0 = lists0 + (uint32_t)buf * (kB2ListCap * 2);
      const int tile0 = (g * T + my_t) * kBM;                // first row of this warp's tile
      const int row_end = min(n_out, tile0 + kBM);
      // this lane's four consecutive rows 4 lane .. 4 lane + 3 of the tile, one int4 per offset of the block
      auto load_idx = [&](uint32_t e, int s2) -> int4 {
        int4 v = make_int4(-1, -1, -1, -1);
        const int k = (int)(e & 31u) * kps + s2;
        const int row0 = tile0 + hbase + 4 * lane;
        if (lane < R && ((e >> (5 + my_t)) & 1u) && k < K && row0 < row_end && !(A.dbg & 64)) {
          if (A.dbg & 512) v = __ldg(reinterpret_cast<const int4*>(tbl + (size_t)k * tbl_stride + row0));
          else v = __ldcs(reinterpret_cast<const int4*>(tbl + (size_t)k * tbl_stride + row0));   // read once: evict first
          const int lim = row_end - row0;
          if (lim < 4) { if (lim < 2) v.y = -1; if (lim < 3) v.z = -1; v.w = -1; }
        }
        return v;
      };
      int ib = (my_slot - gblk) & (NB - 1);                  // first block of this group in this warp's slot
      uint32_t e = 0;
      int4 qa = make_int4(-1, -1, -1, -1), qb = qa;
      if (ib < nblocks) {
        e = lds_u16(list0 + 2u * (uint32_t)ib);
        qa = load_idx(e, 0);
        if (kps == 2) qb = load_idx(e, 1);
      }
#pragma unroll 1
      for (; ib < nblocks; ib += NB) {
        const bool live = (e >> (5 + my_t)) & 1u;
        if (live && !A.g4) {
          // indices -> scratch, transposed so that the copies of row residue o read four consecutive r with one LDS.128
          if (lane < R) {
            sts32(scr + 4u * lane, qa.x); sts32(scr + 4u * R + 4u * lane, qa.y);
            sts32(scr + 8u * R + 4u * lane, qa.z); sts32(scr + 12u * R + 4u * lane, qa.w);
            if (kps == 2) {
              sts32(scr + 16u * R + 4u * lane, qb.x); sts32(scr + 20u * R + 4u * lane, qb.y);
              sts32(scr + 24u * R + 4u * lane, qb.z); sts32(scr + 28u * R + 4u * lane, qb.w);
            }
          }
        }
        const uint32_t cunits = (kps == 2 ? 0u : (e >> 9) * 8u) + (piece >> 4);
        // prefetch the entry and the indices of this warp's next block (NB blocks ahead)
        uint32_t e_n = 0;
        int4 qa_n = make_int4(-1, -1, -1, -1), qb_n = qa_n;
        if (ib + NB < nblocks) {
          e_n = lds_u16(list0 + 2u * (uint32_t)(ib + NB));
          qa_n = load_idx(e_n, 0);
          if (kps == 2) qb_n = load_idx(e_n, 1);
        }
        __syncwarp();
        if (PW == 1 && live && A.g4) {
          // TMA gather: lane l moves rows 4l .. 4l + 3 of the tile (its qa) with ONE gather4; a negative row is zero filled;
          // the stage's barrier counts the bytes, so the warp neither waits for the data nor fences it
          c0 = prof ? clock64() : 0;
          mbar_wait(bar_empty, phase);
          if (prof) { pw_empty += clock64() - c0; ++p_steps; }
          phase ^= 1;
          if (lane == 0) mbar_arrive_expect_tx(bar_full, (A.dbg & 1) ? 0u : (uint32_t)kB2AStage);
          __syncwarp();
          if (!(A.dbg & 1))
            tma_gather4(a_stage + (uint32_t)lane * 512u, &in_map, (int)(e >> 9) * 32, qa.x, qa.y, qa.z, qa.w, bar_full);
        } else if (live) {
          c0 = prof ? clock64() : 0;
          mbar_wait(bar_empty, phase);                         // the MMAs that read the stage's previous tile are done
          if (prof) { pw_empty += clock64() - c0; ++p_steps; }
          phase ^= 1;
#pragma unroll
          for (int q = 0; q < R / 4; ++q) {
            if (A.dbg & 1) break;
            const int4 ii = lds128i(scr_rd + 16u * q);
            const int idx[4] = {ii.x, ii.y, ii.z, ii.w};
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const int r = 4 * q + rr;                        // row 4r + o; (4r + o) & 7 = 4 (r & 1) + o
              const uint32_t dst = a_stage + (uint32_t)r * 512u + dst_lane + (uint32_t)((c8 ^ (4 * (r & 1) + o)) << 4);
              const uint32_t off = (uint32_t)max(idx[rr], 0) * row_units + cunits;
              cp_async16_zfill_hint(dst, in_bytes + ((size_t)off << 4), (idx[rr] >= 0 && !(A.dbg & 16)) ? 16u : 0u, keep_policy);
            }
          }
          cp_async_commit();
          // While these copies fly: pull the rows of this warp's NEXT block into L2 (its indices were requested before the
          // copies were issued).  A stage is published when its slowest row has landed, and with 20-25 % of the sectors
          // coming from DRAM nearly every stage has such a row; the next block's copies then find all their lines in L2.
          if (kps == 1 && !(A.dbg & 1024) && ((e_n >> (5 + my_t)) & 1u)) {
            const uint32_t cu = (e_n >> 9) * 8u;
            const int pi[4] = {qa_n.x, qa_n.y, qa_n.z, qa_n.w};
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
              if (pi[rr] >= 0)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(in_bytes + ((size_t)((uint32_t)pi[rr] * row_units + cu) << 4)));
          }
          c0 = prof ? clock64() : 0;
          cp_async_wait<0>();                                  // this lane's pieces have landed ...
          if (prof) pw_data += clock64() - c0;
          fence_proxy_async();                                 // ... and are visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_full);
        }
        e = e_n; qa = qa_n; qb = qb_n;
      }
      gblk += nblocks;
      __syncwarp();                                          // every lane has read its last list entry
      if (lane == 0) mbar_arrive(smem_u32(bar_list_empty + buf));
    }
    if (prof && warp == 0 && lane == 0) {
      long long* po = A.prof + (size_t)blockIdx.x * 16;
      po[0] = clock64() - p_t0; po[1] = pw_empty; po[2] = pw_data; po[3] = pw_list; po[4] = p_steps;
    }
  } else if (warp == kB2UtilWarp) {
    // ===================== utility warp: block lists, weight tiles, index prefetch =====================
    // Block list of a group: lane = offset step kk; entries ordered (kk, chunk); offsets at which no tile of the group has
    // a neighbour are skipped.  entry = kk | live-tile nibble << 5 | chunk << 9
    // Group of list jj: static (CTA b takes b, b + grid, ...) or, with a scheduler, the next group nobody has taken -- the
    // grouped rulebooks order their rows by DEscending number of live offset triples (grouping.cuh), so the groups come
    // heaviest first and the CTAs finish within one light group of each other (a 27-offset group of 128 channels is up to
    // 108 blocks = 55 us, a sixth of the whole launch: with the static round robin the slowest CTA set the time).
    auto build_list = [&](int jj) {
      const int buf = jj & 1;
      mbar_wait(smem_u32(bar_list_empty + buf), ((uint32_t)(jj >> 1) & 1u) ^ 1u);
      int g = (int)blockIdx.x + jj * gstep;
      if (A.sched) {
        if (lane == 0) g = (int)atomicAdd(A.sched + blockIdx.y, 1u);
        g = __shfl_sync(0xffffffffu, g, 0);
      }
      if (g >= n_groups) {
        if (lane == 0) { s_group[buf] = -1; s_nblocks[buf] = 0; }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(bar_list_full + buf));
        return;
      }
      if (lane == 0) s_group[buf] = g;
      const int tile_first = g * T;
      const int Tr = min(T, A.n_tiles - tile_first);
      uint16_t* blocks = reinterpret_cast<uint16_t*>(lists) + buf * kB2ListCap;
      uint32_t tmask[T];
#pragma unroll
      for (int t = 0; t < T; ++t)                                   // all loads in flight together
        tmask[t] = (A.tile_masks && t < Tr) ? (uint32_t)__ldg(A.tile_masks + tile_first + t) : 0xffffffffu;
      uint32_t nib = 0;
      if (lane < KS) {
#pragma unroll
        for (int t = 0; t < T; ++t) {
          if (t >= Tr) break;
          uint32_t m = tmask[t];
          m &= K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
          if (m == 0) m = 1;                                        // a tile always runs at least one block (zero rows)
          const uint32_t live = kps == 1 ? (m >> lane) & 1u : ((m >> (2 * lane)) & 3u) != 0;
          nib |= live << t;
        }
      }
      // chunk-major order: the live offsets of one 32-channel chunk run back to back, so the 128 B pieces a tile gathers
      // for neighbouring offsets (the same rows, shifted) are re-read while still in L2 -- with offset-major order a
      // 188 x 188 x 512-channel map (290 MB) streamed from DRAM once per offset (ncu: 2.5 GB read, 8 % L2 hits)
      const unsigned live_mask = __ballot_sync(0xffffffffu, nib != 0);
      const int n_live = __popc(live_mask), rank = __popc(live_mask & ((1u << lane) - 1u));
      if (nib)
        for (int c = 0; c < NCHUNK; ++c)
          if (c * n_live + rank < kB2ListCap) blocks[c * n_live + rank] = (uint16_t)(lane | (nib << 5) | (c << 9));
      const int total = n_live * NCHUNK;
      if (lane == 0) s_nblocks[buf] = (uint32_t)min(total, kB2ListCap);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bar_list_full + buf));
    };
    // All lanes pull the neighbour-table lines of the blocks kIdxAhead ahead into L2 (the table is 27 x 4 B per row, larger
    // than the feature maps, and streams from DRAM; the producers' index loads then hit L2).  lane = (tile, offset of the
    // block, 128 B line of the tile's 512 B index slice).  Lane 0 also feeds the weight ring.
    constexpr int kIdxAhead = 8;
    const uint8_t* packed = reinterpret_cast<const uint8_t*>(A.packed) + (size_t)blockIdx.y * KS * NCHUNK * B_STAGE;
    const int* tbl = A.tbl;
    const int tbl_stride = A.tbl_stride;
    int gblk = 0;
    int j = 0;
    build_list(0);
#pragma unroll 1
    for (;; ++j) {
      const int buf = j & 1;
      const int g = s_group[buf];
      if (g < 0) break;
      build_list(j + 1);
      const int nblocks = (int)s_nblocks[buf];
      const uint32_t list0 = lists0 + (uint32_t)buf * (kB2ListCap * 2);
      const int row_end = min(n_out, (g + 1) * T * kBM);
      auto prefetch_idx = [&](int ib) {
        const int pt = lane >> 3, ps = (lane >> 2) & 1, pl = lane & 3;
        if (TMA || ib >= nblocks || pt >= T || ps >= kps || (A.dbg & 64)) return;
        const uint32_t e = lds_u16(list0 + 2u * (uint32_t)ib);
        if ((e >> 9) != 0u || !((e >> (5 + pt)) & 1u)) return;      // all chunks of an offset use the same indices
        const int k = (int)(e & 31u) * kps + ps;
        const int row = (g * T + pt) * kBM + 32 * pl;                        // 128 rows x 4 B = four 128 B lines
        if (k < K && row < row_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(tbl + (size_t)k * tbl_stride + row));
      };
      for (int ib = 0; ib < kIdxAhead; ++ib) prefetch_idx(ib);
#pragma unroll 1
      for (int ib = 0; ib < nblocks; ++ib) {
        prefetch_idx(ib + kIdxAhead);
        if (lane == 0) {
          const uint32_t e = lds_u16(list0 + 2u * (uint32_t)ib);
          const int gb = gblk + ib, slot = gb % SB;
          const int bstep = (int)(e & 31u) * NCHUNK + (int)(e >> 9);
          mbar_wait(smem_u32(bar_b_empty + slot), ((uint32_t)(gb / SB) & 1u) ^ 1u);
          const uint32_t bar = smem_u32(bar_b_full + slot);
          if (A.dbg & 4) {
            mbar_arrive(bar);
          } else {
            mbar_arrive_expect_tx(bar, B_STAGE);
            bulk_copy_g2s(smem_u32(b_ring + (size_t)slot * B_STAGE), packed + (size_t)bstep * B_STAGE, B_STAGE, bar);
          }
        }
        __syncwarp();
      }
      gblk += nblocks;
    }
  } else if (warp < Cfg::EPI_WARP0) {
    // ===================== MMA warps: warp t issues the MMAs of tile t =====================
    // One issuing thread needs ~500-700 clk per step for six tcgen05.mma and a commit (measured), more than the gather;
    // with a warp per tile the T accumulators advance in parallel.  Warp-uniform control flow and values (the entry is
    // broadcast with a shuffle) keep descriptors and barrier addresses in uniform registers; one elected lane issues.
    constexpr uint32_t desc_hi = 64u | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
    if constexpr (SWAP) {
      // ---- swapped operands (COUT = 128, T = 2): D[cout, row] = W^T x X^T, two warps issue ALTERNATE blocks ----
      // Measured (tools/mma_probe.cu, tools/sync_probe.cu): SS-mode MMAs of M = 128, K = 16 run at their floor of N / 2 clk
      // (64 clk at N = 128 even under shared-memory write traffic); the tensor pipe queues only ~1.5 MMAs behind the one
      // executing, so every clock beyond ~170 that the issuing warp spends between two blocks (barrier waits, list entry,
      // commits, and the ~140 clk before the uniform registers of an issued UTCHMMA may be rewritten) is a clock the pipe
      // idles: with one issuer per tile the Cout = 128 kernels sat at 60-70 % tensor activity.  Here the WEIGHT tile is the
      // M = 128 operand and the gathered tiles of both tiles of the group ONE N = 256 operand (stages (slot, 0) and (slot, 1)
      // are adjacent: 256 rows x 128 B), a block is six 128 clk MMAs, and two warps take turns: while one issues block i, the
      // other waits for the operands of block i + 1 and builds its descriptors, then only needs the `turn` hand-off (a block
      // counter in shared memory that the waiting warp polls).  The turn also fixes the order in which the MMAs reach the
      // pipe, so sums are reproducible.  The accumulator is
      // transposed (TMEM lane = output channel, column = row of the group); the epilogue transposes it back while staging.
      const int p = warp - kB2MmaWarp0;                           // this warp issues the blocks with (gblk + ib) & 1 == p
      constexpr uint32_t idesc2 = make_idesc_bf16(COUT, 2 * kBM), idesc1 = make_idesc_bf16(COUT, kBM);
      const uint32_t a_lo0 = ((smem_u32(a_ring) >> 4) & 0x3FFF) | (1u << 16);
      const uint32_t b_lo0 = ((smem_u32(b_ring) >> 4) & 0x3FFF) | (1u << 16);
      const uint32_t bar_a_full0 = smem_u32(bar_a_full), bar_a_empty0 = smem_u32(bar_a_empty);
      const uint32_t bar_b_full0 = smem_u32(bar_b_full), bar_b_empty0 = smem_u32(bar_b_empty);
      const bool leader = elect_one();
      const uint64_t hi64 = (uint64_t)desc_hi << 32;
      // six MMAs: (weight slice, activation slice) pairs, small terms first (same pairs as the unswapped issue below)
      auto issue6 = [&](uint32_t d, uint32_t w_lo, uint32_t x_lo, uint32_t idesc, uint32_t acc) {
        if (kps == 1) {
          umma_bf16_ss(d, hi64 | (w_lo + 0u), hi64 | (x_lo + 4u), idesc, acc);
          umma_bf16_ss(d, hi64 | (w_lo + 2u), hi64 | (x_lo + 6u), idesc, 1u);
          umma_bf16_ss(d, hi64 | (w_lo + 4u), hi64 | (x_lo + 0u), idesc, 1u);
          umma_bf16_ss(d, hi64 | (w_lo + 6u), hi64 | (x_lo + 2u), idesc, 1u);
          umma_bf16_ss(d, hi64 | (w_lo + 0u), hi64 | (x_lo + 0u), idesc, 1u);
          umma_bf16_ss(d, hi64 | (w_lo + 2u), hi64 | (x_lo + 2u), idesc, 1u);
        } else {
          umma_bf16_ss(d, hi64 | (w_lo + 0u), hi64 | (x_lo + 2u), idesc, acc);
          umma_bf16_ss(d, hi64 | (w_lo + 2u), hi64 | (x_lo + 0u), idesc, 1u);
          umma_bf16_ss(d, hi64 | (w_lo + 4u), hi64 | (x_lo + 6u), idesc, 1u);
          umma_bf16_ss(d, hi64 | (w_lo + 6u), hi64 | (x_lo + 4u), idesc, 1u);
          umma_bf16_ss(d, hi64 | (w_lo + 0u), hi64 | (x_lo + 0u), idesc, 1u);
          umma_bf16_ss(d, hi64 | (w_lo + 4u), hi64 | (x_lo + 4u), idesc, 1u);
        }
      };
      uint32_t a_ph = 0;                                          // phase bit per stage (this warp's slots only)
      int gblk = 0, j = 0;
      long long mw_a = 0, mw_b = 0, mw_acc = 0, mw_list = 0, m_steps = 0;
      const long long m_t0 = prof ? clock64() : 0;
      for (;; ++j) {
        const int buf = j & 1;
        long long c0 = prof ? clock64() : 0;
        mbar_wait(smem_u32(bar_list_full + buf), (uint32_t)(j >> 1) & 1u);
        if (prof) { const long long c1 = clock64(); mw_list += c1 - c0; c0 = c1; }
        if (s_group[buf] < 0) break;
        mbar_wait(smem_u32(bar_acc_empty + buf), ((uint32_t)(j >> 1) & 1u) ^ 1u);
        if (prof) mw_acc += clock64() - c0;
        tc_fence_after();
        const int nblocks = __shfl_sync(0xffffffffu, (int)s_nblocks[buf], 0);
        const uint32_t list0 = lists0 + (uint32_t)buf * (kB2ListCap * 2);
        const uint32_t d0 = tmem_base + (uint32_t)(buf * Cfg::ACC_BUF);
        uint32_t accm = 0;                                        // bit t: tile t's columns hold a partial sum (both warps track it)
        bool issued = false;
        uint32_t e_next = lds_u16(list0);
#pragma unroll 1
        for (int ib = 0; ib < nblocks; ++ib) {
          const uint32_t e = __shfl_sync(0xffffffffu, e_next, 0);
          e_next = lds_u16(list0 + 2u * (uint32_t)min(ib + 1, nblocks - 1));
          const int gb = gblk + ib;
          const uint32_t nib = (e >> 5) & 3u;
          if ((gb & 1) == p) {
            const int slot = gb & (NB - 1), bslot = gb % SB;
            c0 = prof ? clock64() : 0;
            mbar_wait(bar_b_full0 + 8 * bslot, (uint32_t)(gb / SB) & 1u);
            if (prof) { const long long c1 = clock64(); mw_b += c1 - c0; c0 = c1; }
            const int stage0 = slot * 2;
            if (nib & 1u) { mbar_wait(bar_a_full0 + 8 * stage0, (a_ph >> stage0) & 1u); a_ph ^= 1u << stage0; }
            if (nib & 2u) { mbar_wait(bar_a_full0 + 8 * (stage0 + 1), (a_ph >> (stage0 + 1)) & 1u); a_ph ^= 2u << stage0; }
            if (prof) { mw_a += clock64() - c0; ++m_steps; }
            const uint32_t x_lo = a_lo0 + (uint32_t)stage0 * (kB2AStage >> 4);
            const uint32_t w_lo = b_lo0 + (uint32_t)bslot * (B_STAGE >> 4);
            // the turn: block gb may issue once every block before it has (a counter in shared memory, polled; the
            // hand-off is a plain store after the last UTCHMMA of a block -- shorter than an mbarrier round trip)
            for (uint32_t spin = 0; *s_turn != (uint32_t)gb; ++spin)
              if (spin > (1u << 26)) __trap();                   // bounded: a protocol bug must not hang the GPU
            tc_fence_after();
            if (leader) {
              if (!(A.dbg & 2)) {
                if (nib == 3u && (accm == 0u || accm == 3u)) {
                  issue6(d0, w_lo, x_lo, idesc2, accm ? 1u : 0u);
                } else {
                  if (nib & 1u) issue6(d0, w_lo, x_lo, idesc1, accm & 1u);
                  if (nib & 2u) issue6(d0 + (uint32_t)kBM, w_lo, x_lo + (uint32_t)(kB2AStage >> 4), idesc1, (accm >> 1) & 1u);
                }
              }
              *s_turn = (uint32_t)gb + 1u;
              if (nib & 1u) umma_commit(bar_a_empty0 + 8 * stage0);
              if (nib & 2u) umma_commit(bar_a_empty0 + 8 * (stage0 + 1));
              if (nib) umma_commit(bar_b_empty0 + 8 * bslot);
              else mbar_arrive(bar_b_empty0 + 8 * bslot);
            }
            issued = issued || nib != 0u;
            __syncwarp();
          }
          accm |= nib;
        }
        gblk += nblocks;
        if (leader) {
          if (issued) umma_commit(smem_u32(bar_acc_full + buf));    // this warp's MMAs of the group are complete
          else mbar_arrive(smem_u32(bar_acc_full + buf));
          mbar_arrive(smem_u32(bar_list_empty + buf));
        }
        __syncwarp();
      }
      if (prof && p == 0 && lane == 0) {
        long long* po = A.prof + (size_t)blockIdx.x * 16;
        po[5] = clock64() - m_t0; po[6] = mw_a; po[7] = mw_b; po[8] = mw_acc; po[9] = mw_list; po[10] = m_steps;
      }
    } else {
    const int t = warp - kB2MmaWarp0;
    constexpr uint32_t idesc = make_idesc_bf16(kBM, COUT);
    const uint32_t a_lo0 = ((smem_u32(a_ring) >> 4) & 0x3FFF) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(b_ring) >> 4) & 0x3FFF) | (1u << 16);
    const uint32_t bar_a_full0 = smem_u32(bar_a_full), bar_a_empty0 = smem_u32(bar_a_empty);
    const uint32_t bar_b_full0 = smem_u32(bar_b_full), bar_b_empty0 = smem_u32(bar_b_empty);
    // (A slice, B slice) of the six MMAs of a tile and block, small terms first; a slice = 32 B = 16 BF16 k positions.
    //   Cin >= 32: A = [hi 0-15 | hi 16-31 | lo 0-15 | lo 16-31], B = [w1 0-15 | w1 16-31 | w2 0-15 | w2 16-31]
    //              pairs (2,0) (3,1) (0,2) (1,3) (0,0) (1,1)
    //   Cin == 16: A = [hi k0 | lo k0 | hi k1 | lo k1],            B = [w1 k0 | w2 k0 | w1 k1 | w2 k1]
    //              pairs (1,0) (0,1) (3,2) (2,3) (0,0) (2,2)
    const bool leader = elect_one();
    uint32_t a_ph = 0;                                            // phase bit per slot of this tile's a_full barriers
    int gblk = 0;
    int j = 0;
    long long mw_a = 0, mw_b = 0, mw_acc = 0, mw_list = 0, m_steps = 0;
    const long long m_t0 = prof ? clock64() : 0;
    for (;; ++j) {
      const int buf = j & 1;
      long long c0 = prof ? clock64() : 0;
      mbar_wait(smem_u32(bar_list_full + buf), (uint32_t)(j >> 1) & 1u);
      if (prof) { const long long c1 = clock64(); mw_list += c1 - c0; c0 = c1; }
      if (s_group[buf] < 0) break;
      mbar_wait(smem_u32(bar_acc_empty + buf), ((uint32_t)(j >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator set
      if (prof) mw_acc += clock64() - c0;
      tc_fence_after();
      const int nblocks = __shfl_sync(0xffffffffu, (int)s_nblocks[buf], 0);
      const uint32_t list0 = lists0 + (uint32_t)buf * (kB2ListCap * 2);
      const uint32_t d = tmem_base + (uint32_t)(buf * Cfg::ACC_BUF + t * Cfg::ACC_STRIDE);
      uint32_t acc = 0;
      uint32_t e_next = lds_u16(list0);
#pragma unroll 1
      for (int ib = 0; ib < nblocks; ++ib) {
        const uint32_t e = __shfl_sync(0xffffffffu, e_next, 0);
        e_next = lds_u16(list0 + 2u * (uint32_t)min(ib + 1, nblocks - 1));
        const int gb = gblk + ib, slot = gb & (NB - 1), bslot = gb % SB;
        const bool live = (e >> (5 + t)) & 1u;
        c0 = prof ? clock64() : 0;
        mbar_wait(bar_b_full0 + 8 * bslot, (uint32_t)(gb / SB) & 1u);     // every block: this warp sees every phase
        if (prof) mw_b += clock64() - c0;
        if (live) {
          const int stage = slot * T + t;
          c0 = prof ? clock64() : 0;
          mbar_wait(bar_a_full0 + 8 * stage, (a_ph >> slot) & 1u);
          if (prof) { mw_a += clock64() - c0; ++m_steps; }
          a_ph ^= 1u << slot;
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (kB2AStage >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)bslot * (B_STAGE >> 4);
          if (leader && !(A.dbg & 2)) {
            const uint64_t hi64 = (uint64_t)desc_hi << 32;
            if (kps == 1) {
              umma_bf16_ss(d, hi64 | (a_lo + 4u), hi64 | (b_lo + 0u), idesc, acc);
              umma_bf16_ss(d, hi64 | (a_lo + 6u), hi64 | (b_lo + 2u), idesc, 1u);
              umma_bf16_ss(d, hi64 | (a_lo + 0u), hi64 | (b_lo + 4u), idesc, 1u);
              umma_bf16_ss(d, hi64 | (a_lo + 2u), hi64 | (b_lo + 6u), idesc, 1u);
              umma_bf16_ss(d, hi64 | (a_lo + 0u), hi64 | (b_lo + 0u), idesc, 1u);
              umma_bf16_ss(d, hi64 | (a_lo + 2u), hi64 | (b_lo + 2u), idesc, 1u);
            } else {
              umma_bf16_ss(d, hi64 | (a_lo + 2u), hi64 | (b_lo + 0u), idesc, acc);
              umma_bf16_ss(d, hi64 | (a_lo + 0u), hi64 | (b_lo + 2u), idesc, 1u);
              umma_bf16_ss(d, hi64 | (a_lo + 6u), hi64 | (b_lo + 4u), idesc, 1u);
              umma_bf16_ss(d, hi64 | (a_lo + 4u), hi64 | (b_lo + 6u), idesc, 1u);
              umma_bf16_ss(d, hi64 | (a_lo + 0u), hi64 | (b_lo + 0u), idesc, 1u);
              umma_bf16_ss(d, hi64 | (a_lo + 4u), hi64 | (b_lo + 4u), idesc, 1u);
            }
          }
          acc = 1u;
          if (leader) {
            umma_commit(bar_a_empty0 + 8 * stage);                  // the gathered tile is reusable once read ...
            umma_commit(bar_b_empty0 + 8 * bslot);                  // ... and so is this warp's share of the weight tile
          }
        } else {
          if (leader) mbar_arrive(bar_b_empty0 + 8 * bslot);        // dead tile: nothing to read
        }
        __syncwarp();
      }
      gblk += nblocks;
      if (leader) {
        if (acc) umma_commit(smem_u32(bar_acc_full + buf));         // this tile's accumulator is complete
        else mbar_arrive(smem_u32(bar_acc_full + buf));             // tile past the end of the tensor: nothing issued
        mbar_arrive(smem_u32(bar_list_empty + buf));
      }
      __syncwarp();
    }
    if (prof && t == 0 && lane == 0) {
      long long* po = A.prof + (size_t)blockIdx.x * 16;
      po[5] = clock64() - m_t0; po[6] = mw_a; po[7] = mw_b; po[8] = mw_acc; po[9] = mw_list; po[10] = m_steps;
    }
    }
  } else {
    // ===================== epilogue warps: TMEM -> staged rows -> coalesced BN / residual / activation / stores ==========
    // A warp reads the accumulator rows of its TMEM lane quarter (lane = row), stages 32 (16) channels per row in shared
    // memory and re-reads them with 8 (4) lanes per row, so that every global access is a whole 128 B (64 B) row piece.
    // The BN affine + activation (GELU: erff) + split of a 128 x 128 tile is ~150 k warp instructions; with four warps the
    // 1x1 layers of the neck were epilogue-bound (256->256 at 188 x 188: 585 us against 180 us for the same shape with ReLU).
    const int g4 = warp & 3;
    const int half = (warp - Cfg::EPI_WARP0) >> 2;                  // which of the quarter's two warps
    const uint32_t stg = smem_u32(epi) + (uint32_t)(warp - Cfg::EPI_WARP0) * (Cfg::EPI_ROWS * kB2EpiRow);
    const float* __restrict__ scale = A.scale ? A.scale + cblk : nullptr;
    const float* __restrict__ shift = A.shift ? A.shift + cblk : nullptr;
    const int act = A.act, res_after = A.res_after_act;
    constexpr int EPC = Cfg::EPC;
    constexpr int PPR = EPC / 4;                                    // 16 B pieces per staged row
    constexpr int RPI = 32 / PPR;                                   // rows covered by one warp-wide access
    const int pr = lane / PPR, pp = lane % PPR;
    int j = 0;
#pragma unroll 1
    for (;; ++j) {
      const int buf = j & 1;
      mbar_wait(smem_u32(bar_list_full + buf), (uint32_t)(j >> 1) & 1u);
      const int g = s_group[buf];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bar_list_empty + buf));   // the group id has been read (the list itself is not used here)
      if (g < 0) break;
      const int tile_first = g * T;
      const int Tr = min(T, A.n_tiles - tile_first);
      const int tile0 = tile_first * kBM;
      const int row_end = min(n_out, tile0 + T * kBM);
      mbar_wait(smem_u32(bar_acc_full + buf), (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      // staged 32 rows x EPC channels of this warp -> affine / residual / activation -> whole-row stores.
      //   rbase: first of the 32 rows inside the group, c0: first channel of the staged piece, orow_l: lane r's output row
      auto store_staged = [&](int rbase, int c0, int orow_l, int lane0, int iters) {
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale + c0) + pp);
        if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift + c0) + pp);
#pragma unroll
        for (int jj = 0; jj < PPR; ++jj) {
          if (jj >= iters) break;
          const int r = pr + RPI * jj;
          const int orow = __shfl_sync(0xffffffffu, orow_l, lane0 + r);
          const int row = tile0 + rbase + r;
          float4 v = lds128(stg + (uint32_t)r * kB2EpiRow + 16u * pp);
          if (row < row_end && orow >= 0 && !(A.dbg & 8)) {
            v.x = v.x * sc.x + sh.x; v.y = v.y * sc.y + sh.y; v.z = v.z * sc.z + sh.z; v.w = v.w * sc.w + sh.w;
            float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
            // read-once / written-once streams must not push the gathered map out of L2: evict-first loads and stores
            if (A.residual) rr = __ldcs(reinterpret_cast<const float4*>(A.residual + (size_t)orow * A.res_ld + cblk + c0) + pp);
            if (!res_after) { v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w; }
            v.x = b2_act(v.x, act); v.y = b2_act(v.y, act); v.z = b2_act(v.z, act); v.w = b2_act(v.w, act);
            if (res_after) { v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w; }
            if (A.out) __stcs(reinterpret_cast<float4*>(A.out + (size_t)orow * A.out_ld + cblk + c0) + pp, v);
            if (A.out_split) {
              // split row: per chunk of EPC channels [EPC/2 words hi | EPC/2 words lo]; this lane owns channels 4pp..4pp+3
              uint32_t h0, l0, h1, l1;
              split_pair(v.x, v.y, h0, l0);
              split_pair(v.z, v.w, h1, l1);
              uint32_t* dst = A.out_split + (size_t)orow * A.split_ld + cblk + c0;
              *reinterpret_cast<uint2*>(dst + 2 * pp) = make_uint2(h0, h1);
              *reinterpret_cast<uint2*>(dst + EPC / 2 + 2 * pp) = make_uint2(l0, l1);
            }
          }
        }
      };
      int item = 0;
      if constexpr (SWAP) {
        // transposed accumulator: this warp's TMEM lane quarter = output channels 32 g4 .. 32 g4 + 31, a column = a row of the
        // group.  An item = 32 rows: lane c reads its channel of 32 consecutive rows and writes them down a staging column.
        const int c0 = g4 * 32;
#pragma unroll 1
        for (int rb = 0; rb < Tr * kBM; rb += 32) {
          if (A.dbg & 128) break;
          if ((item++ & 1) != half) continue;
          if (tile0 + rb >= row_end) break;
          const int row_l = tile0 + rb + lane;
          const int orow_l = (row_l < row_end && A.out_rows) ? __ldg(A.out_rows + row_l) : row_l;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {                              // 16 rows at a time (Cfg::EPI_ROWS)
            uint32_t acc[16];
            tmem_ld<16>(tmem_base + ((uint32_t)(g4 * 32) << 16) + (uint32_t)(buf * Cfg::ACC_BUF + rb + 16 * h), acc);
#pragma unroll
            for (int q = 0; q < 16; ++q) sts32(stg + (uint32_t)q * kB2EpiRow + 4u * lane, (int)acc[q]);
            __syncwarp();
            store_staged(rb + 16 * h, c0, orow_l, 16 * h, PPR / 2);
            __syncwarp();
          }
        }
      } else {
#pragma unroll 1
      for (int t = 0; t < Tr; ++t) {
        if (A.dbg & 128) break;
        const int row_l = tile0 + t * kBM + g4 * 32 + lane;
        const int orow_l = (row_l < row_end && A.out_rows) ? __ldg(A.out_rows + row_l) : row_l;
#pragma unroll 1
        for (int c0 = 0; c0 < COUT; c0 += EPC) {
          if ((item++ & 1) != half) continue;
          uint32_t acc[EPC];
          tmem_ld<EPC>(tmem_base + ((uint32_t)(g4 * 32) << 16) + (uint32_t)(buf * Cfg::ACC_BUF + t * Cfg::ACC_STRIDE + c0), acc);
#pragma unroll
          for (int q = 0; q < PPR; ++q)
            sts128(stg + (uint32_t)lane * kB2EpiRow + 16u * q, __uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]),
                   __uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3]));
          __syncwarp();
          store_staged(t * kBM + g4 * 32, c0, orow_l, 0, PPR);
          __syncwarp();                                             // staged rows are consumed before the next chunk lands
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bar_acc_empty + buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kB2UtilWarp) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (A.sched && tid == 0) {                                     // the last CTA to leave rearms the scheduler slot
    __threadfence();
    if (atomicAdd(A.sched + 15, 1u) == gridDim.x * gridDim.y - 1u)
      for (int i = 0; i < 16; ++i) A.sched[i] = 0u;
  }
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

// ---------------------------------------------------------------------------------------------
// Row grouping (s2d_table_group_rows): cut the EXECUTED work of the tile kernel.
//
// The kernel multiplies whole 128-row tiles per kernel offset, although only 17 % (stage 0) .. 64 % (stage 3) of the
// (row, offset) pairs of a LiDAR scene exist: a voxel on a flat surface has no neighbour above or below, one on a thin
// vertical structure none to the sides.  In scan order a tile mixes all kinds of rows, so nearly every offset is live for
// some row of every tile (79-96 %).  Output rows may be processed in ANY order, though: the kernel scatters row p of the
// permuted table to out_rows[p], so the results are bit-identical.  Rows are therefore grouped by WHICH of the nine
// (dz, dy) offset triples have a neighbour (9-bit key, stable counting sort, so spatial locality survives inside a group);
// tiles become homogeneous and whole triples of offsets are skipped through the tile masks: live (tile, offset) pairs drop
// to 35 % / 62 % / 68 % / 76 % on the four SubM stages and to 23-49 % on the strided layers.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGrpRows) group_keys_hist_kernel(const int* __restrict__ tbl, int stride, int K, int n,
                                                                   unsigned short* __restrict__ keys, int* __restrict__ counts) {
  __shared__ __align__(16) unsigned short s_cnt[32][kGrpBuckets];
  const int row = blockIdx.x * kGrpRows + threadIdx.x;
  unsigned key = 0;
  if (row < n) {
    for (int k = 0; k < K; ++k)
      if (__ldg(tbl + (size_t)k * stride + row) >= 0) key |= 1u << (k / 3);
    keys[row] = (unsigned short)key;
  }
  group_block_counts(key, row < n, s_cnt);
  if (threadIdx.x < kGrpBuckets) {
    int total = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) total += s_cnt[w][threadIdx.x];
    counts[(size_t)blockIdx.x * kGrpBuckets + threadIdx.x] = total;
  }
}

// tbl_out[k][p] = tbl[k][perm[p]] and the live-offset mask of every 128-row tile of tbl_out, one block per tile
__global__ void __launch_bounds__(128) group_permute_table_kernel(const int* __restrict__ tbl, int stride, int K, int n,
                                                                  const int* __restrict__ perm, int* __restrict__ out,
                                                                  int out_stride, int* __restrict__ masks) {
  __shared__ unsigned s_mask;
  if (threadIdx.x == 0) s_mask = 0u;
  __syncthreads();
  const int p = blockIdx.x * 128 + threadIdx.x;
  const int src = p < n ? __ldg(perm + p) : -1;
  unsigned m = 0;
#pragma unroll 9
  for (int k = 0; k < K; ++k) {
    const int v = src >= 0 ? __ldg(tbl + (size_t)k * stride + src) : -1;
    if (p < n) out[(size_t)k * out_stride + p] = v;
    if (__ballot_sync(0xffffffffu, v >= 0)) m |= 1u << k;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicOr(&s_mask, m);
  __syncthreads();
  if (threadIdx.x == 0) masks[blockIdx.x] = (int)s_mask;
}

// dynamic group scheduler slots: [next group per blockIdx.y (<= 8) ... | CTAs that have left at [15]]; zero at load, rearmed
// by the last CTA of the launch that used the slot; consecutive launches rotate over the slots so that launches running
// concurrently on different streams never share one
constexpr int kB2SchedSlots = 256;
__device__ unsigned int g_b2_sched[kB2SchedSlots][16];
// next slot, shared by every instantiation of the kernel (one rotation for the whole library)
static unsigned int* b2_sched_slot() {
  static unsigned int* slots = nullptr;
  static std::atomic<unsigned> next{0};
  if (!slots && cudaGetSymbolAddress(reinterpret_cast<void**>(&slots), g_b2_sched) != cudaSuccess) return nullptr;
  return slots + (size_t)(next.fetch_add(1) % kB2SchedSlots) * 16;
}
static int g_b2_variant = 0;
static int g_b2_dbg = 0;
static long long* g_b2_prof = nullptr;

static int b2_cout_block(int Cout) {
  return Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : (Cout % 16 == 0 ? 16 : 0)));
}
bool bf2_supported(int Cin, int Cout) { return (Cin == 16 || (Cin >= 32 && Cin % 32 == 0)) && b2_cout_block(Cout) != 0; }

template <int COUT, int T, int S, bool TMA, bool SWAP = false, int PW = 1>
static int launch_b2(const B2Args& a, const CUtensorMap& map, int Cout, cudaStream_t st) {
  using Cfg = B2Cfg<COUT, T, S, SWAP, PW>;
  static bool configured = false;
  if (!configured) {
    S2D_CUDA(cudaFuncSetAttribute(conv_bf2_kernel<COUT, T, S, TMA, SWAP, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    configured = true;
  }
  B2Args b = a;
  S2D_REQUIRE(a.ksteps * a.nchunk <= kB2ListCap && a.nchunk <= 128,
              "s2d_conv_fwd(bf16x2): %d x %d contraction blocks per tile group exceed %d (or more than 128 chunks)", a.ksteps,
              a.nchunk, kB2ListCap);
  // persistent CTAs, one per SM: CTA b works on the tile groups b, b + grid, b + 2 grid, ...
  b.n_tiles = div_up(a.n_out, kBM);
  b.n_groups = div_up(b.n_tiles, T);
  const int gy = Cout / COUT;
  int gx = kNumSMs * Cfg::CTAS_PER_SM / gy;
  if (gx < 1) gx = 1;
  if (gx > b.n_groups) gx = b.n_groups;
  const dim3 grid(gx, gy);
  b.sched = nullptr;
  if (a.tile_masks && gy <= 8 && !(g_b2_variant & 4)) {   // masked (grouped) launches: groups differ in cost -> dynamic scheduler
    b.sched = b2_sched_slot();
    S2D_REQUIRE(b.sched, "s2d_conv_fwd(bf16x2): cannot resolve the scheduler slots");
  }
  conv_bf2_kernel<COUT, T, S, TMA, SWAP, PW><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(b, map);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

// 4-D tensor map over a regular map of split rows [B, H, W, ld words]: box = 32 words x kGridTW x kGridTH x 1, 128B swizzle,
// zeros outside (the convolution's padding and the ragged last tiles).
static int make_grid_map(const void* base, int ld_words, int C, int B, int H, int W, CUtensorMap* map) {
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)ld_words * 4, (cuuint64_t)ld_words * 4 * W, (cuuint64_t)ld_words * 4 * W * H};
  const cuuint32_t box[4] = {32u, (cuuint32_t)kGridTW, (cuuint32_t)kGridTH, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {             // resolved at run time: the library must load on a machine without a driver (build / ABI checks)
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    S2D_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    S2D_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "s2d_conv_fwd_grid: cuTensorMapEncodeTiled not available in this driver");
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("s2d_conv_fwd_grid: cuTensorMapEncodeTiled failed (%d) for base=%p ld=%d C=%d B=%d H=%d W=%d", (int)r, base, ld_words,
              C, B, H, W);
    return S2D_ERR_CUDA;
  }
  return S2D_OK;
}

// 2-D tensor map over split rows [n rows, ld words]: box = 32 words x 1 row, 128B swizzle, zero fill outside -- what
// tile::gather4 needs (four such rows per instruction, a negative row index reads as zeros)
static int make_rows_map(const void* base, int ld_words, int C, int n, CUtensorMap* map) {
  const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)(n > 0 ? n : 1)};
  const cuuint64_t strides[1] = {(cuuint64_t)ld_words * 4};
  const cuuint32_t box[2] = {32u, 1u};
  const cuuint32_t estr[2] = {1u, 1u};
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  S2D_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  S2D_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "s2d_conv_fwd: cuTensorMapEncodeTiled not available in this driver");
  const CUresult r = reinterpret_cast<EncodeFn>(fn)(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides,
                                                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("s2d_conv_fwd: cuTensorMapEncodeTiled failed (%d) for rows map base=%p ld=%d C=%d n=%d", (int)r, base, ld_words, C, n);
    return S2D_ERR_CUDA;
  }
  return S2D_OK;
}

// grid_k > 0: dense-grid (TMA) mode on a [grid_b, grid_h, grid_w] map with a grid_k x grid_k kernel and padding grid_pad
static int conv_fwd_bf2_impl(const s2d_conv_params& p, int grid_b, int grid_h, int grid_w, int grid_k, int grid_pad,
                             cudaStream_t st) {
  const bool tma = grid_k > 0;
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
  S2D_REQUIRE(tma || (p.tbl_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(p.tbl) & 15) == 0),
              "s2d_conv_fwd(bf16x2): the neighbour table must be 16 B aligned with tbl_stride %% 4 == 0 (got %d)", p.tbl_stride);
  S2D_REQUIRE((unsigned long long)p.n_in * (unsigned long long)p.in_split_ld * 4ull < (1ull << 36),
              "s2d_conv_fwd: input tensor larger than 64 GiB (32-bit gather offsets in 16 B units)");
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
  a.prof = g_b2_prof;
  a.grid_tiles_x = a.grid_tiles_y = a.grid_kw = a.grid_pad = 0;
  a.g4 = 0;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (tma) {
    S2D_REQUIRE(p.Cin % 32 == 0 && p.K == grid_k * grid_k && p.out_rows && !p.tile_masks,
                "s2d_conv_fwd_grid: needs Cin %% 32 == 0, K = k*k, out_rows (s2d_grid2d_tile_rows) and no tile masks");
    a.grid_tiles_x = div_up(grid_w, kGridTW);
    a.grid_tiles_y = div_up(grid_h, kGridTH);
    a.grid_kw = grid_k;
    a.grid_pad = grid_pad;
    S2D_REQUIRE(p.n_out == grid_b * a.grid_tiles_x * a.grid_tiles_y * kBM, "s2d_conv_fwd_grid: n_out must be tiles x 128");
    const int rc = make_grid_map(p.in_split, p.in_split_ld, p.Cin, grid_b, grid_h, grid_w, &map);
    if (rc != S2D_OK) return rc;
    if (cb == 128) return (g_b2_variant & 3) == 2 ? launch_b2<128, 2, 8, true>(a, map, p.Cout, st) : launch_b2<128, 2, 8, true, true>(a, map, p.Cout, st);
    if (cb == 64) return launch_b2<64, 4, 8, true>(a, map, p.Cout, st);
    if (cb == 32) return launch_b2<32, 4, 8, true>(a, map, p.Cout, st);
    return launch_b2<16, 4, 8, true>(a, map, p.Cout, st);
  }
  const int v = g_b2_variant;
  a.g4 = 0;
  if ((v & 8) && a.kps == 1) {        // experimental: TMA gather4 producers
    const int rc = make_rows_map(p.in_split, p.in_split_ld, p.Cin, p.n_in, &map);
    if (rc != S2D_OK) return rc;
    a.g4 = 1;
  }
  // variant 0: production choice; 1: T = 2.  (Two CTAs per SM with half the stage ring each, S = 4, measured no faster: the
  // kernel is bound by shared-memory bandwidth and the tensor pipe, which both CTAs share, not by barrier latency.)
  // 128 output channels: swapped operands (variant 2: the unswapped kernel, for A/B timing).  With only a few blocks per
  // group (1x1 layers) the kernel is bound by its epilogue, and the unswapped one (8 x STS.128 per item instead of 32 x
  // STS.32) is the faster of the two there.
  if (cb == 128) {
    if ((v & 3) == 2 || a.ksteps * a.nchunk < 16) return launch_b2<128, 2, 8, false>(a, map, p.Cout, st);
    // variant 16: two producer warps per stage -- measured no faster (0.390 ms either way on the 128 -> 128 layer): the
    // gather is bound by what the SM can pull from L2 for a map of this size, not by the fill latency of a stage
    if ((v & 16) && !a.g4) return launch_b2<128, 2, 8, false, true, 2>(a, map, p.Cout, st);
    return launch_b2<128, 2, 8, false, true>(a, map, p.Cout, st);
  }
  if (cb == 64) return (v & 3) == 1 ? launch_b2<64, 2, 8, false>(a, map, p.Cout, st) : launch_b2<64, 4, 8, false>(a, map, p.Cout, st);
  if (cb == 32) return (v & 3) == 1 ? launch_b2<32, 2, 8, false>(a, map, p.Cout, st) : launch_b2<32, 4, 8, false>(a, map, p.Cout, st);
  return (v & 3) == 1 ? launch_b2<16, 2, 8, false>(a, map, p.Cout, st) : launch_b2<16, 4, 8, false>(a, map, p.Cout, st);
}

int conv_fwd_bf2(const s2d_conv_params& p, cudaStream_t st) { return conv_fwd_bf2_impl(p, 0, 0, 0, 0, 0, st); }

// out_rows of the dense-grid mode: tile-order row -> pixel row (b * H + y) * W + x, or -1 outside the map
__global__ void __launch_bounds__(128) grid_tile_rows_kernel(int B, int H, int W, int tiles_x, int tiles_y, int* __restrict__ rows) {
  const int tile = blockIdx.x, r = threadIdx.x;
  const int b = tile / (tiles_x * tiles_y), rem = tile - b * tiles_x * tiles_y;
  const int y = (rem / tiles_x) * kGridTH + r / kGridTW, x = (rem % tiles_x) * kGridTW + r % kGridTW;
  rows[(size_t)tile * kBM + r] = (y < H && x < W) ? (b * H + y) * W + x : -1;
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

// which instantiation conv_fwd_bf2 launches (tuning aid; not part of the public header): low two bits 1 = two tiles per group
// for Cout <= 64, 2 = the unswapped kernel for Cout = 128; bit 4 = static round robin instead of the dynamic group scheduler;
// bit 8 = TMA gather4 producers (experiment, 2x slower); bit 16 = two producer warps per stage at Cout = 128 (no gain)
extern "C" void s2d_debug_bf2_variant(int v) { g_b2_variant = v; }
extern "C" void s2d_debug_bf2_flags(int f) { g_b2_dbg = f; }
extern "C" void s2d_debug_bf2_prof(long long* p) { g_b2_prof = p; }

extern "C" int s2d_table_tile_masks(const int* tbl, int tbl_stride, int K, int n_rows, int* tile_masks, void* stream) {
  S2D_REQUIRE(K >= 1 && K <= 31 && n_rows >= 0 && tbl_stride >= n_rows, "s2d_table_tile_masks: bad argument");
  if (n_rows == 0) return S2D_OK;
  S2D_REQUIRE(tbl && tile_masks, "s2d_table_tile_masks: null argument");
  tile_masks_kernel<<<div_up(n_rows, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(tbl, tbl_stride, K, n_rows, tile_masks);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" size_t s2d_table_group_rows_workspace_bytes(int n_rows) {
  if (n_rows < 0) return 0;
  return group_workspace_bytes(n_rows);
}

extern "C" int s2d_table_group_rows(const int* tbl, int tbl_stride, int K, int n_rows, int* perm, int* tbl_out,
                                    int out_stride, int* tile_masks, void* workspace, size_t workspace_bytes, void* stream) {
  S2D_REQUIRE(K >= 1 && K <= 27 && n_rows >= 0 && tbl_stride >= n_rows && out_stride >= n_rows,
              "s2d_table_group_rows: bad argument (K = %d must be <= 27)", K);
  if (n_rows == 0) return S2D_OK;
  S2D_REQUIRE(tbl && perm && tbl_out && tile_masks && workspace, "s2d_table_group_rows: null argument");
  S2D_REQUIRE(workspace_bytes >= s2d_table_group_rows_workspace_bytes(n_rows), "s2d_table_group_rows: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = div_up(n_rows, kGrpRows);
  int* counts = static_cast<int*>(workspace);
  int* tails = counts + (size_t)nblk * kGrpBuckets;
  unsigned short* keys = reinterpret_cast<unsigned short*>(tails + kGrpSegs * kGrpBuckets);
  group_keys_hist_kernel<<<nblk, kGrpRows, 0, st>>>(tbl, tbl_stride, K, n_rows, keys, counts);
  group_scan(counts, nblk, tails, st);
  group_scatter_kernel<<<nblk, kGrpRows, 0, st>>>(keys, n_rows, nblk, counts, tails, perm);
  group_permute_table_kernel<<<div_up(n_rows, 128), 128, 0, st>>>(tbl, tbl_stride, K, n_rows, perm, tbl_out, out_stride,
                                                                   tile_masks);
  S2D_LAUNCH_CHECK();
  count_launches(5);
  return S2D_OK;
}

extern "C" int s2d_grid2d_tile_rows_count(int B, int H, int W) {
  return (B < 1 || H < 1 || W < 1) ? 0 : B * div_up(H, kGridTH) * div_up(W, kGridTW) * kBM;
}

extern "C" int s2d_grid2d_tile_rows(int B, int H, int W, int* rows, void* stream) {
  S2D_REQUIRE(B >= 1 && H >= 1 && W >= 1 && rows, "s2d_grid2d_tile_rows: bad argument");
  const int tx = div_up(W, kGridTW), ty = div_up(H, kGridTH);
  grid_tile_rows_kernel<<<B * tx * ty, 128, 0, static_cast<cudaStream_t>(stream)>>>(B, H, W, tx, ty, rows);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_conv_fwd_grid(const s2d_conv_params* params, int B, int H, int W, int k, int pad, void* stream) {
  S2D_REQUIRE(params, "s2d_conv_fwd_grid: null params");
  const s2d_conv_params& p = *params;
  S2D_REQUIRE(p.precision == S2D_PRECISION_BF16X2, "s2d_conv_fwd_grid: only S2D_PRECISION_BF16X2");
  S2D_REQUIRE(B >= 1 && H >= 1 && W >= 1 && k >= 1 && k * k <= 27 && pad >= 0 && p.n_in == B * H * W,
              "s2d_conv_fwd_grid: bad grid (B=%d H=%d W=%d k=%d pad=%d n_in=%d)", B, H, W, k, pad, p.n_in);
  S2D_REQUIRE(p.act >= S2D_ACT_NONE && p.act <= S2D_ACT_GELU && p.weights, "s2d_conv_fwd_grid: bad argument");
  return conv_fwd_bf2_impl(p, B, H, W, k, pad, static_cast<cudaStream_t>(stream));
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
