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
#include "tc_ptx.cuh"

namespace s2d {

constexpr int kTcMaxK = 27;

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
constexpr int kGroups = 2;                 // producer groups: group g gathers / converts the steps s = g (mod kGroups)
constexpr int kGroupWarps = 8;             // 8 warps x 16 rows = one 128-row step
constexpr int kProducerWarps = kGroups * kGroupWarps;
constexpr int kLoaderWarp = kProducerWarps;
constexpr int kMmaWarp0 = kProducerWarps + 1;   // first MMA warp; MMA warp m issues the MMAs of row tile t = m
constexpr int kMmaWarps = 4;                     // = max T: a single issuing thread cannot keep the tensor pipe fed
constexpr int kTcThreads = 32 * (kProducerWarps + 1 + kMmaWarps);
constexpr int kNbrSlots = 4;
constexpr int kMaxKps = 2;   // kernel offsets folded into one 32-channel contraction step (Cin = 16: two)

// A tile element (row m, channel ch of the 32-channel chunk) lives in TMEM lane m, column a_col(ch).  The
// permutation makes the 8 columns a thread owns in a tcgen05.st.16x256b.x4 fragment
// (columns 8n + 2j + e, j = lane % 4) equal to 8 CONTIGUOUS channels [8j, 8j+8) of its row, so the gather is
// two LDG.128 per row.  The weight tiles are packed with the same K permutation (pack_weights_kernel).
__host__ __device__ constexpr int a_col_of_channel(int ch) { return 8 * ((ch % 8) / 2) + 2 * (ch / 8) + (ch % 2); }
// Same idea for the BF16 correction operand: two values per 32-bit column, 16 columns per 32 channels; the
// thread's 4 columns {2j, 2j+1, 8+2j, 8+2j+1} hold channels 8j..8j+7 at k positions 4j..4j+3, 16+4j..16+4j+3.
__host__ __device__ constexpr int bf16_kpos_of_channel(int ch) { return 16 * ((ch % 8) / 4) + 4 * (ch / 8) + (ch % 4); }

// PASSES: 1 = single TF32 pass; 3 = split TF32 (Alo*Bhi + Ahi*Blo + Ahi*Bhi, three TF32 MMAs per K slice);
//         2 = TF32 main term + BF16 correction terms: Ahi*Bhi in TF32, and [Alo | A] x [B ; Blo] as ONE BF16
//             contraction of twice the depth (BF16 runs at twice the TF32 rate, so the corrections cost one TF32
//             pass instead of two).  The correction terms are 2^-11 of the result, so their own BF16 rounding
//             (2^-9 relative) leaves an error of ~2^-20 per product: fp32-level accuracy at 2/3 of the tensor time.
template <int COUT, int PASSES>
struct TcCfg {
  static constexpr int NPART = PASSES == 1 ? 1 : 2;
  static constexpr int T = COUT >= 128 ? 2 : 4;        // 128-row tiles per CTA sharing every weight tile
  static constexpr int ROWS = T * kBM;
  static constexpr int B_TILE = COUT * kBK * 4;        // 2 / 4 / 8 / 16 KB (COUT rows x 128 B)
  static constexpr int B_STAGE = NPART * B_TILE;
  static constexpr int SB = COUT >= 128 ? 3 : 4;       // weight-tile ring (shared memory)
  static constexpr int ACC_STRIDE = COUT < 32 ? 32 : COUT;   // TMEM allocation granule
  static constexpr int ACC_COLS = T * ACC_STRIDE;      // TMEM: accumulators ...
  static constexpr int A_COLS = NPART * kBK;           // ... and one gathered A stage: 32 TF32 columns (+ 32 more)
  static constexpr int SA_RAW = (512 - ACC_COLS) / A_COLS;
  static constexpr int SA_CAP = SA_RAW > 8 ? 8 : SA_RAW;
  static constexpr int SA = SA_CAP - SA_CAP % kGroups;   // gathered-tile ring (TMEM); see "phase aliasing" below
  static constexpr int TMEM_COLS = 512;
  static constexpr int DEPTH = COUT >= 128 ? 3 : 4;    // gathered steps in flight per producer warp (cp.async ring)
  static constexpr int A_WARP_STAGE = 16 * 128;        // 16 rows x 128 B per producer warp and step
  static constexpr int A_RING_BYTES = kProducerWarps * DEPTH * A_WARP_STAGE;
  static constexpr int EPC = COUT < 32 ? COUT : 32;    // accumulator columns per epilogue tcgen05.ld
  static constexpr int NBR_BYTES = kNbrSlots * kMaxKps * ROWS * 4;
  static constexpr int SMEM_BYTES = SB * B_STAGE + A_RING_BYTES + NBR_BYTES + 1024 + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(COUT % 16 == 0 && COUT >= 16 && COUT <= 128, "UMMA N constraint for M=128 / TMEM budget");
  // mbarrier waits are parity waits: a waiter must never be more than one phase behind a barrier it waits on.
  //  * producers: a stage's uses are the steps s = stage (mod SA); with SA a multiple of kGroups they all belong to ONE
  //    producer group, whose warps therefore see every phase of that stage's barriers (an odd SA broke this);
  //  * MMA warps: tile t consumes the steps s = t (mod Tr).  It sees every phase of a stage iff Tr divides SA; for
  //    even Tr the phases it skips were delivered earlier by the same producer group (in order), so they are complete.
  //    An odd Tr that does not divide SA (3 with SA = 4 or 8) is NOT safe: launch_tc never hands out 3 tiles then.
  static constexpr bool TR3_OK = SA % 3 == 0;
  static_assert(SA >= 2 && SA >= kGroups && SA % kGroups == 0, "gathered-tile ring: see phase aliasing");
  static_assert(T >= kGroups, "producer groups split the steps of an offset step");
};

// Arguments of one launch.  Rows may be strided (in_ld / out_ld / res_ld, in floats) so that layers can read
// from and write into channel slices of wider buffers (concatenations); blockIdx.y selects a block of COUT
// output channels (weights are packed per block); out_rows optionally remaps output rows (sub-pixel
// transposed convolutions).  A contraction step covers 32 input channels: one 32-channel chunk of one kernel
// offset (kps = 1, nchunk = Cin/32) or the 16 channels of two consecutive offsets (kps = 2, Cin = 16).
// Ablation switches (tools/ablate_spconv.py) exist only in builds with -DS2D_TC_ABLATE; they cost instructions in
// the producer loop, which is issue bound.
#ifdef S2D_TC_ABLATE
#define S2D_DBG(args, bit) ((args).dbg & (bit))
#else
#define S2D_DBG(args, bit) 0
#endif

struct ConvArgs {
  const float* in;
  const float* packed;
  const int* tbl;
  const float* scale;
  const float* shift;
  const float* residual;
  float* out;
  const int* out_rows;
  int in_ld, out_ld, res_ld, tbl_stride, n_out, K, nchunk, kps, ksteps, act, res_after_act, use_tma;
  int tile_unit, unit_base, unit_rem, n_tiles;   // CTA b owns (unit_base + (b < unit_rem)) * tile_unit consecutive tiles
  int t_min;                 // tiles a CTA executes at least (virtual empty tiles) so that Tr * nchunk >= kGroups
  int dbg;   // ablation switches for tools/ablate_spconv.py (0 in production): 1 no gather, 2 no TMEM store, 4 no MMA
};

__device__ __forceinline__ float apply_act(float y, int act) {
  if (act == S2D_ACT_RELU) return fmaxf(y, 0.f);
  if (act == S2D_ACT_GELU) return 0.5f * y * (1.f + erff(y * 0.70710678118654752440f));   // exact GELU (torch default)
  return y;
}

// One CTA = T row tiles.  Loop nest: offset step kk > channel chunk c > tile t; the weight tile of (kk,c) is
// loaded once and feeds T MMA groups (one per tile accumulator).  The gathered A operand never touches
// shared memory: the producer warps write it straight into TMEM (tcgen05.st) and the MMAs run in TS mode,
// so shared-memory bandwidth only carries the small weight slices.
template <int COUT, int PASSES>
__global__ void __launch_bounds__(kTcThreads, 1) spconv_tc_kernel(const __grid_constant__ ConvArgs A,
                                                                  const __grid_constant__ CUtensorMap in_map) {
  using Cfg = TcCfg<COUT, PASSES>;
  constexpr int SA = Cfg::SA, SB = Cfg::SB, T = Cfg::T;
  constexpr int B_TILE = Cfg::B_TILE, B_STAGE = Cfg::B_STAGE, A_COLS = Cfg::A_COLS;
  const int NCHUNK = A.nchunk, K = A.K, KS = A.ksteps, kps = A.kps, n_out = A.n_out;
  const size_t blk_tiles = (size_t)blockIdx.y * KS * NCHUNK;           // weight tiles before this channel block
  const float* __restrict__ packed = A.packed + blk_tiles * 2 * (B_TILE / 4);
  const int* __restrict__ tbl = A.tbl;
  const int tbl_stride = A.tbl_stride;
  const int cblk = blockIdx.y * COUT;                                  // first output channel of this block

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_ring = smem;                                       // SB x [B_hi | B_lo or B_bf16]
  uint8_t* a_ring = b_ring + SB * B_STAGE;                      // 8 warps x DEPTH x [16 rows x 128 B], 128B-swizzled
  int* s_nbr = reinterpret_cast<int*>(a_ring + Cfg::A_RING_BYTES);   // kNbrSlots x kMaxKps x [T*128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_nbr) + Cfg::NBR_BYTES);
  uint64_t* bar_a_full = bars;
  uint64_t* bar_a_empty = bar_a_full + SA;
  uint64_t* bar_b_full = bar_a_empty + SA;
  uint64_t* bar_b_empty = bar_b_full + SB;
  uint64_t* bar_n_full = bar_b_empty + SB;
  uint64_t* bar_n_empty = bar_n_full + kNbrSlots;
  uint64_t* bar_accum = bar_n_empty + kNbrSlots;
  uint64_t* bar_gather = bar_accum + 1;                         // kProducerWarps x DEPTH: one per warp and ring stage
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_gather + kProducerWarps * Cfg::DEPTH);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Balanced tiling: the 128-row tiles are dealt evenly over a grid that is a multiple of the SM count, so the last
  // wave is as loaded as the others (a CTA's time is proportional to its tile count Tr <= T).
  const int bx = (int)blockIdx.x;
  const int t_alloc = (A.unit_base + (bx < A.unit_rem ? 1 : 0)) * A.tile_unit;
  const int tile_first = (bx * A.unit_base + min(bx, A.unit_rem)) * A.tile_unit;
  if (t_alloc == 0 || tile_first >= A.n_tiles) return;
  const int Tr = max(t_alloc, A.t_min);       // tiles past the CTA's real ones are virtual: no rows loaded, none stored
  const int tile0 = tile_first * kBM;
  const int row_end = min(n_out, tile0 + t_alloc * kBM);
  const int nsteps = KS * NCHUNK * Tr;

  if (warp == kMmaWarp0 && lane == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(smem_u32(bar_a_full + s), kGroupWarps);           // one arrival per producer warp of the step's group
      mbar_init(smem_u32(bar_a_empty + s), 1);                    // one tcgen05.commit
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(smem_u32(bar_b_full + s), 1);                     // arrive.expect_tx of the loader
      mbar_init(smem_u32(bar_b_empty + s), Tr);                   // one tcgen05.commit per row tile
    }
    for (int s = 0; s < kNbrSlots; ++s) {
      mbar_init(smem_u32(bar_n_full + s), 32);                    // every loader lane arrives
      mbar_init(smem_u32(bar_n_empty + s), kProducerWarps);
    }
    mbar_init(smem_u32(bar_accum), Tr);
    for (int s = 0; s < kProducerWarps * Cfg::DEPTH; ++s) mbar_init(smem_u32(bar_gather + s), 1);
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
    // channels [8*(l%4), +8) of the step's 32 contraction channels (see a_col_of_channel).
    const int grp = warp / kGroupWarps, gw = warp % kGroupWarps;
    const int row16 = 32 * (gw & 3) + 16 * (gw >> 2);
    const int rl = lane >> 2;                                // fragment rows rl and rl + 8 of the warp's 16
    const int j = lane & 3;
    const int o = lane >> 3, c = lane & 7;                   // gather: lane = (row octet, 16 B chunk of the row)
    const int sub = kps == 2 ? (c >> 2) : 0;                 // which of the step's offsets this lane gathers
    const int c4 = kps == 2 ? (c & 3) : c;                   // float4 index inside the source row chunk
    const uint32_t bar_a_full0 = smem_u32(bar_a_full), bar_a_empty0 = smem_u32(bar_a_empty);
    const uint32_t bar_n_full0 = smem_u32(bar_n_full), bar_n_empty0 = smem_u32(bar_n_empty);
    const uint32_t tmem_mine = tmem_a0 + ((uint32_t)row16 << 16);
    const int* nbr_mine = s_nbr + sub * Cfg::ROWS + row16 + 4 * o;   // this lane gathers rows 4o .. 4o+3 of the warp's 16
    constexpr int DEPTH = Cfg::DEPTH;
    const uint32_t ring0 = smem_u32(a_ring) + (uint32_t)warp * (DEPTH * Cfg::A_WARP_STAGE);
    // gather destination of this lane inside a stage (row 4o + i, chunk c, 128B swizzle) and fragment sources:
    // dst_i = (g_off + i*128) ^ (i << 4), because (4o + i) & 7 = 4*(o & 1) + i for i < 4
    const uint32_t g_off = (uint32_t)o * 512u + (uint32_t)((c ^ (4 * (o & 1))) << 4);
    const uint32_t f_off0 = (uint32_t)rl * 128u + (uint32_t)(((2 * j) ^ rl) << 4);
    const uint32_t f_off1 = (uint32_t)rl * 128u + (uint32_t)(((2 * j + 1) ^ rl) << 4);

    // The gather runs DEPTH steps ahead of the conversion: every lane copies 16 B pieces of the neighbour rows
    // straight into the warp's shared-memory ring with cp.async (a missing neighbour is a zero-fill), so the
    // number of rows in flight is bounded by the ring, not by registers or scoreboards.  A warp gathers exactly
    // the 16 rows it later converts, so only warp-level synchronisation is needed.
    // This warp's steps are s = grp, grp + kGroups, ...; step s = (lk * NCHUNK + lc) * Tr + lt.  The host guarantees
    // Tr * NCHUNK >= kGroups, so every group has a step in every offset step (the slot protocol relies on it).
    int lk = 0, lc = 0, lt = grp;                // gather stream position (offset step, chunk, tile)
    while (lt >= Tr) { lt -= Tr; ++lc; }
    int lk_ready = -1;                           // last offset step whose neighbour slice this warp has waited for
    int gstage = 0, cstage = 0;                  // ring positions of the gather / convert streams
    uint32_t cphase = 0;                         // mbarrier phase of the convert stream (TMA gather)
    int sstage = grp;                            // TMEM ring position of step s: s % SA
    uint32_t sphase = 1;                         // first pass over the TMEM ring: slots are free
    // byte offsets are 32 bit (the host checks n_in * in_ld * 4 < 4 GiB); neighbour indices are read with ld.shared
    const char* in_bytes = reinterpret_cast<const char*>(A.in);
    const uint32_t row_bytes = (uint32_t)A.in_ld * 4u;
    const uint32_t nbr_addr0 = smem_u32(nbr_mine);
    // The four neighbour indices of the next gather are fetched with ONE LDS.128 well before they are needed
    // (gather_prefetch at the top of a conversion step, gather_issue at its end), which takes the index load and
    // its shared-memory queueing delay off the critical path.
    int4 pidx = make_int4(-1, -1, -1, -1);
    auto gather_prefetch = [&]() {
      const int slot = lk & (kNbrSlots - 1);
      if (lk != lk_ready) { mbar_wait(bar_n_full0 + 8 * slot, (lk / kNbrSlots) & 1); lk_ready = lk; }
      pidx = lds128i(nbr_addr0 + (uint32_t)(slot * (kMaxKps * Cfg::ROWS) + lt * kBM) * 4u);
    };
    const bool use_tma = A.use_tma != 0;         // kps == 1: rows are whole 128 B chunks -> TMA gather4
    const uint32_t bar_g0 = smem_u32(bar_gather + warp * DEPTH);
    auto gather_issue = [&]() {
      const int slot = lk & (kNbrSlots - 1);
      if (use_tma) {
        // lane 8o issues ONE gather4 for rows 4o..4o+3 (their indices are its pidx): 4 instructions per warp and
        // step instead of 4 x 32 cp.async lanes, and the copies run on the TMA engine, not through the LSU pipe.
        const uint32_t bar = bar_g0 + 8u * (uint32_t)gstage;
        if (lane == 0) mbar_arrive_expect_tx(bar, S2D_DBG(A, 1) ? 0u : (uint32_t)Cfg::A_WARP_STAGE);
        __syncwarp();
        if (c == 0 && !S2D_DBG(A, 1))
          tma_gather4(ring0 + (uint32_t)gstage * Cfg::A_WARP_STAGE + (uint32_t)o * 512u, &in_map, lc * kBK, pidx.x, pidx.y,
                      pidx.z, pidx.w, bar);
      } else {
        const uint32_t dst0 = ring0 + (uint32_t)gstage * Cfg::A_WARP_STAGE + g_off;
        const uint32_t cbytes = (uint32_t)(lc * kBK + c4 * 4) * 4u;
        const int idx[4] = {pidx.x, pidx.y, pidx.z, pidx.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (S2D_DBG(A, 1)) break;
          const uint32_t off = (uint32_t)max(idx[i], 0) * row_bytes + cbytes;
          const uint32_t dst = (dst0 + (uint32_t)i * 128u) ^ ((uint32_t)i << 4);
          cp_async16_zfill(dst, in_bytes + off, idx[i] >= 0 ? 16u : 0u);
        }
        cp_async_commit();
      }
      if (++gstage == DEPTH) gstage = 0;
      lt += kGroups;
      while (lt >= Tr) {
        lt -= Tr;
        if (++lc == NCHUNK) {
          lc = 0;
          __syncwarp();                          // every lane has read this step's indices
          if (lane == 0) mbar_arrive(bar_n_empty0 + 8 * slot);
          ++lk;
        }
      }
    };
    auto convert_step = [&](bool more) {
      if (more) gather_prefetch();
      if (use_tma) {
        mbar_wait(bar_g0 + 8u * (uint32_t)cstage, cphase);   // the four gather4 of the oldest step have landed
      } else {
        cp_async_wait<DEPTH - 1>();              // this lane's pieces of the oldest step in flight have landed
        __syncwarp();                            // ... and so have the other lanes'
      }
      const uint32_t src = ring0 + (uint32_t)cstage * Cfg::A_WARP_STAGE;
      if (++cstage == DEPTH) { cstage = 0; cphase ^= 1; }
      float4 v[4];
      if (S2D_DBG(A, 8)) {
        v[0] = v[1] = v[2] = v[3] = make_float4(1.f, 2.f, 3.f, 4.f);
      } else {
        v[0] = lds128(src + f_off0); v[1] = lds128(src + f_off1);
        v[2] = lds128(src + f_off0 + 1024u); v[3] = lds128(src + f_off1 + 1024u);
      }
      // fragment order of tcgen05.st.16x256b.x4: reg 4n+e = row rl, column 8n+2j+e; reg 4n+2+e = row rl+8
      const float xa[8] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w};
      const float xb[8] = {v[2].x, v[2].y, v[2].z, v[2].w, v[3].x, v[3].y, v[3].z, v[3].w};
      uint32_t hi[16], lo[16];
      if constexpr (PASSES == 2) {
        float la[8], lb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float ah = tf32_trunc(xa[i]), bh = tf32_trunc(xb[i]);   // x = hi + lo exactly; one LOP each (lo < 2^-10 |x|,
                                                                          // its BF16 truncation then costs ~2^-18 |x|)
          hi[4 * (i >> 1) + (i & 1)] = __float_as_uint(ah);
          hi[4 * (i >> 1) + 2 + (i & 1)] = __float_as_uint(bh);
          la[i] = xa[i] - ah;
          lb[i] = xb[i] - bh;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i0 = 4 * h + 2 * e;
            lo[4 * h + e] = pack_bf16x2(la[i0], la[i0 + 1]);             // columns [0,16): A_lo  (x B)
            lo[4 * h + 2 + e] = pack_bf16x2(lb[i0], lb[i0 + 1]);
            lo[4 * (h + 2) + e] = pack_bf16x2(xa[i0], xa[i0 + 1]);       // columns [16,32): A    (x B_lo)
            lo[4 * (h + 2) + 2 + e] = pack_bf16x2(xb[i0], xb[i0 + 1]);
          }
      } else {
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float a = xa[2 * n + e], cc = xb[2 * n + e];
            if constexpr (PASSES == 3) {
              const float ah = tf32_trunc(a), ch = tf32_trunc(cc);   // x = hi + lo exactly
              hi[4 * n + e] = __float_as_uint(ah);      lo[4 * n + e] = __float_as_uint(a - ah);
              hi[4 * n + 2 + e] = __float_as_uint(ch);  lo[4 * n + 2 + e] = __float_as_uint(cc - ch);
            } else {
              hi[4 * n + e] = __float_as_uint(tf32_round(a));
              hi[4 * n + 2 + e] = __float_as_uint(tf32_round(cc));
            }
          }
      }
      mbar_wait(bar_a_empty0 + 8 * sstage, sphase);
      tc_fence_after();
      const uint32_t dst = tmem_mine + (uint32_t)(sstage * A_COLS);
      if (!S2D_DBG(A, 2)) {
        tmem_st_16x256b_x4(dst, hi);
        if constexpr (PASSES != 1) tmem_st_16x256b_x4(dst + kBK, lo);
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();                              // also: every lane is done reading the ring stage
      if (lane == 0) mbar_arrive(bar_a_full0 + 8 * sstage);      // one arrival per producer warp and stage
      sstage += kGroups;
      if (sstage >= SA) { sstage -= SA; sphase ^= 1; }
    };

    const int my_steps = (nsteps - grp + kGroups - 1) / kGroups;
#pragma unroll 1
    for (int i = 0; i < DEPTH; ++i) {
      if (i < my_steps) { gather_prefetch(); gather_issue(); } else if (!use_tma) cp_async_commit();
    }
#pragma unroll 1
    for (int s = 0; s < my_steps; ++s) {
      const bool more = s + DEPTH < my_steps;
      convert_step(more);
      if (more) gather_issue(); else if (!use_tma) cp_async_commit();
    }

    // ===================== epilogue (same 8 warps) =====================
    // TMEM lane == row inside a tile; warp w may touch lanes [32*(w%4), +32).
    mbar_wait(smem_u32(bar_accum), 0);
    tc_fence_after();
    const int g = warp & 3;
    // (tile, column block) units: warp>>2 = 0..3 ; T = 4: one tile each; T = 2: tile = unit & 1, column half = unit >> 1
    const int unit = warp >> 2;
    const int t_first = unit % T, c_lo = (unit / T) * (COUT / (4 / T)), c_hi = c_lo + COUT / (4 / T);
    const float* __restrict__ scale = A.scale ? A.scale + cblk : nullptr;
    const float* __restrict__ shift = A.shift ? A.shift + cblk : nullptr;
    const int act = A.act, res_after = A.res_after_act;
    constexpr int EPC = Cfg::EPC;
#pragma unroll 1
    for (int t = t_first; t < Tr; t += 4) {
      const int row = tile0 + t * kBM + g * 32 + lane;
      const int orow = (row < row_end && A.out_rows) ? __ldg(A.out_rows + row) : row;
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_hi; c0 += EPC) {
        uint32_t acc[EPC];
        tmem_ld<EPC>(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(t * Cfg::ACC_STRIDE + c0), acc);
        if (row < row_end && !S2D_DBG(A, 16)) {
          float* dst = A.out + (size_t)orow * A.out_ld + cblk + c0;
          const float* res = A.residual ? A.residual + (size_t)orow * A.res_ld + cblk + c0 : nullptr;
#pragma unroll
          for (int q = 0; q < EPC / 4; ++q) {
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
    auto load_nbr = [&](int kk) {
      const int slot = kk % kNbrSlots;
      mbar_wait(smem_u32(bar_n_empty + slot), ((kk / kNbrSlots) & 1) ^ 1);
      int v[kMaxKps][Cfg::ROWS / 32];
#pragma unroll
      for (int sb = 0; sb < kMaxKps; ++sb) {
        const int k = kk * kps + sb;
        const bool live = sb < kps && k < K && !S2D_DBG(A, 64);   // a padded offset (odd K, kps = 2) gathers nothing
        const int* src = tbl + (size_t)k * tbl_stride + tile0;
#pragma unroll
        for (int i = 0; i < Cfg::ROWS / 32; ++i)     // all loads in flight before the first store
          v[sb][i] = (live && tile0 + lane + 32 * i < row_end) ? __ldg(src + lane + 32 * i) : -1;
      }
#pragma unroll
      for (int sb = 0; sb < kMaxKps; ++sb) {
        int* dst = s_nbr + (slot * kMaxKps + sb) * Cfg::ROWS;
        if (sb < kps) {
#pragma unroll
          for (int i = 0; i < Cfg::ROWS / 32; ++i) dst[lane + 32 * i] = v[sb][i];
        }
      }
      mbar_arrive(smem_u32(bar_n_full + slot));
    };
    if (0 < KS) load_nbr(0);
    if (1 < KS) load_nbr(1);
    for (int kk = 0; kk < KS; ++kk) {
      if (kk + 2 < KS) load_nbr(kk + 2);
      if (lane == 0) {
        for (int c = 0; c < NCHUNK; ++c) {
          const int bstep = kk * NCHUNK + c;
          const int bs = bstep % SB;
          mbar_wait(smem_u32(bar_b_empty + bs), ((bstep / SB) & 1) ^ 1);
          const uint32_t bar = smem_u32(bar_b_full + bs);
          if (S2D_DBG(A, 32)) { mbar_arrive(bar); continue; }
          mbar_arrive_expect_tx(bar, B_STAGE);
          bulk_copy_g2s(smem_u32(b_ring + (size_t)bs * B_STAGE), packed + (size_t)bstep * 2 * (B_TILE / 4), B_STAGE,
                        bar);
        }
      }
      __syncwarp();
    }
  } else if (warp - kMmaWarp0 < Tr) {
    // ===================== MMA issuers: warp m owns row tile t = m =====================
    // One thread per tile issues that tile's MMAs (steps s = bstep * T + t): the issue loop of a single thread
    // (~100 instructions per step with the barrier handling) was the bottleneck of the whole kernel when one
    // warp served all T tiles.  A comes from TMEM (TS mode), B from the shared weight ring.
    if (lane == 0) {
      const int t = warp - kMmaWarp0;
      constexpr uint32_t idesc = make_idesc_tf32(kBM, COUT);
      constexpr uint32_t idesc16 = make_idesc_bf16(kBM, COUT);
      // constant high word of the SW128 K-major descriptor: SBO = 1024 B, version 1, layout SWIZZLE_128B
      constexpr uint32_t desc_hi = 64u | (1u << 14) | (2u << 29);
      const uint32_t b_lo0 = ((smem_u32(b_ring) >> 4) & 0x3FFF) | (1u << 16);
      const uint32_t bar_a_full0 = smem_u32(bar_a_full), bar_a_empty0 = smem_u32(bar_a_empty);
      const uint32_t bar_b_full0 = smem_u32(bar_b_full), bar_b_empty0 = smem_u32(bar_b_empty);
      const uint32_t d = tmem_base + (uint32_t)(t * Cfg::ACC_STRIDE);
      const int nb = KS * NCHUNK;
      int stage = t % SA, bs = 0;
      uint32_t a_phase = (uint32_t)((t / SA) & 1), b_phase = 0;
      for (int b = 0; b < nb; ++b) {
        mbar_wait(bar_b_full0 + 8 * bs, b_phase);
        mbar_wait(bar_a_full0 + 8 * stage, a_phase);
        tc_fence_after();
        const uint32_t a_hi = tmem_a0 + (uint32_t)(stage * A_COLS);
        const uint32_t b_lo = b_lo0 + (uint32_t)bs * (B_STAGE >> 4);
        const uint32_t acc0 = b == 0 ? 0u : 1u;
        if (S2D_DBG(A, 4)) {
        } else if constexpr (PASSES == 2) {
          // corrections first (small terms): 64 BF16 k positions = 32 TMEM columns, 128 B of every B row
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint64_t db = ((uint64_t)desc_hi << 32) | (b_lo + (B_TILE >> 4) + 2 * q);
            umma_bf16_ts(d, a_hi + kBK + 8 * q, db, idesc16, q == 0 ? acc0 : 1u);
          }
#pragma unroll
          for (int q = 0; q < kBK / 8; ++q) {
            const uint64_t db = ((uint64_t)desc_hi << 32) | (b_lo + 2 * q);
            umma_tf32_ts(d, a_hi + 8 * q, db, idesc, 1);
          }
        } else {
#pragma unroll
          for (int q = 0; q < kBK / 8; ++q) {   // UMMA K = 8 for TF32: 8 TMEM columns of A, 32 B of every B row
            const uint64_t db_hi = ((uint64_t)desc_hi << 32) | (b_lo + 2 * q);
            if constexpr (PASSES == 3) {
              const uint64_t db_lo = ((uint64_t)desc_hi << 32) | (b_lo + (B_TILE >> 4) + 2 * q);
              umma_tf32_ts(d, a_hi + kBK + 8 * q, db_hi, idesc, q == 0 ? acc0 : 1u);   // small terms first
              umma_tf32_ts(d, a_hi + 8 * q, db_lo, idesc, 1);
              umma_tf32_ts(d, a_hi + 8 * q, db_hi, idesc, 1);
            } else {
              umma_tf32_ts(d, a_hi + 8 * q, db_hi, idesc, q == 0 ? acc0 : 1u);
            }
          }
        }
        umma_commit(bar_a_empty0 + 8 * stage);                  // gathered tile reusable once read
        umma_commit(bar_b_empty0 + 8 * bs);                     // weight tile: one of T arrivals
        if (b == nb - 1) umma_commit(smem_u32(bar_accum));      // this tile's accumulator complete
        stage += Tr;
        while (stage >= SA) { stage -= SA; a_phase ^= 1; }
        if (++bs == SB) { bs = 0; b_phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kLoaderWarp) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// weight packing: W [K, Cin, Cout] -> per (cout block, contraction step): [tile 0 | tile 1], each CB rows x 128 B
// in the exact shared-memory image the kernel's B descriptor expects (K-major, 128B swizzle, k permuted like
// the TMEM A layout).  tile 0 = TF32(w).  tile 1 = TF32(w - tile0) for the TF32 / TF32X3 modes, or the BF16
// pair [bf16(w) | bf16(w - tile0)] (64 k positions) for the BF16-correction mode.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ W, int K, int Cin, int Cout,
                                                           int CB, int kps, int nchunk, int ksteps, int bf16c,
                                                           float* __restrict__ packed) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nstep = ksteps * nchunk;
  const long long total = (long long)(Cout / CB) * nstep * CB * kBK;
  if (idx >= total) return;
  const int kk32 = (int)(idx % kBK);
  const int n = (int)((idx / kBK) % CB);
  const int step = (int)((idx / ((long long)kBK * CB)) % nstep);
  const int blk = (int)(idx / ((long long)kBK * CB * nstep));
  int k, ci;
  if (kps == 1) {
    k = step / nchunk;
    ci = (step % nchunk) * kBK + kk32;
  } else {
    const int cs = kBK / kps;                                   // = Cin
    k = step * kps + kk32 / cs;
    ci = kk32 % cs;
  }
  const int co = blk * CB + n;
  const float w = k < K ? W[((size_t)k * Cin + ci) * Cout + co] : 0.f;
  const float hi = tf32_rna(w);
  const float lo = w - hi;
  const size_t tile = (size_t)CB * kBK;                         // floats per tile
  const size_t base = ((size_t)blk * nstep + step) * 2 * tile;
  const int col = a_col_of_channel(kk32);                       // k position the TF32 MMA sees (matches the TMEM A layout)
  const size_t pos = (sw128_chunk_offset(n, col >> 2) >> 2) + (col & 3);
  packed[base + pos] = hi;
  if (!bf16c) {
    packed[base + tile + pos] = tf32_rna(lo);
  } else {
    __nv_bfloat16* t1 = reinterpret_cast<__nv_bfloat16*>(packed + base + tile);
    const int kp = bf16_kpos_of_channel(kk32);                  // [0,32): pairs with A_lo; +32: pairs with A
    t1[(sw128_chunk_offset(n, kp >> 3) >> 1) + (kp & 7)] = __float2bfloat16_rn(w);
    t1[(sw128_chunk_offset(n, (kp + 32) >> 3) >> 1) + (kp & 7)] = __float2bfloat16_rn(lo);
  }
}

static int g_tc_debug = 0;
static int g_tc_gather = 0;   // 0: cp.async gather ring (default), 2: TMA gather4 where possible

static int cout_block(int Cout) {
  return Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : (Cout % 16 == 0 ? 16 : 0)));
}
static int kps_of(int Cin) { return Cin == 16 ? 2 : 1; }
// S2D_PRECISION_AUTO: both modes have fp32-level accuracy; the BF16-correction mode needs 2/3 of the tensor time
// but twice the conversion instructions in the producer warps, so it only wins where the layer is tensor bound.
static int resolve_precision(int precision, int Cout) {
  if (precision != S2D_PRECISION_AUTO) return precision;
  return cout_block(Cout) >= 128 ? S2D_PRECISION_TF32_BF16C : S2D_PRECISION_TF32X3;
}

static bool tc_supported(int Cin, int Cout) {
  return (Cin == 16 || (Cin >= kBK && Cin % kBK == 0)) && cout_block(Cout) != 0;
}

template <int COUT, int PASSES>
static int launch_tc(const ConvArgs& a, const CUtensorMap& in_map, int Cout, cudaStream_t st) {
  using Cfg = TcCfg<COUT, PASSES>;
  static bool configured = false;
  if (!configured) {
    S2D_CUDA(cudaFuncSetAttribute(spconv_tc_kernel<COUT, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    configured = true;
  }
  // Deal the 128-row tiles over a grid that is a whole number of waves (one CTA per SM), at most T per CTA; small
  // problems get one tile per CTA (more SMs busy).  Where 3 tiles per CTA would alias mbarrier phases (TcCfg), tiles
  // are dealt in pairs (2 or 4 per CTA).
  ConvArgs b = a;
  const int n_tiles = div_up(a.n_out, kBM);
  const int g_full = div_up(n_tiles, Cfg::T);
  int gx = div_up(g_full, kNumSMs) * kNumSMs;
  if (gx > n_tiles) gx = n_tiles;
  b.tile_unit = (!Cfg::TR3_OK && Cfg::T >= 3 && n_tiles / gx >= 2) ? 2 : 1;
  const int units = div_up(n_tiles, b.tile_unit);
  b.unit_base = units / gx;
  b.unit_rem = units % gx;
  b.n_tiles = n_tiles;
  b.t_min = a.nchunk >= kGroups ? 1 : kGroups;                      // Tr * NCHUNK >= kGroups
  const dim3 grid(gx, Cout / COUT);
  spconv_tc_kernel<COUT, PASSES><<<grid, kTcThreads, Cfg::SMEM_BYTES, st>>>(b, in_map);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

template <int COUT>
static int launch_tc_prec(const ConvArgs& a, const CUtensorMap& m, int Cout, int precision, cudaStream_t st) {
  if (precision == S2D_PRECISION_TF32X3) return launch_tc<COUT, 3>(a, m, Cout, st);
  if (precision == S2D_PRECISION_TF32_BF16C) return launch_tc<COUT, 2>(a, m, Cout, st);
  return launch_tc<COUT, 1>(a, m, Cout, st);
}

// 2-D tensor map over the input rows [n_in, Cin] (row stride in_ld floats), box = 32 columns x 1 row, 128B swizzle,
// zero fill outside: what tile::gather4 needs (four such rows per instruction).
static int make_input_map(const s2d_conv_params& p, CUtensorMap* map) {
  const cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)(p.n_in > 0 ? p.n_in : 1)};
  const cuuint64_t strides[1] = {(cuuint64_t)p.in_ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, 1u};
  const cuuint32_t estr[2] = {1u, 1u};
  // The driver entry point is resolved at run time so that the library has no load-time dependency on libcuda.so.1
  // (it must load on a machine without a driver: build / ABI checks run there).
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    S2D_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    S2D_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "s2d_conv_fwd: cuTensorMapEncodeTiled not available in this driver");
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.in), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("s2d_conv_fwd: cuTensorMapEncodeTiled failed (%d) for in=%p n_in=%d Cin=%d in_ld=%d", (int)r, (const void*)p.in,
              p.n_in, p.Cin, p.in_ld);
    return S2D_ERR_CUDA;
  }
  return S2D_OK;
}

int conv_fwd_tf32(const s2d_conv_params& p, cudaStream_t st) {
  if (!tc_supported(p.Cin, p.Cout)) {
    set_error("s2d_conv_fwd: no tcgen05 kernel for Cin=%d Cout=%d (need Cin == 16 or Cin %% 32 == 0, Cout %% 16 == 0)",
              p.Cin, p.Cout);
    return S2D_ERR_UNSUPPORTED;
  }
  S2D_REQUIRE(p.in_ld % 4 == 0 && p.out_ld % 4 == 0 && (!p.residual || p.res_ld % 4 == 0),
              "s2d_conv_fwd: row strides must be multiples of 4 floats");
  S2D_REQUIRE((unsigned long long)p.n_in * (unsigned long long)p.in_ld * 4ull < (1ull << 32),
              "s2d_conv_fwd: input tensor larger than 4 GiB (32-bit gather offsets)");
  ConvArgs a;
  a.in = p.in; a.packed = p.weights; a.tbl = p.tbl; a.scale = p.scale; a.shift = p.shift; a.residual = p.residual;
  a.out = p.out; a.out_rows = p.out_rows; a.in_ld = p.in_ld; a.out_ld = p.out_ld; a.res_ld = p.res_ld;
  a.tbl_stride = p.tbl_stride; a.n_out = p.n_out; a.K = p.K; a.kps = kps_of(p.Cin);
  a.nchunk = a.kps == 1 ? p.Cin / kBK : 1; a.ksteps = div_up(p.K, a.kps); a.act = p.act;
  a.res_after_act = p.res_after_act;
  a.dbg = g_tc_debug;
  // The TMA gather4 path (whole 128 B row chunks: kps == 1, 16 B aligned base) is functional but measured SLOWER
  // than the cp.async ring on B200 (one gather4 ~ 38 clk of TMA time: 32->32 0.60 vs 0.38 ms, 64->64 0.68 vs 0.53 ms),
  // so it is opt-in (s2d_debug_tc_gather(2), tools/microbench_spconv.py --gather 2).
  a.use_tma = (a.kps == 1 && g_tc_gather == 2 && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0) ? 1 : 0;
  alignas(64) CUtensorMap in_map;
  memset(&in_map, 0, sizeof(in_map));
  if (a.use_tma) {
    const int rc = make_input_map(p, &in_map);
    if (rc != S2D_OK) return rc;
  }
  const int cb = cout_block(p.Cout);
  const int prec = resolve_precision(p.precision, p.Cout);
  if (cb == 128) return launch_tc_prec<128>(a, in_map, p.Cout, prec, st);
  if (cb == 64) return launch_tc_prec<64>(a, in_map, p.Cout, prec, st);
  if (cb == 32) return launch_tc_prec<32>(a, in_map, p.Cout, prec, st);
  return launch_tc_prec<16>(a, in_map, p.Cout, prec, st);
}

}  // namespace s2d

using namespace s2d;

// ablation switches of the tcgen05 kernel (profiling aid, see ConvArgs::dbg); not part of the public header
extern "C" void s2d_debug_tc_flags(int flags) { g_tc_debug = flags; }
extern "C" void s2d_debug_tc_gather(int mode) { g_tc_gather = mode; }

namespace s2d {
int pack_weights_bf2(const float* W, int K, int Cin, int Cout, void* packed, cudaStream_t st);   // conv_bf2.cu
}

extern "C" int s2d_spconv_tf32_supported(int Cin, int Cout) { return tc_supported(Cin, Cout) ? 1 : 0; }

extern "C" size_t s2d_spconv_packed_bytes(int K, int Cin, int Cout) {
  if (K < 1 || !tc_supported(Cin, Cout)) return 0;
  const int kps = kps_of(Cin);
  const int nchunk = kps == 1 ? Cin / kBK : 1;
  return (size_t)div_up(K, kps) * nchunk * kBK * Cout * 2 * sizeof(float);
}

extern "C" int s2d_spconv_pack_weights(const float* W, int K, int Cin, int Cout, int precision, float* packed,
                                       void* stream) {
  S2D_REQUIRE(W && packed && K >= 1 && K <= kTcMaxK, "s2d_spconv_pack_weights: bad argument");
  S2D_REQUIRE(tc_supported(Cin, Cout), "s2d_spconv_pack_weights: unsupported shape Cin=%d Cout=%d", Cin, Cout);
  if (precision == S2D_PRECISION_BF16X2) return pack_weights_bf2(W, K, Cin, Cout, packed, static_cast<cudaStream_t>(stream));
  precision = resolve_precision(precision, Cout);
  S2D_REQUIRE(precision == S2D_PRECISION_TF32 || precision == S2D_PRECISION_TF32X3 ||
                  precision == S2D_PRECISION_TF32_BF16C,
              "s2d_spconv_pack_weights: precision %d has no packed image", precision);
  const int kps = kps_of(Cin);
  const int nchunk = kps == 1 ? Cin / kBK : 1;
  const int ksteps = div_up(K, kps);
  const long long total = (long long)ksteps * nchunk * kBK * Cout;
  pack_weights_kernel<<<div_up(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      W, K, Cin, Cout, cout_block(Cout), kps, nchunk, ksteps, precision == S2D_PRECISION_TF32_BF16C ? 1 : 0, packed);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
