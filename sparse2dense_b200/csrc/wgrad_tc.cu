// wgrad_tc.cu -- the weight gradient of every gather-convolution on the 5th-gen tensor cores.
//
//   dW[k][a][b] = sum_i G[tbl[k][i]][a] * D[d_rows ? d_rows[i] : i][b]          (G = layer input rows, D = dOut rows)
//
// replaces spconv's indice_conv_backward filter gradient (per-offset gather -> mm) and cuDNN's wgrad of the reference's dense
// layers (det3d/models/necks/rpn.py, bbox_heads/center_head.py) for channel counts that are multiples of 32; train.cu keeps
// the CUDA-core kernels for the remaining shapes (5-channel input, 1..3-channel heads, 16-channel stage).
//
// The contraction runs over ROWS, so both operands are "MN-major" for the tensor core: a gathered row is contiguous in the
// channel (M / N) dimension and the k index is the row.  The split-row format of conv_bf2.cu -- per 32-channel chunk one
// 128 B row [32 BF16 hi | 32 BF16 lo], x = hi + lo to 16 mantissa bits -- is then EXACTLY one swizzle-atom row of the
// canonical MN-major SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO)) [16 B units]: 64 M-elements = [hi | lo] of a chunk,
// LBO = distance between chunks, SBO = 1024 B between groups of eight rows.  One tcgen05.mma.kind::f16 with
// A = [G_hi | G_lo] (M = 128: two chunks) and B = [D_hi | D_lo] (N = 64 per chunk) therefore produces all four products
// hi*hi, hi*lo, lo*hi, lo*lo as quadrants of the fp32 accumulator in TMEM; the epilogue adds the quadrants.  No operand is
// ever converted or transposed in registers: the gather is cp.async straight into the operand tile.
//
//   CTA = (row chunk, kernel offset k, block of <= 4 G-chunks x <= 4 D-chunks) -> 128 x 128 channels of dW[k] for its rows
//   warps 0-7  gather 32-row stages (eight lanes per 128 B row, one row per thread: up to 8 x 16 B cp.async per stage and
//              thread), indices loaded two stages ahead, table lines prefetched into L2; a stage is published three
//              stages after it was issued (cp.async.wait_group 3 -> fence.proxy.async -> mbarrier).  Warps 0-3 then run the
//              epilogue: tcgen05.ld, quadrant sums (hi/lo lanes live in different warps: exchanged through shared memory),
//              partial result of the chunk to the workspace
//   warp 8     TMEM allocation; one elected lane issues (na + 1) / 2 x 2 MMAs (M = 128, N = 64 nb, K = 16 rows) per stage and
//              releases it with tcgen05.commit
// Row chunks bound the length of one TMEM accumulation (the tensor core truncates when it aligns addends) and
// fill the GPU; their partial results are summed in a fixed order (deterministic).
#include "tc_ptx.cuh"

namespace s2d {

constexpr int kWtRows = 32;                // gathered rows per stage = 2 MMA k-steps of 16 rows
constexpr int kWtChunk = kWtRows * 128;    // bytes of one 32-channel chunk of a stage (32 swizzled 128 B rows)
// Stages a gather thread keeps in flight behind the one it is issuing.  A stage is published when the thread issues the
// stage kWtLag later, and a stage can only be issued once the MMAs of the stage S earlier are done: the gather runs
// S - kWtLag stages ahead of the tensor core.  (First version: 64-row stages, S = 3, lag 2 -> one stage of slack, i.e. MMA,
// issue and barrier round trip fully serialised: 4.1 k clk per 64 rows against 1 k clk of MMA time.)
constexpr int kWtLag = 3;
// Eight gather warps (one row per thread and stage) + the MMA warp.  With four (two rows per thread) the gather warps were
// bound by their own instruction stream -- 64-bit address arithmetic for 16 copies per thread and stage: ncu showed them
// issuing or in fixed-latency dependency stalls ~75 % of the time -- while shared memory held room for more copies in flight.
constexpr int kWtGatherWarps = 8;
constexpr int kWtRowGroups = kWtGatherWarps * 4;          // row groups of a stage: eight lanes copy one 128 B row
constexpr int kWtThreads = 32 * (kWtGatherWarps + 1);

struct WtArgs {
  const uint32_t* g;
  const uint32_t* d;
  const int* d_rows;
  const int* tbl;
  float* partial;                          // [chunks][K][Cg][Cd]
  int g_ld, d_ld, tbl_stride, n_rows, K, Cg, Cd, ca, cb, rows_per_chunk, n_ablk;
  int tpc;                                 // kernel offsets per CTA: 4 / ca when Cg <= 64 (their G-chunks share one D tile), else 1
};

template <int NA, int NB>
struct WtCfg {
  static constexpr int STAGE = (NA + NB) * kWtChunk;
  static constexpr int S0 = (200 * 1024) / STAGE;
  static constexpr int S = S0 > 8 ? 8 : S0;
  static constexpr int MB = (NA + 1) / 2;                 // 128-lane accumulator blocks: two G-chunks x [hi | lo] each
  static constexpr int COLS = MB * 64 * NB;
  static constexpr int TMEM_COLS = COLS <= 32 ? 32 : COLS <= 64 ? 64 : COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
  static constexpr int SMEM = S * STAGE + 1024 + 256;
  static_assert(S >= kWtLag + 2 && COLS <= 512, "stage ring / TMEM budget");
};

// MN-major, 128B-swizzled shared-memory matrix descriptor: start | LBO (between 64-element groups of M / N) | SBO (between
// groups of eight k rows) | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t wt_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void wt_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void wt_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int NA, int NB>
__global__ void __launch_bounds__(kWtThreads, 1) wgrad_tc_kernel(const __grid_constant__ WtArgs A) {
  using Cfg = WtCfg<NA, NB>;
  constexpr int S = Cfg::S;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE);
  uint64_t* bar_empty = bar_full + S;
  uint64_t* bar_acc = bar_empty + S;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acc + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x, k0 = blockIdx.y * A.tpc;
  const int ablk = blockIdx.z % A.n_ablk, bblk = blockIdx.z / A.n_ablk;
  const int a0 = ablk * NA, b0 = bblk * NB;
  // A-chunk c of the CTA = G-chunk a0 + c % cpt of kernel offset k0 + c / cpt (cpt = G-chunks per offset in this CTA)
  const int cpt = A.tpc > 1 ? A.ca : min(NA, A.ca - a0);
  const int ntaps = min(A.tpc, A.K - k0);
  const int na = cpt * ntaps, nb = min(NB, A.cb - b0);
  const int row0 = chunk * A.rows_per_chunk;
  const int row_end = min(A.n_rows, row0 + A.rows_per_chunk);
  const int n_slabs = row_end > row0 ? (row_end - row0 + kWtRows - 1) / kWtRows : 0;
  const uint32_t smem0 = smem_u32(smem);

  if (warp == kWtGatherWarps) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(smem_u32(bar_full + s), 32 * kWtGatherWarps);   // every gather thread, once its copies have landed
        mbar_init(smem_u32(bar_empty + s), 1);              // tcgen05.commit of the MMAs that read the stage
      }
      mbar_init(smem_u32(bar_acc), 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < kWtGatherWarps) {
    // ===================== gather: thread = (16 B piece of a 128 B chunk row, row group); four rows per thread ==========
    // Eight lanes copy one whole 128 B row of a chunk, so a warp instruction moves four full rows (the first version gave
    // a thread half a row: 32 half-used sectors per instruction, and the kernel ran at 10 B/clk/SM).
    constexpr int RQ = kWtRows / kWtRowGroups;               // rows per thread and stage
    const int piece = tid & 7, rg = tid >> 3;                // row rg of the stage
    const char* gb = reinterpret_cast<const char*>(A.g) + (size_t)a0 * 128 + piece * 16;
    const char* db = reinterpret_cast<const char*>(A.d) + (size_t)b0 * 128 + piece * 16;
    const size_t g_row = (size_t)A.g_ld * 4, d_row = (size_t)A.d_ld * 4;
    const int* tk = A.tbl + (size_t)k0 * A.tbl_stride;
    auto indices = [&](int slab, int (&ia)[RQ][4], int (&id)[RQ]) {
#pragma unroll
      for (int q = 0; q < RQ; ++q) {
        const int i = row0 + slab * kWtRows + q * kWtRowGroups + rg;
        ia[q][0] = ia[q][1] = ia[q][2] = ia[q][3] = id[q] = -1;
        if (i < row_end) {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (t < ntaps) ia[q][t] = __ldg(tk + (size_t)t * A.tbl_stride + i);
          id[q] = A.d_rows ? __ldg(A.d_rows + i) : i;
        }
      }
    };
    // Index loads run TWO stages ahead of their use and the table lines are pulled into L2 eight stages ahead: with one
    // stage of look-ahead the gather threads spent 30 % of their samples waiting for the indices of the stage they were
    // about to copy (ncu source page: long-scoreboard stall on the first use of the index registers) -- the table is
    // streamed once per launch and offset, so every one of those loads went to DRAM.
    constexpr int kIdxAhead = 8;
    auto prefetch_table = [&](int slab) {
      if (slab >= n_slabs || (tid & 7) != 0) return;           // one lane per row group; 32 rows x 4 B = one 128 B line per tap
      const int i = row0 + slab * kWtRows;
      if (rg < ntaps) asm volatile("prefetch.global.L2 [%0];" ::"l"(tk + (size_t)rg * A.tbl_stride + i));
      if (rg == 4 && A.d_rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.d_rows + i));
    };
    // three index register sets rotate by NAME over a loop unrolled by three (a copy "cur = next" at the end of an
    // iteration would be the first use of the loaded registers and stall right there)
    int ia0[RQ][4], id0[RQ], ia1[RQ][4], id1[RQ], ia2[RQ][4], id2[RQ];
    for (int pj = 0; pj < kIdxAhead; ++pj) prefetch_table(pj);
    indices(0, ia0, id0);
    indices(1, ia1, id1);
    const bool multi = A.tpc > 1;                            // several kernel offsets per CTA: chunk c -> offset c / cpt
    // one stage: copies of slab j from the indices (ia, id); the indices of slab j + 2 go into (la, ld)
    auto stage_step = [&](int j, int (&ia)[RQ][4], int (&id)[RQ], int (&la)[RQ][4], int (&ld)[RQ]) {
      const int s = j % S;
      if (j >= S) mbar_wait(smem_u32(bar_empty + s), (uint32_t)((j / S) - 1) & 1u);
      prefetch_table(j + kIdxAhead);
      indices(j + 2, la, ld);                                // (past the end: all -1, no loads)
      const uint32_t stage = smem0 + (uint32_t)s * Cfg::STAGE;
#pragma unroll
      for (int q = 0; q < RQ; ++q) {
        const int r = q * kWtRowGroups + rg;
        const uint32_t dst = stage + (uint32_t)r * 128u + (uint32_t)((piece ^ (r & 7)) << 4);
#pragma unroll
        for (int c = 0; c < NA; ++c) {
          if (c < na) {
            // static register indices only (a runtime index would put ia[][] in local memory): cpt is 1 or 2 when multi
            const int it = !multi ? ia[q][0] : (cpt == 1 ? ia[q][c & 3] : ia[q][(c >> 1) & 3]);
            const int cc = !multi ? c : (cpt == 1 ? 0 : (c & 1));
            cp_async16_zfill(dst + (uint32_t)c * kWtChunk, gb + (size_t)max(it, 0) * g_row + cc * 128, it >= 0 ? 16u : 0u);
          }
        }
        const char* ds = db + (size_t)max(id[q], 0) * d_row;
        const uint32_t da = id[q] >= 0 ? 16u : 0u;
#pragma unroll
        for (int c = 0; c < NB; ++c)
          if (c < nb) cp_async16_zfill(dst + (uint32_t)(NA + c) * kWtChunk, ds + c * 128, da);
      }
      cp_async_commit();
      if (j >= kWtLag) {
        cp_async_wait<kWtLag>();                             // the copies of stage j - 3 have landed ...
        fence_proxy_async();                                 // ... and are visible to the tensor core (async proxy)
        mbar_arrive(smem_u32(bar_full + (j - kWtLag) % S));
      }
    };
#pragma unroll 1
    for (int j = 0; j < n_slabs; j += 3) {
      stage_step(j, ia0, id0, ia2, id2);
      if (j + 1 < n_slabs) stage_step(j + 1, ia1, id1, ia0, id0);
      if (j + 2 < n_slabs) stage_step(j + 2, ia2, id2, ia1, id1);
    }
    // drain: publish the last (up to kWtLag) stages in order
    if (n_slabs >= 3) {
      cp_async_wait<2>();
      fence_proxy_async();
      mbar_arrive(smem_u32(bar_full + (n_slabs - 3) % S));
    }
    if (n_slabs >= 2) {
      cp_async_wait<1>();
      fence_proxy_async();
      mbar_arrive(smem_u32(bar_full + (n_slabs - 2) % S));
    }
    if (n_slabs >= 1) {
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(smem_u32(bar_full + (n_slabs - 1) % S));
    }

    // ===================== epilogue (warps 0-3, one per TMEM lane quarter): quadrant sums -> partial[chunk][k][a][b] =========
    if (warp < 4) {
    if (n_slabs > 0) {
      mbar_wait(smem_u32(bar_acc), 0u);
      tc_fence_after();
    }
    float* stg = reinterpret_cast<float*>(smem);             // [2 G-chunks of the block][32 lanes][33]: lo-lane sums
    const int cpair = warp >> 1, is_lo = warp & 1;
    float* out_c = A.partial + (size_t)chunk * A.K * (size_t)A.Cg * A.Cd;
#pragma unroll 1
    for (int mb = 0; mb < Cfg::MB; ++mb) {
      const int achunk = 2 * mb + cpair;
      const bool active = achunk < na;                       // warp-uniform
#pragma unroll 1
      for (int cbk = 0; cbk < nb; ++cbk) {
        float sum[32];
        if (active && n_slabs > 0) {
          uint32_t vh[32], vl[32];
          const uint32_t t0 = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mb * 64 * NB + cbk * 64);
          tmem_ld32(t0, vh);
          tmem_ld32(t0 + 32u, vl);
#pragma unroll
          for (int c = 0; c < 32; ++c) sum[c] = __uint_as_float(vh[c]) + __uint_as_float(vl[c]);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) sum[c] = 0.f;
        }
        if (is_lo) {
#pragma unroll
          for (int c = 0; c < 32; ++c) stg[(cpair * 32 + lane) * 33 + c] = sum[c];
        }
        wt_bar_sync();
        if (!is_lo && active) {
          const int t = A.tpc > 1 ? achunk / cpt : 0, cc = A.tpc > 1 ? achunk % cpt : achunk;
          float* dst = out_c + ((size_t)(k0 + t) * A.Cg + (size_t)((a0 + cc) * 32 + lane)) * A.Cd + (size_t)(b0 + cbk) * 32;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v;
            v.x = sum[4 * q] + stg[(cpair * 32 + lane) * 33 + 4 * q];
            v.y = sum[4 * q + 1] + stg[(cpair * 32 + lane) * 33 + 4 * q + 1];
            v.z = sum[4 * q + 2] + stg[(cpair * 32 + lane) * 33 + 4 * q + 2];
            v.w = sum[4 * q + 3] + stg[(cpair * 32 + lane) * 33 + 4 * q + 3];
            reinterpret_cast<float4*>(dst)[q] = v;
          }
        }
        wt_bar_sync();
      }
    }
    }
  } else {
    // ===================== MMA issue (one elected lane, warp-uniform control flow) =====================
    // The whole warp walks the loop and only the issue is predicated: every descriptor is a warp-uniform value, so ptxas
    // keeps it in uniform registers.  (Issuing from inside `if (lane == 0)` made it wrap each UTCHMMA in an R2UR waterfall
    // loop that reuses the same uniform registers and therefore waits ~140 clk for the previous MMA to release them:
    // tools/mma_probe.cu.)
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc_bf16(kBM, 64 * nb) | (1u << 15) | (1u << 16);       // A and B MN-major
      const int mbs = (na + 1) / 2;
#pragma unroll 1
      for (int j = 0; j < n_slabs; ++j) {
        const int s = j % S;
        mbar_wait(smem_u32(bar_full + s), (uint32_t)(j / S) & 1u);
        tc_fence_after();
        const uint32_t stage = smem0 + (uint32_t)s * Cfg::STAGE;
        // all descriptors first, then the MMAs back to back in one predicated block (distinct uniform registers: an MMA
        // holds its descriptor registers for ~140 clk after issue, rewriting them would wait that long)
        uint64_t da[Cfg::MB][kWtRows / 16], dbb[kWtRows / 16];
#pragma unroll
        for (int ks = 0; ks < kWtRows / 16; ++ks) {
          dbb[ks] = wt_desc(stage + (uint32_t)NA * kWtChunk + (uint32_t)ks * 2048u, (uint32_t)kWtChunk, 1024u);
#pragma unroll
          for (int mb = 0; mb < Cfg::MB; ++mb) {
            const uint32_t lbo_a = (2 * mb + 1 < na) ? (uint32_t)kWtChunk : 0u;   // a lone chunk is mirrored into lanes 64-127
            da[mb][ks] = wt_desc(stage + (uint32_t)(2 * mb) * kWtChunk + (uint32_t)ks * 2048u, lbo_a, 1024u);
          }
        }
        if (leader) {
#pragma unroll
          for (int mb = 0; mb < Cfg::MB; ++mb) {
            if (mb < mbs) {
#pragma unroll
              for (int ks = 0; ks < kWtRows / 16; ++ks)
                wt_mma(tmem_base + (uint32_t)(mb * 64 * NB), da[mb][ks], dbb[ks], idesc, (j > 0 || ks > 0) ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(bar_empty + s));
        }
        __syncwarp();
      }
      if (n_slabs > 0 && leader) umma_commit(smem_u32(bar_acc));
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWtGatherWarps) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

__global__ void wt_reduce_kernel(const float* __restrict__ partial, int n_chunks, long long n_elem, int accumulate,
                                 float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < n_chunks; ++c) s += partial[(size_t)c * n_elem + i];      // fixed order: deterministic
    out[i] = accumulate ? out[i] + s : s;
  }
}

struct WtPlan {
  int NA, NB, n_ablk, n_bblk, chunks, rpc, tpc, ktiles;
};
static WtPlan wt_plan(int n_rows, int K, int Cg, int Cd) {
  WtPlan p;
  const int ca = Cg / 32, cb = Cd / 32;
  p.tpc = (ca <= 2 && K > 1) ? 4 / ca : 1;          // narrow G: several kernel offsets share the CTA's D tile
  p.NA = (ca >= 3 || p.tpc > 1) ? 4 : ca;
  p.NB = cb >= 3 ? 4 : cb;
  p.n_ablk = p.tpc > 1 ? 1 : div_up(ca, p.NA);
  p.n_bblk = div_up(cb, p.NB);
  p.ktiles = div_up(K, p.tpc);
  const long long tiles = (long long)p.ktiles * p.n_ablk * p.n_bblk;
  long long want = (2LL * kNumSMs) / tiles;                            // at most two full waves of CTAs (one CTA per SM)
  const long long max_chunks = (n_rows + 511) / 512;                   // at least 512 rows per chunk
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  int rpc = (int)((n_rows + want - 1) / want);
  rpc = (rpc + kWtRows - 1) / kWtRows * kWtRows;
  if (rpc < kWtRows) rpc = kWtRows;
  p.rpc = rpc;
  p.chunks = n_rows > 0 ? div_up(n_rows, rpc) : 1;
  return p;
}

template <int NA, int NB>
static int wt_launch(const WtArgs& a, const WtPlan& p, cudaStream_t st) {
  using Cfg = WtCfg<NA, NB>;
  static bool configured = false;
  if (!configured) {
    S2D_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<NA, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  const dim3 grid(p.chunks, p.ktiles, p.n_ablk * p.n_bblk);
  wgrad_tc_kernel<NA, NB><<<grid, kWtThreads, Cfg::SMEM, st>>>(a);
  S2D_LAUNCH_CHECK();
  return S2D_OK;
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_conv_wgrad_bf2_supported(int Cg, int Cd) { return Cg >= 32 && Cd >= 32 && Cg % 32 == 0 && Cd % 32 == 0; }

extern "C" size_t s2d_conv_wgrad_bf2_workspace_bytes(int n_rows, int K, int Cg, int Cd) {
  if (n_rows < 0 || K < 1 || !s2d_conv_wgrad_bf2_supported(Cg, Cd)) return 0;
  return (size_t)wt_plan(n_rows, K, Cg, Cd).chunks * K * Cg * Cd * sizeof(float);
}

extern "C" int s2d_conv_wgrad_bf2(const void* g_split, int g_ld, int n_g, int Cg, const void* d_split, int d_ld,
                                  const int* d_rows, int Cd, const int* tbl, int tbl_stride, int n_rows, int K, float* out,
                                  int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  S2D_REQUIRE(s2d_conv_wgrad_bf2_supported(Cg, Cd), "s2d_conv_wgrad_bf2: Cg = %d and Cd = %d must be multiples of 32", Cg, Cd);
  S2D_REQUIRE(g_split && d_split && tbl && out && workspace, "s2d_conv_wgrad_bf2: null pointer");
  S2D_REQUIRE(n_rows >= 0 && K >= 1 && K <= 65535 && n_g >= 0 && g_ld >= Cg && d_ld >= Cd && g_ld % 4 == 0 && d_ld % 4 == 0 &&
                  tbl_stride >= n_rows,
              "s2d_conv_wgrad_bf2: bad sizes");
  S2D_REQUIRE(((reinterpret_cast<uintptr_t>(g_split) | reinterpret_cast<uintptr_t>(d_split)) & 15) == 0,
              "s2d_conv_wgrad_bf2: split rows must be 16 B aligned");
  S2D_REQUIRE(workspace_bytes >= s2d_conv_wgrad_bf2_workspace_bytes(n_rows, K, Cg, Cd), "s2d_conv_wgrad_bf2: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const WtPlan p = wt_plan(n_rows, K, Cg, Cd);
  WtArgs a;
  a.g = static_cast<const uint32_t*>(g_split); a.d = static_cast<const uint32_t*>(d_split); a.d_rows = d_rows; a.tbl = tbl;
  a.partial = static_cast<float*>(workspace); a.g_ld = g_ld; a.d_ld = d_ld; a.tbl_stride = tbl_stride; a.n_rows = n_rows;
  a.K = K; a.Cg = Cg; a.Cd = Cd; a.ca = Cg / 32; a.cb = Cd / 32; a.rows_per_chunk = p.rpc; a.n_ablk = p.n_ablk; a.tpc = p.tpc;
  int rc = S2D_OK;
#define S2D_WT(NA_, NB_) \
  if (p.NA == NA_ && p.NB == NB_) rc = wt_launch<NA_, NB_>(a, p, st);
  S2D_WT(1, 1) S2D_WT(1, 2) S2D_WT(1, 4) S2D_WT(2, 1) S2D_WT(2, 2) S2D_WT(2, 4) S2D_WT(4, 1) S2D_WT(4, 2) S2D_WT(4, 4)
#undef S2D_WT
  if (rc != S2D_OK) return rc;
  const long long n_elem = (long long)K * Cg * Cd;
  long long blocks = (n_elem + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  wt_reduce_kernel<<<(int)blocks, 256, 0, st>>>(a.partial, p.chunks, n_elem, accumulate, out);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  return S2D_OK;
}
