// train.cu -- backward / training-mode kernels of the distillation training step.
//
// Every convolution of the path (sparse 3-D, dense 2-D, transposed, 1x1 "Linear") is one gather-GEMM over a k-major
// neighbour table  out[i] = sum_k in[tbl[k][i]] . W[k]  (spconv_tc.cu / spconv_simt.cu).  Its backward needs
//   * data gradient   dIn[j] = sum_k dOut[inv[k][j]] . W[k]^T : the SAME forward kernel over the transposed table, so the
//     only new kernel is s2d_table_transpose (inv[k][j] = i  <=>  tbl[k][i] = j; for a fixed tap k the map i -> j is
//     injective for every table of this library);
//   * weight gradient dW[k] = sum_i in[tbl[k][i]]^T . dOut[i] : s2d_conv_wgrad, a row-reduction GEMM with two-stage
//     (deterministic) accumulation.
// Training-mode BatchNorm over rows ([N, C]: the active voxels of the batch, scn.py:100-107 / the pixels of a BEV map,
// rpn.py:126-145) is split into statistics (s2d_bn_train_stats), the affine + activation (+ residual) pass
// (s2d_rows_affine_act) and the two backward passes (s2d_rows_affine_act_bwd, s2d_bn_train_bwd).  LayerNorm([C,H,W])
// and the depthwise 7x7 convolution of the ConvNeXt blocks (rpn.py:204-222) get their backward here too, and the
// optimizer is one fused launch over a flat parameter buffer (fastai-style true weight decay + Adam,
// det3d/solver/fastai_optim.py:158-174).
#include <initializer_list>

#include "common.cuh"

namespace s2d {

// --------------------------------------------------------------------------------------------------------------------
// table transpose
// --------------------------------------------------------------------------------------------------------------------
__global__ void table_transpose_kernel(const int* __restrict__ tbl, int tbl_stride, int K, int n_out,
                                       const int* __restrict__ out_rows, int* __restrict__ inv, int inv_stride, int n_in) {
  const long long total = (long long)K * n_out;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t / n_out), i = (int)(t - (long long)k * n_out);
    const int j = __ldg(tbl + (size_t)k * tbl_stride + i);
    if (j >= 0 && j < n_in) inv[(size_t)k * inv_stride + j] = out_rows ? __ldg(out_rows + i) : i;
  }
}

// --------------------------------------------------------------------------------------------------------------------
// weight gradient: out[k][a][b] = sum_i G[tbl[k][i]][a] * D[drow(i)][b]
// CTA = one tap k, one (16*TA x 16*TB) channel tile, one chunk of rows; 256 threads as 16 x 16, TA x TB outputs each.
// --------------------------------------------------------------------------------------------------------------------
constexpr int kWgSlab = 32;      // rows staged per step

template <int TA, int TB>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ G, int g_ld, int n_g, int Cg,
                                                         const float* __restrict__ D, int d_ld, const int* __restrict__ d_rows,
                                                         int Cd, const int* __restrict__ tbl, int tbl_stride, int n_rows,
                                                         int rows_per_chunk, int na_tiles, int vec4,
                                                         float* __restrict__ partial) {
  constexpr int CA = 16 * TA, CB = 16 * TB;
  // column of register v of thread tb: 8-wide tiles are split in two float4 halves 64 columns apart so that the 16 lanes of
  // a row read consecutive 16-byte words (no shared-memory bank conflicts)
  auto bcol = [](int tb, int v) { return TB == 8 ? (v < 4 ? tb * 4 + v : 64 + tb * 4 + (v - 4)) : tb * TB + v; };
  __shared__ __align__(16) float Gs[kWgSlab][CA];
  __shared__ __align__(16) float Ds[kWgSlab][CB];
  __shared__ int js[kWgSlab], is[kWgSlab];
  const int k = blockIdx.y;
  const int a_tile = blockIdx.z % na_tiles, b_tile = blockIdx.z / na_tiles;
  const int a0 = a_tile * CA, b0 = b_tile * CB;
  const int ta = threadIdx.x >> 4, tb = threadIdx.x & 15;
  float acc[TA][TB];
#pragma unroll
  for (int u = 0; u < TA; ++u)
#pragma unroll
    for (int v = 0; v < TB; ++v) acc[u][v] = 0.f;

  const int r_begin = blockIdx.x * rows_per_chunk;
  const int r_end = min(n_rows, r_begin + rows_per_chunk);
  for (int r0 = r_begin; r0 < r_end; r0 += kWgSlab) {
    if (threadIdx.x < kWgSlab) {
      const int i = r0 + threadIdx.x;
      int j = -1, di = -1;
      if (i < r_end) {
        j = __ldg(tbl + (size_t)k * tbl_stride + i);
        if (j >= n_g) j = -1;
        di = d_rows ? __ldg(d_rows + i) : i;
      }
      js[threadIdx.x] = j;
      is[threadIdx.x] = j >= 0 ? di : -1;
    }
    __syncthreads();
    if (vec4) {                                   // rows are 16-byte aligned and the channel counts multiples of 4
      for (int e = threadIdx.x; e < kWgSlab * (CA / 4); e += 256) {
        const int r = e / (CA / 4), c = (e - r * (CA / 4)) * 4;
        const int j = js[r];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j >= 0 && a0 + c < Cg) v = __ldg(reinterpret_cast<const float4*>(G + (size_t)j * g_ld + a0 + c));
        *reinterpret_cast<float4*>(&Gs[r][c]) = v;
      }
      for (int e = threadIdx.x; e < kWgSlab * (CB / 4); e += 256) {
        const int r = e / (CB / 4), c = (e - r * (CB / 4)) * 4;
        const int i = is[r];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i >= 0 && b0 + c < Cd) v = __ldg(reinterpret_cast<const float4*>(D + (size_t)i * d_ld + b0 + c));
        *reinterpret_cast<float4*>(&Ds[r][c]) = v;
      }
    } else {
      for (int e = threadIdx.x; e < kWgSlab * CA; e += 256) {
        const int r = e / CA, c = e - r * CA;
        const int j = js[r];
        Gs[r][c] = (j >= 0 && a0 + c < Cg) ? __ldg(G + (size_t)j * g_ld + a0 + c) : 0.f;
      }
      for (int e = threadIdx.x; e < kWgSlab * CB; e += 256) {
        const int r = e / CB, c = e - r * CB;
        const int i = is[r];
        Ds[r][c] = (i >= 0 && b0 + c < Cd) ? __ldg(D + (size_t)i * d_ld + b0 + c) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kWgSlab; ++r) {
      float ga[TA], db[TB];
#pragma unroll
      for (int u = 0; u < TA; ++u) ga[u] = Gs[r][ta * TA + u];
#pragma unroll
      for (int v = 0; v < TB; ++v) db[v] = Ds[r][bcol(tb, v)];
#pragma unroll
      for (int u = 0; u < TA; ++u)
#pragma unroll
        for (int v = 0; v < TB; ++v) acc[u][v] = fmaf(ga[u], db[v], acc[u][v]);
    }
    __syncthreads();
  }
  // partial[chunk][k][a][b]
  float* dst = partial + ((size_t)blockIdx.x * gridDim.y + k) * (size_t)Cg * Cd;
#pragma unroll
  for (int u = 0; u < TA; ++u) {
    const int a = a0 + ta * TA + u;
    if (a >= Cg) continue;
#pragma unroll
    for (int v = 0; v < TB; ++v) {
      const int b = b0 + bcol(tb, v);
      if (b < Cd) dst[(size_t)a * Cd + b] = acc[u][v];
    }
  }
}

// The 128 x 128 tile with the next slab in flight: rows go global -> shared with cp.async (16 B, zero-fill for missing
// neighbours) into the other half of a double buffer while the FFMA loop runs on the current one.  Needs 16-byte aligned
// rows and channel counts that are multiples of 4 (every 128 / 256 / 512-channel layer of the path).
__device__ __forceinline__ void wg_cp16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(256) conv_wgrad_pipe_kernel(const float* __restrict__ G, int g_ld, int n_g, int Cg,
                                                              const float* __restrict__ D, int d_ld,
                                                              const int* __restrict__ d_rows, int Cd,
                                                              const int* __restrict__ tbl, int tbl_stride, int n_rows,
                                                              int rows_per_chunk, int na_tiles, float* __restrict__ partial) {
  constexpr int TA = 8, TB = 8, CA = 128, CB = 128;
  extern __shared__ __align__(16) float wg_smem[];              // [2][32][128] G, then [2][32][128] D
  float(*Gs)[kWgSlab][CA] = reinterpret_cast<float(*)[kWgSlab][CA]>(wg_smem);
  float(*Ds)[kWgSlab][CB] = reinterpret_cast<float(*)[kWgSlab][CB]>(wg_smem + 2 * kWgSlab * CA);
  auto bcol = [](int t, int v) { return v < 4 ? t * 4 + v : 64 + t * 4 + (v - 4); };
  const int k = blockIdx.y;
  const int a_tile = blockIdx.z % na_tiles, b_tile = blockIdx.z / na_tiles;
  const int a0 = a_tile * CA, b0 = b_tile * CB;
  const int ta = threadIdx.x >> 4, tb = threadIdx.x & 15;
  float acc[TA][TB];
#pragma unroll
  for (int u = 0; u < TA; ++u)
#pragma unroll
    for (int v = 0; v < TB; ++v) acc[u][v] = 0.f;
  const int r_begin = blockIdx.x * rows_per_chunk;
  const int r_end = min(n_rows, r_begin + rows_per_chunk);
  const int c4 = (threadIdx.x & 31) * 4, rr = threadIdx.x >> 5;   // this thread copies column chunk c4 of rows rr, rr+8, ...
  const uint32_t gs_base = (uint32_t)__cvta_generic_to_shared(&Gs[0][0][0]);
  const uint32_t ds_base = (uint32_t)__cvta_generic_to_shared(&Ds[0][0][0]);

  auto issue = [&](int r0, int buf) {
#pragma unroll
    for (int q = 0; q < kWgSlab / 8; ++q) {
      const int r = rr + 8 * q, i = r0 + r;
      int j = -1, di = 0;
      if (i < r_end) {
        j = __ldg(tbl + (size_t)k * tbl_stride + i);
        if (j >= n_g) j = -1;
        di = d_rows ? __ldg(d_rows + i) : i;
      }
      const bool ga = j >= 0 && a0 + c4 < Cg, da = j >= 0 && b0 + c4 < Cd;
      const uint32_t off = (uint32_t)((buf * kWgSlab + r) * CA + c4) * 4u;
      wg_cp16(gs_base + off, ga ? (const void*)(G + (size_t)j * g_ld + a0 + c4) : (const void*)G, ga ? 16u : 0u);
      wg_cp16(ds_base + off, da ? (const void*)(D + (size_t)di * d_ld + b0 + c4) : (const void*)D, da ? 16u : 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (r_begin < r_end) issue(r_begin, 0);
  int buf = 0;
  for (int r0 = r_begin; r0 < r_end; r0 += kWgSlab, buf ^= 1) {
    const bool more = r0 + kWgSlab < r_end;
    if (more) {
      issue(r0 + kWgSlab, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kWgSlab; ++r) {
      float ga[TA], db[TB];
#pragma unroll
      for (int u = 0; u < TA; ++u) ga[u] = Gs[buf][r][ta * TA + u];
#pragma unroll
      for (int v = 0; v < TB; ++v) db[v] = Ds[buf][r][bcol(tb, v)];
#pragma unroll
      for (int u = 0; u < TA; ++u)
#pragma unroll
        for (int v = 0; v < TB; ++v) acc[u][v] = fmaf(ga[u], db[v], acc[u][v]);
    }
    __syncthreads();                                              // everyone is done with `buf` before it is refilled
  }
  float* dst = partial + ((size_t)blockIdx.x * gridDim.y + k) * (size_t)Cg * Cd;
#pragma unroll
  for (int u = 0; u < TA; ++u) {
    const int a = a0 + ta * TA + u;
    if (a >= Cg) continue;
#pragma unroll
    for (int v = 0; v < TB; ++v) {
      const int b = b0 + bcol(tb, v);
      if (b < Cd) dst[(size_t)a * Cd + b] = acc[u][v];
    }
  }
}

// Square channel tiles up to 64 x 64 (the 16 / 32 / 64-channel sparse stages, the 64-channel head branches): one CTA takes
// FOUR taps at once, 64 threads per tap, so that the rows of D (dOut, independent of the tap) are staged once for four
// taps and every thread still owns a T x T register tile (T = 2, 4, 8 for 16, 32, 64 channels).
constexpr int kWgTaps = 4;

template <int T>
__global__ void __launch_bounds__(256) conv_wgrad_mt_kernel(const float* __restrict__ G, int g_ld, int n_g, int Cg,
                                                            const float* __restrict__ D, int d_ld, const int* __restrict__ d_rows,
                                                            int Cd, const int* __restrict__ tbl, int tbl_stride, int n_rows, int K,
                                                            int rows_per_chunk, int vec4, float* __restrict__ partial) {
  constexpr int C = 8 * T;
  __shared__ __align__(16) float Gs[kWgTaps][kWgSlab][C];
  __shared__ __align__(16) float Ds[kWgSlab][C];
  __shared__ int js[kWgTaps][kWgSlab], is[kWgSlab];
  auto col = [](int t, int v) { return T == 8 ? (v < 4 ? t * 4 + v : 32 + t * 4 + (v - 4)) : t * T + v; };
  const int k0 = blockIdx.y * kWgTaps;
  const int kt = threadIdx.x >> 6, t = threadIdx.x & 63, ta = t >> 3, tb = t & 7;
  float acc[T][T];
#pragma unroll
  for (int u = 0; u < T; ++u)
#pragma unroll
    for (int v = 0; v < T; ++v) acc[u][v] = 0.f;
  const int r_begin = blockIdx.x * rows_per_chunk;
  const int r_end = min(n_rows, r_begin + rows_per_chunk);
  for (int r0 = r_begin; r0 < r_end; r0 += kWgSlab) {
    if (threadIdx.x < kWgSlab) {
      const int i = r0 + threadIdx.x;
      is[threadIdx.x] = i < r_end ? (d_rows ? __ldg(d_rows + i) : i) : -1;
    }
    if (threadIdx.x < kWgTaps * kWgSlab) {
      const int kk = threadIdx.x / kWgSlab, r = threadIdx.x - kk * kWgSlab, i = r0 + r;
      int j = -1;
      if (k0 + kk < K && i < r_end) {
        j = __ldg(tbl + (size_t)(k0 + kk) * tbl_stride + i);
        if (j >= n_g) j = -1;
      }
      js[kk][r] = j;
    }
    __syncthreads();
    if (vec4) {
      for (int e = threadIdx.x; e < kWgSlab * (C / 4); e += 256) {
        const int r = e / (C / 4), c = (e - r * (C / 4)) * 4;
        const int i = is[r];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i >= 0 && c < Cd) v = __ldg(reinterpret_cast<const float4*>(D + (size_t)i * d_ld + c));
        *reinterpret_cast<float4*>(&Ds[r][c]) = v;
      }
      for (int e = threadIdx.x; e < kWgTaps * kWgSlab * (C / 4); e += 256) {
        const int kk = e / (kWgSlab * (C / 4)), rem = e - kk * (kWgSlab * (C / 4));
        const int r = rem / (C / 4), c = (rem - r * (C / 4)) * 4;
        const int j = js[kk][r];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j >= 0 && c < Cg) v = __ldg(reinterpret_cast<const float4*>(G + (size_t)j * g_ld + c));
        *reinterpret_cast<float4*>(&Gs[kk][r][c]) = v;
      }
    } else {
      for (int e = threadIdx.x; e < kWgSlab * C; e += 256) {
        const int r = e / C, c = e - r * C;
        const int i = is[r];
        Ds[r][c] = (i >= 0 && c < Cd) ? __ldg(D + (size_t)i * d_ld + c) : 0.f;
      }
      for (int e = threadIdx.x; e < kWgTaps * kWgSlab * C; e += 256) {
        const int kk = e / (kWgSlab * C), rem = e - kk * (kWgSlab * C);
        const int r = rem / C, c = rem - r * C;
        const int j = js[kk][r];
        Gs[kk][r][c] = (j >= 0 && c < Cg) ? __ldg(G + (size_t)j * g_ld + c) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kWgSlab; ++r) {
      float ga[T], db[T];
#pragma unroll
      for (int u = 0; u < T; ++u) ga[u] = Gs[kt][r][col(ta, u)];
#pragma unroll
      for (int v = 0; v < T; ++v) db[v] = Ds[r][col(tb, v)];
#pragma unroll
      for (int u = 0; u < T; ++u)
#pragma unroll
        for (int v = 0; v < T; ++v) acc[u][v] = fmaf(ga[u], db[v], acc[u][v]);
    }
    __syncthreads();
  }
  const int k = k0 + kt;
  if (k >= K) return;
  float* dst = partial + ((size_t)blockIdx.x * K + k) * (size_t)Cg * Cd;
#pragma unroll
  for (int u = 0; u < T; ++u) {
    const int a = col(ta, u);
    if (a >= Cg) continue;
#pragma unroll
    for (int v = 0; v < T; ++v) {
      const int b = col(tb, v);
      if (b < Cd) dst[(size_t)a * Cd + b] = acc[u][v];
    }
  }
}

// T of the multi-tap kernel, or 0: both channel counts in the same bucket 9..16 / 17..32 / 33..64
static int wgrad_mt_tile(int Cg, int Cd) {
  auto bucket = [](int c) { return c <= 8 ? 0 : c <= 16 ? 2 : c <= 32 ? 4 : c <= 64 ? 8 : 0; };
  return bucket(Cg) == bucket(Cd) ? bucket(Cg) : 0;
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int n_chunks, long long n_elem, int accumulate,
                                    float* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem; e += (long long)gridDim.x * blockDim.x) {
    float s = accumulate ? out[e] : 0.f;
    for (int c = 0; c < n_chunks; ++c) s += partial[(size_t)c * n_elem + e];     // fixed order: deterministic
    out[e] = s;
  }
}

// Small channel counts (the 5-channel input layer, the 1/3-channel heads and generators of the PCR branch): a 16x16-thread
// channel tile would idle most lanes, so here a THREAD owns rows and keeps the whole Cg x Cd block in registers; rows of
// consecutive threads are consecutive in memory (coalesced), the block is reduced once per CTA with shuffles.
template <int CG, int CD>
__global__ void __launch_bounds__(256) conv_wgrad_small_kernel(const float* __restrict__ G, int g_ld, int n_g, int Cg,
                                                               const float* __restrict__ D, int d_ld,
                                                               const int* __restrict__ d_rows, int Cd,
                                                               const int* __restrict__ tbl, int tbl_stride, int n_rows,
                                                               int rows_per_chunk, int d_vec4, float* __restrict__ partial) {
  const int k = blockIdx.y;
  float acc[CG][CD];
#pragma unroll
  for (int a = 0; a < CG; ++a)
#pragma unroll
    for (int b = 0; b < CD; ++b) acc[a][b] = 0.f;
  const int r_begin = blockIdx.x * rows_per_chunk;
  const int r_end = min(n_rows, r_begin + rows_per_chunk);
  for (int i = r_begin + threadIdx.x; i < r_end; i += 256) {
    const int j = __ldg(tbl + (size_t)k * tbl_stride + i);
    if (j < 0 || j >= n_g) continue;
    const int di = d_rows ? __ldg(d_rows + i) : i;
    float g[CG], d[CD];
#pragma unroll
    for (int a = 0; a < CG; ++a) g[a] = a < Cg ? __ldg(G + (size_t)j * g_ld + a) : 0.f;
    if (CD % 4 == 0 && d_vec4) {                      // Cd % 4 == 0, 16-byte aligned rows
#pragma unroll
      for (int b = 0; b < CD; b += 4) {
        const float4 v = b < Cd ? __ldg(reinterpret_cast<const float4*>(D + (size_t)di * d_ld + b)) : make_float4(0.f, 0.f, 0.f, 0.f);
        d[b] = v.x; d[b + 1] = v.y; d[b + 2] = v.z; d[b + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int b = 0; b < CD; ++b) d[b] = b < Cd ? __ldg(D + (size_t)di * d_ld + b) : 0.f;
    }
#pragma unroll
    for (int a = 0; a < CG; ++a)
#pragma unroll
      for (int b = 0; b < CD; ++b) acc[a][b] = fmaf(g[a], d[b], acc[a][b]);
  }
  __shared__ float sh[8][CG * CD];
#pragma unroll
  for (int a = 0; a < CG; ++a)
#pragma unroll
    for (int b = 0; b < CD; ++b) {
      float v = acc[a][b];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][a * CD + b] = v;
    }
  __syncthreads();
  float* dst = partial + ((size_t)blockIdx.x * gridDim.y + k) * (size_t)Cg * Cd;
  for (int e = threadIdx.x; e < CG * CD; e += 256) {
    const int a = e / CD, b = e - a * CD;
    if (a < Cg && b < Cd) {
      float v = 0.f;
      for (int w = 0; w < 8; ++w) v += sh[w][e];            // fixed order
      dst[(size_t)a * Cd + b] = v;
    }
  }
}

// 0: tiled kernel; otherwise the (CG, CD) bucket of the small kernel encoded as CG * 100 + CD
static int wgrad_small_bucket(int Cg, int Cd) {
  if (Cg <= 4 && Cd <= 4) return 404;
  if (Cg <= 4 && Cd <= 16) return 416;
  if (Cg <= 8 && Cd <= 16) return 816;
  if (Cg <= 32 && Cd <= 4) return 3204;
  return 0;
}

static int wgrad_tile(int c) { return c <= 16 ? 1 : c <= 32 ? 2 : c <= 64 ? 4 : 8; }

static int wgrad_chunks(int n_rows, int K, int Cg, int Cd) {
  if (wgrad_small_bucket(Cg, Cd)) {
    long long want = (148LL * 4 + K - 1) / K;
    const long long max_chunks = (n_rows + 2047) / 2048;                         // at least 8 rows per thread
    if (want > max_chunks) want = max_chunks;
    return (int)(want < 1 ? 1 : want);
  }
  if (wgrad_mt_tile(Cg, Cd)) {
    const long long tiles = (K + kWgTaps - 1) / kWgTaps;
    long long want = (148LL * 4 + tiles - 1) / tiles;
    const long long max_chunks = (n_rows + 4 * kWgSlab - 1) / (4 * kWgSlab);
    if (want > max_chunks) want = max_chunks;
    if (want > 65535) want = 65535;
    return (int)(want < 1 ? 1 : want);
  }
  const int ta = wgrad_tile(Cg), tb = wgrad_tile(Cd);
  const long long tiles = (long long)K * ((Cg + 16 * ta - 1) / (16 * ta)) * ((Cd + 16 * tb - 1) / (16 * tb));
  long long want = (148LL * 4 + tiles - 1) / tiles;                              // ~4 CTAs per SM in total
  const long long max_chunks = (n_rows + 4 * kWgSlab - 1) / (4 * kWgSlab);       // at least 128 rows per chunk
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  return (int)want;
}

// --------------------------------------------------------------------------------------------------------------------
// column reductions over rows: per-block partial sums in double, fixed-order final pass
// --------------------------------------------------------------------------------------------------------------------
constexpr int kColBlocks = 1184;  // 8 per SM: enough loads in flight to approach the HBM rate on the 0.5 GB PCR tensors
constexpr int kColThreads = 256;

__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == S2D_ACT_RELU) return fmaxf(z, 0.f);
  if (act == S2D_ACT_GELU) return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));
  return z;
}
__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == S2D_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == S2D_ACT_GELU) {
    const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
    return cdf + z * pdf;
  }
  return 1.f;
}

// Threads of a block walk the [n, C] matrix in row-major element order with a block-stride that is a multiple of C
// where possible, so that a thread stays on one column; the general case keeps per-thread accumulators keyed by column
// through shared-memory atomics-free staging: each thread owns column (tid % C) when 256 % C == 0 or C % 256 == 0.
// To stay simple and deterministic for every C, the kernel assigns columns to threads (c = tid, tid + 256, ...) and
// rows to (blockIdx, lane group): thread t of block b handles rows r = b*RG + g, stepping by gridDim*RG, where the block
// is split into RG = max(1, 256 / Cpad) row groups of Cpad = min(C, 256) columns.
template <int NV, class F>
__device__ __forceinline__ void column_partial(int n, int C, double* __restrict__ partial, F f) {
  const int cpad = C < kColThreads ? C : kColThreads;
  const int rg = kColThreads / cpad;                      // row groups per block
  const int g = threadIdx.x / cpad, c0 = threadIdx.x - g * cpad;
  __shared__ double sh[NV][kColThreads];
  for (int cb = 0; cb < C; cb += kColThreads) {
    const int c = cb + c0;
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = 0.0;
    if (g < rg && c < C) {
      float facc[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) facc[v] = 0.f;
      int since = 0;
      for (int r = blockIdx.x * rg + g; r < n; r += gridDim.x * rg) {
        f(r, c, facc);
        if (++since == 64) {                                // flush the float accumulators: bounded rounding error
#pragma unroll
          for (int v = 0; v < NV; ++v) { acc[v] += (double)facc[v]; facc[v] = 0.f; }
          since = 0;
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] += (double)facc[v];
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) sh[v][threadIdx.x] = acc[v];
    __syncthreads();
    if (g == 0 && c < C) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        double s = 0.0;
        for (int q = 0; q < rg; ++q) s += sh[v][q * cpad + c0];
        partial[((size_t)blockIdx.x * NV + v) * C + c] = s;
      }
    }
    __syncthreads();
  }
}

// sums[v][c] = sum over blocks: one warp per output, lane l adds blocks l, l+32, ... in order, then a fixed shuffle tree
// (deterministic for a given block count)
__global__ void column_final_kernel(const double* __restrict__ partial, int nblocks, int nv_c, double* __restrict__ sums) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= nv_c) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partial[(size_t)b * nv_c + e];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) sums[e] = s;
}

__global__ void __launch_bounds__(kColThreads) bn_stats_partial_kernel(const float* __restrict__ x, int ld, int n, int C,
                                                                      double* __restrict__ partial) {
  column_partial<2>(n, C, partial, [&](int r, int c, float (&a)[2]) {
    const float v = __ldg(x + (size_t)r * ld + c);
    a[0] += v;
    a[1] = fmaf(v, v, a[1]);
  });
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double n, const double* __restrict__ n_dev, int C, float eps, float momentum,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ scale,
                                   float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (n_dev) n = *n_dev;
  const double m = sums[c] / n;
  double var = sums[C + c] / n - m * m;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)m;
  invstd[c] = is;
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * is;
  shift[c] = b - (float)m * g * is;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
  if (running_var) {
    const double unbiased = n > 1.0 ? var * (n / (n - 1.0)) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void rows_affine_act_kernel(const float* __restrict__ x, int ld, long long n, int C,
                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                       const float* __restrict__ residual, int res_ld, int act, int res_after_act,
                                       float* __restrict__ out, int out_ld) {
  const long long total = n * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int c = (int)(e - r * C);
    float z = x[(size_t)r * ld + c];
    if (scale) z *= __ldg(scale + c);
    if (shift) z += __ldg(shift + c);
    const float res = residual ? residual[(size_t)r * res_ld + c] : 0.f;
    float y;
    if (res_after_act) y = act_fwd(z, act) + res;
    else y = act_fwd(z + res, act);
    out[(size_t)r * out_ld + c] = y;
  }
}

// the same, four channels per thread (C, the row strides and the base addresses multiples of 4 floats)
__global__ void rows_affine_act_vec4_kernel(const float* __restrict__ x, int ld, long long n, int C,
                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                            const float* __restrict__ residual, int res_ld, int act, int res_after_act,
                                            float* __restrict__ out, int out_ld) {
  const int c4n = C >> 2;
  const long long total = n * c4n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / c4n;
    const int c = (int)(e - r * c4n) << 2;
    const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)r * ld + c);
    const float4 sc = scale ? __ldg(reinterpret_cast<const float4*>(scale + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 sh = shift ? __ldg(reinterpret_cast<const float4*>(shift + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 rv = residual ? *reinterpret_cast<const float4*>(residual + (size_t)r * res_ld + c)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    const float zx[4] = {xv.x, xv.y, xv.z, xv.w}, s4[4] = {sc.x, sc.y, sc.z, sc.w}, h4[4] = {sh.x, sh.y, sh.z, sh.w},
                r4[4] = {rv.x, rv.y, rv.z, rv.w};
    float y[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float z = zx[q];
      if (scale) z *= s4[q];
      if (shift) z += h4[q];
      y[q] = res_after_act ? act_fwd(z, act) + r4[q] : act_fwd(z + r4[q], act);
    }
    *reinterpret_cast<float4*>(out + (size_t)r * out_ld + c) = make_float4(y[0], y[1], y[2], y[3]);
  }
}

// dz = dy * act'(pre-activation); partial sums of dz and dz * x per column
__global__ void __launch_bounds__(kColThreads) rows_affine_act_bwd_kernel(
    const float* __restrict__ x, int ld, int n, int C, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ residual, int res_ld, int act, int res_after_act, const float* __restrict__ dy, int dy_ld,
    float* __restrict__ dz, int dz_ld, double* __restrict__ partial) {
  column_partial<2>(n, C, partial, [&](int r, int c, float (&a)[2]) {
    const float xv = x[(size_t)r * ld + c];
    float z = xv;
    if (scale) z *= __ldg(scale + c);
    if (shift) z += __ldg(shift + c);
    if (residual && !res_after_act) z += residual[(size_t)r * res_ld + c];
    const float g = dy[(size_t)r * dy_ld + c] * act_grad(z, act);
    dz[(size_t)r * dz_ld + c] = g;
    a[0] += g;
    a[1] = fmaf(g, xv, a[1]);
  });
}

// dx = gamma*invstd * (dz - mean(dz) - xhat * mean(dz*xhat)); also dgamma / dbeta
__global__ void bn_bwd_coeff_kernel(const double* __restrict__ sums, double n, const double* __restrict__ n_dev, int C, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (n_dev) n = *n_dev;
  const double sdz = sums[c], sdzx = sums[C + c];
  const double m = mean[c], is = invstd[c];
  const double dg = is * (sdzx - m * sdz);               // sum dz * xhat
  if (dgamma) dgamma[c] = (float)dg;
  if (dbeta) dbeta[c] = (float)sdz;
  if (!coef) return;
  const double g = gamma ? (double)gamma[c] : 1.0;
  // dx = a*dz + b*x + d  with  a = g*is,  b = -g*is*is*dg/n,  d = -g*is*(sdz/n) + g*is*is*m*dg/n
  coef[c] = (float)(g * is);
  coef[C + c] = (float)(-g * is * is * dg / n);
  coef[2 * C + c] = (float)(-g * is * sdz / n + g * is * is * m * dg / n);
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, int ld, long long n, int C, const float* __restrict__ dz,
                                    int dz_ld, const float* __restrict__ coef, float* __restrict__ dx, int dx_ld) {
  const long long total = n * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int c = (int)(e - r * C);
    dx[(size_t)r * dx_ld + c] = fmaf(__ldg(coef + c), dz[(size_t)r * dz_ld + c],
                                     fmaf(__ldg(coef + C + c), x[(size_t)r * ld + c], __ldg(coef + 2 * C + c)));
  }
}

__global__ void bn_bwd_apply_vec4_kernel(const float* __restrict__ x, int ld, long long n, int C, const float* __restrict__ dz,
                                         int dz_ld, const float* __restrict__ coef, float* __restrict__ dx, int dx_ld) {
  const int c4n = C >> 2;
  const long long total = n * c4n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / c4n;
    const int c = (int)(e - r * c4n) << 2;
    const float4 a = __ldg(reinterpret_cast<const float4*>(coef + c)), b = __ldg(reinterpret_cast<const float4*>(coef + C + c)),
                 d = __ldg(reinterpret_cast<const float4*>(coef + 2 * C + c));
    const float4 g = *reinterpret_cast<const float4*>(dz + (size_t)r * dz_ld + c);
    const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)r * ld + c);
    *reinterpret_cast<float4*>(dx + (size_t)r * dx_ld + c) =
        make_float4(fmaf(a.x, g.x, fmaf(b.x, xv.x, d.x)), fmaf(a.y, g.y, fmaf(b.y, xv.y, d.y)),
                    fmaf(a.z, g.z, fmaf(b.z, xv.z, d.z)), fmaf(a.w, g.w, fmaf(b.w, xv.w, d.w)));
  }
}

static bool vec4_ok(int C, std::initializer_list<int> lds, std::initializer_list<const void*> ptrs) {
  if (C % 4) return false;
  for (int l : lds) if (l % 4) return false;
  for (const void* p : ptrs) if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return false;
  return true;
}

__global__ void sums_to_float_kernel(const double* __restrict__ sums, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (float)sums[c];
}

// --------------------------------------------------------------------------------------------------------------------
// LayerNorm([C, H, W]) backward on rows x[b*HW + p][c], weight / bias indexed [c][p]
// --------------------------------------------------------------------------------------------------------------------
constexpr int kLnBlocks = 64;      // per sample

// partial[b][blk][4] = {sum x, sum x^2, -, -}   or   {sum g, sum g*xhat}
__global__ void __launch_bounds__(256) ln_stats_kernel(const float* __restrict__ x, int C, int HW, double* __restrict__ partial) {
  const int b = blockIdx.y;
  const long long per = (long long)C * HW;
  const float* xb = x + (size_t)b * per;
  double s = 0.0, q = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += (long long)gridDim.x * blockDim.x) {
    const double v = xb[e];
    s += v;
    q += v * v;
  }
  __shared__ double sh[2][8];
  for (int d = 16; d > 0; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); q += __shfl_xor_sync(0xffffffffu, q, d); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < 8; ++w) { a += sh[0][w]; c += sh[1][w]; }
    partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + 0] = a;
    partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = c;
  }
}

__global__ void ln_finalize_kernel(const double* __restrict__ partial, int nblk, long long per, float eps,
                                   float* __restrict__ stats /*[B][2] mean, invstd*/) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < nblk; ++k) { s += partial[((size_t)b * nblk + k) * 2]; q += partial[((size_t)b * nblk + k) * 2 + 1]; }
  const double m = s / per;
  double var = q / per - m * m;
  if (var < 0.0) var = 0.0;
  stats[b * 2] = (float)m;
  stats[b * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void __launch_bounds__(256) ln_bwd_sums_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          const float* __restrict__ w, int C, int HW,
                                                          const float* __restrict__ stats, double* __restrict__ partial) {
  const int b = blockIdx.y;
  const long long per = (long long)C * HW;
  const float* xb = x + (size_t)b * per;
  const float* db = dy + (size_t)b * per;
  const float m = stats[b * 2], is = stats[b * 2 + 1];
  double s = 0.0, q = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(e / C), c = (int)(e - (long long)p * C);
    const float g = db[e] * (w ? __ldg(w + (size_t)c * HW + p) : 1.f);
    const float xh = (xb[e] - m) * is;
    s += (double)g;
    q += (double)g * (double)xh;
  }
  __shared__ double sh[2][8];
  for (int d = 16; d > 0; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); q += __shfl_xor_sync(0xffffffffu, q, d); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int k = 0; k < 8; ++k) { a += sh[0][k]; c += sh[1][k]; }
    partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + 0] = a;
    partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = c;
  }
}

__global__ void ln_bwd_finalize_kernel(const double* __restrict__ partial, int nblk, long long per, float* __restrict__ gs) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < nblk; ++k) { s += partial[((size_t)b * nblk + k) * 2]; q += partial[((size_t)b * nblk + k) * 2 + 1]; }
  gs[b * 2] = (float)(s / per);
  gs[b * 2 + 1] = (float)(q / per);
}

// dx, and dw[c][p] = sum_b dy*xhat, db[c][p] = sum_b dy  (one thread per (p, c), loop over the batch)
__global__ void ln_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ w,
                                    int B, int C, int HW, const float* __restrict__ stats, const float* __restrict__ gs,
                                    float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ dbias) {
  const long long per = (long long)C * HW;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(e / C), c = (int)(e - (long long)p * C);
    const float wv = w ? __ldg(w + (size_t)c * HW + p) : 1.f;
    float aw = 0.f, ab = 0.f;
    for (int b = 0; b < B; ++b) {
      const float m = stats[b * 2], is = stats[b * 2 + 1];
      const float d = dy[(size_t)b * per + e];
      const float xh = (x[(size_t)b * per + e] - m) * is;
      aw = fmaf(d, xh, aw);
      ab += d;
      dx[(size_t)b * per + e] = is * (d * wv - gs[b * 2] - xh * gs[b * 2 + 1]);
    }
    if (dw) dw[(size_t)c * HW + p] = aw;
    if (dbias) dbias[(size_t)c * HW + p] = ab;
  }
}

// --------------------------------------------------------------------------------------------------------------------
// depthwise k x k convolution, weight gradient: dw[c][ky][kx] = sum_{b,y,x} dy[b,y,x,c] * x[b,y+ky-p,x+kx-p,c]
// grid (chunks, k*k), threads over channels; partial[chunk][tap][c] then a fixed-order reduction
// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B,
                                                           int H, int W, int C, int k, int pad, int rows_per_chunk,
                                                           float* __restrict__ partial) {
  const int tap = blockIdx.y, ky = tap / k, kx = tap - ky * k;
  const int n = B * H * W;
  const int r0 = blockIdx.x * rows_per_chunk, r1 = min(n, r0 + rows_per_chunk);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) {
      const int b = r / (H * W), rem = r - b * H * W, y = rem / W, xx = rem - y * W;
      const int sy = y + ky - pad, sx = xx + kx - pad;
      if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
      acc = fmaf(dy[(size_t)r * C + c], x[((size_t)(b * H + sy) * W + sx) * C + c], acc);
    }
    partial[((size_t)blockIdx.x * gridDim.y + tap) * C + c] = acc;
  }
}

// out[c][tap] (torch depthwise weight layout [C,1,k,k]) = sum over chunks
__global__ void dwconv_wgrad_reduce_kernel(const float* __restrict__ partial, int n_chunks, int taps, int C,
                                           float* __restrict__ dw) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= taps * C) return;
  const int tap = e / C, c = e - tap * C;
  float s = 0.f;
  for (int q = 0; q < n_chunks; ++q) s += partial[((size_t)q * taps + tap) * C + c];
  dw[(size_t)c * taps + tap] = s;
}

// --------------------------------------------------------------------------------------------------------------------
// optimizer: p *= 1 - wd*lr (true weight decay), then Adam (torch.optim.Adam arithmetic, amsgrad off)
// --------------------------------------------------------------------------------------------------------------------
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                                 float decay, float bc1, float bc2_sqrt, const float* __restrict__ grad_scale) {
  const float gs = grad_scale ? *grad_scale : 1.f;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float gr = g[e] * gs;
    const float mm = beta1 * m[e] + (1.f - beta1) * gr;
    const float vv = beta2 * v[e] + (1.f - beta2) * gr * gr;
    m[e] = mm;
    v[e] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[e] = p[e] * decay - (lr / bc1) * (mm / denom);
  }
}

// sum of squares of a flat buffer -> partial doubles; final: clip coefficient min(1, max_norm / (norm + 1e-6))
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n, double* __restrict__ partial) {
  double s = 0.0;
  float f = 0.f;
  int since = 0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float v = g[e];
    f = fmaf(v, v, f);
    if (++since == 32) { s += (double)f; f = 0.f; since = 0; }
  }
  s += (double)f;
  __shared__ double sh[8];
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < 8; ++w) a += sh[w];
    partial[blockIdx.x] = a;
  }
}

__global__ void clip_coef_kernel(const double* __restrict__ partial, int nblk, float max_norm, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partial[b];
  const float norm = (float)sqrt(s);
  out[0] = norm;
  const float coef = max_norm / (norm + 1e-6f);
  out[1] = (max_norm > 0.f && coef < 1.f) ? coef : 1.f;
}

static int grid_for(long long n, int threads, int cap) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_table_transpose(const int* tbl, int tbl_stride, int K, int n_out, const int* out_rows, int* inv,
                                   int inv_stride, int n_in, void* stream) {
  S2D_REQUIRE(K >= 1 && n_out >= 0 && n_in >= 0 && inv_stride >= n_in && tbl_stride >= n_out,
              "s2d_table_transpose: bad sizes (K %d, n_out %d, n_in %d, strides %d / %d)", K, n_out, n_in, tbl_stride, inv_stride);
  S2D_REQUIRE((n_out == 0 || tbl) && (n_in == 0 || inv), "s2d_table_transpose: null table");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_in > 0)
    S2D_CUDA(cudaMemset2DAsync(inv, (size_t)inv_stride * sizeof(int), 0xFF, (size_t)n_in * sizeof(int), (size_t)K, st));
  if (n_out > 0 && n_in > 0) {
    table_transpose_kernel<<<grid_for((long long)K * n_out, 256, 148 * 16), 256, 0, st>>>(tbl, tbl_stride, K, n_out, out_rows,
                                                                                      inv, inv_stride, n_in);
    S2D_LAUNCH_CHECK();
    count_launches(1);
  }
  return S2D_OK;
}

extern "C" size_t s2d_conv_wgrad_workspace_bytes(int n_rows, int K, int Cg, int Cd) {
  if (n_rows < 0 || K < 1 || Cg < 1 || Cd < 1) return 0;
  return (size_t)wgrad_chunks(n_rows, K, Cg, Cd) * K * Cg * Cd * sizeof(float);
}

template <int TA, int TB>
static void launch_wgrad(dim3 grid, cudaStream_t st, const float* g, int g_ld, int n_g, int Cg, const float* d, int d_ld,
                         const int* d_rows, int Cd, const int* tbl, int tbl_stride, int n_rows, int rpc, int na, float* partial) {
  const int vec4 = (g_ld % 4 == 0) && (d_ld % 4 == 0) && (Cg % 4 == 0) && (Cd % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
  conv_wgrad_kernel<TA, TB><<<grid, 256, 0, st>>>(g, g_ld, n_g, Cg, d, d_ld, d_rows, Cd, tbl, tbl_stride, n_rows, rpc, na, vec4,
                                                  partial);
}

extern "C" int s2d_conv_wgrad(const float* g, int g_ld, int n_g, int Cg, const float* d, int d_ld, const int* d_rows, int Cd,
                              const int* tbl, int tbl_stride, int n_rows, int K, float* out, int accumulate, void* ws,
                              size_t ws_bytes, void* stream) {
  S2D_REQUIRE(g && d && tbl && out && ws, "s2d_conv_wgrad: null pointer");
  S2D_REQUIRE(n_rows >= 0 && K >= 1 && K <= 65535 && Cg >= 1 && Cd >= 1 && g_ld >= Cg && d_ld >= Cd && tbl_stride >= n_rows,
              "s2d_conv_wgrad: bad sizes");
  S2D_REQUIRE(ws_bytes >= s2d_conv_wgrad_workspace_bytes(n_rows, K, Cg, Cd), "s2d_conv_wgrad: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ta = wgrad_tile(Cg), tb = wgrad_tile(Cd);
  const int na = (Cg + 16 * ta - 1) / (16 * ta), nb = (Cd + 16 * tb - 1) / (16 * tb);
  const int chunks = wgrad_chunks(n_rows, K, Cg, Cd);
  int rpc = (n_rows + chunks - 1) / chunks;
  rpc = ((rpc + kWgSlab - 1) / kWgSlab) * kWgSlab;
  if (rpc < kWgSlab) rpc = kWgSlab;
  dim3 grid(chunks, K, na * nb);
  float* partial = static_cast<float*>(ws);
  const int bucket = wgrad_small_bucket(Cg, Cd);
  const int mt = bucket ? 0 : wgrad_mt_tile(Cg, Cd);
  if (mt) {
    const int vec4 = (g_ld % 4 == 0) && (d_ld % 4 == 0) && (Cg % 4 == 0) && (Cd % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
    dim3 gm(chunks, (K + kWgTaps - 1) / kWgTaps);
    if (mt == 2) conv_wgrad_mt_kernel<2><<<gm, 256, 0, st>>>(g, g_ld, n_g, Cg, d, d_ld, d_rows, Cd, tbl, tbl_stride, n_rows, K, rpc, vec4, partial);
    if (mt == 4) conv_wgrad_mt_kernel<4><<<gm, 256, 0, st>>>(g, g_ld, n_g, Cg, d, d_ld, d_rows, Cd, tbl, tbl_stride, n_rows, K, rpc, vec4, partial);
    if (mt == 8) conv_wgrad_mt_kernel<8><<<gm, 256, 0, st>>>(g, g_ld, n_g, Cg, d, d_ld, d_rows, Cd, tbl, tbl_stride, n_rows, K, rpc, vec4, partial);
    S2D_LAUNCH_CHECK();
    const long long n_elem_m = (long long)K * Cg * Cd;
    wgrad_reduce_kernel<<<grid_for(n_elem_m, 256, 148 * 8), 256, 0, st>>>(partial, chunks, n_elem_m, accumulate, out);
    S2D_LAUNCH_CHECK();
    count_launches(2);
    return S2D_OK;
  }
  if (bucket) {
    dim3 gs(chunks, K);
    const int d_vec4 = (d_ld % 4 == 0) && (Cd % 4 == 0) && (reinterpret_cast<uintptr_t>(d) & 15) == 0;
#define S2D_WGS(CG_, CD_)                                                                                              \
  if (bucket == CG_ * 100 + CD_)                                                                                        \
    conv_wgrad_small_kernel<CG_, CD_><<<gs, 256, 0, st>>>(g, g_ld, n_g, Cg, d, d_ld, d_rows, Cd, tbl, tbl_stride, n_rows, \
                                                          rpc, d_vec4, partial);
    S2D_WGS(4, 4) S2D_WGS(4, 16) S2D_WGS(8, 16) S2D_WGS(32, 4)
#undef S2D_WGS
    S2D_LAUNCH_CHECK();
    const long long n_elem_s = (long long)K * Cg * Cd;
    wgrad_reduce_kernel<<<grid_for(n_elem_s, 256, 148 * 8), 256, 0, st>>>(partial, chunks, n_elem_s, accumulate, out);
    S2D_LAUNCH_CHECK();
    count_launches(2);
    return S2D_OK;
  }
  if (ta == 8 && tb == 8 && (g_ld % 4 == 0) && (d_ld % 4 == 0) && (Cg % 4 == 0) && (Cd % 4 == 0) &&
      ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
    constexpr int smem = 4 * kWgSlab * 128 * (int)sizeof(float);                // 64 KB: two slabs of G and of D
    static bool configured = false;
    if (!configured) {
      S2D_CUDA(cudaFuncSetAttribute(conv_wgrad_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured = true;
    }
    conv_wgrad_pipe_kernel<<<grid, 256, smem, st>>>(g, g_ld, n_g, Cg, d, d_ld, d_rows, Cd, tbl, tbl_stride, n_rows, rpc, na,
                                                    partial);
    S2D_LAUNCH_CHECK();
    const long long n_elem_p = (long long)K * Cg * Cd;
    wgrad_reduce_kernel<<<grid_for(n_elem_p, 256, 148 * 8), 256, 0, st>>>(partial, chunks, n_elem_p, accumulate, out);
    S2D_LAUNCH_CHECK();
    count_launches(2);
    return S2D_OK;
  }
#define S2D_WG(TA_, TB_) \
  if (ta == TA_ && tb == TB_) launch_wgrad<TA_, TB_>(grid, st, g, g_ld, n_g, Cg, d, d_ld, d_rows, Cd, tbl, tbl_stride, n_rows, rpc, na, partial);
  S2D_WG(1, 1) S2D_WG(1, 2) S2D_WG(1, 4) S2D_WG(1, 8)
  S2D_WG(2, 1) S2D_WG(2, 2) S2D_WG(2, 4) S2D_WG(2, 8)
  S2D_WG(4, 1) S2D_WG(4, 2) S2D_WG(4, 4) S2D_WG(4, 8)
  S2D_WG(8, 1) S2D_WG(8, 2) S2D_WG(8, 4) S2D_WG(8, 8)
#undef S2D_WG
  S2D_LAUNCH_CHECK();
  const long long n_elem = (long long)K * Cg * Cd;
  wgrad_reduce_kernel<<<grid_for(n_elem, 256, 148 * 8), 256, 0, st>>>(partial, chunks, n_elem, accumulate, out);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  return S2D_OK;
}

extern "C" size_t s2d_rows_workspace_bytes(int C) {
  if (C < 1) return 0;
  // partial sums [kColBlocks][2][C] + final sums [2][C] (double) + 3*C float coefficients
  return ((size_t)kColBlocks * 2 * C + 2 * (size_t)C) * sizeof(double) + 3 * (size_t)C * sizeof(float);
}

extern "C" int s2d_bn_train_stats(const float* x, int ld, int n, int C, float eps, float momentum, const float* gamma,
                                  const float* beta, float* running_mean, float* running_var, float* mean, float* invstd,
                                  float* scale, float* shift, void* ws, size_t ws_bytes, void* stream) {
  S2D_REQUIRE(x && mean && invstd && scale && shift && ws, "s2d_bn_train_stats: null pointer");
  S2D_REQUIRE(n >= 1 && C >= 1 && ld >= C, "s2d_bn_train_stats: bad sizes (n %d, C %d, ld %d)", n, C, ld);
  S2D_REQUIRE(ws_bytes >= s2d_rows_workspace_bytes(C), "s2d_bn_train_stats: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(ws);
  double* sums = partial + (size_t)kColBlocks * 2 * C;
  bn_stats_partial_kernel<<<kColBlocks, kColThreads, 0, st>>>(x, ld, n, C, partial);
  S2D_LAUNCH_CHECK();
  column_final_kernel<<<(2 * C * 32 + 255) / 256, 256, 0, st>>>(partial, kColBlocks, 2 * C, sums);
  S2D_LAUNCH_CHECK();
  bn_finalize_kernel<<<(C + 255) / 256, 256, 0, st>>>(sums, (double)n, nullptr, C, eps, momentum, gamma, beta, running_mean,
                                                      running_var, mean, invstd, scale, shift);
  S2D_LAUNCH_CHECK();
  count_launches(3);
  return S2D_OK;
}

// ---- the same in two halves, for SyncBatchNorm: the caller all-reduces the column sums (2*C doubles at
// s2d_rows_workspace_sums_offset(C) inside the workspace) and the row counts between the two calls ----------------------
extern "C" size_t s2d_rows_workspace_sums_offset(int C) { return C < 1 ? 0 : (size_t)kColBlocks * 2 * C * sizeof(double); }

extern "C" int s2d_bn_train_sums(const float* x, int ld, int n, int C, void* ws, size_t ws_bytes, void* stream) {
  S2D_REQUIRE(ws && (n == 0 || x), "s2d_bn_train_sums: null pointer");
  S2D_REQUIRE(n >= 0 && C >= 1 && ld >= C, "s2d_bn_train_sums: bad sizes (n %d, C %d, ld %d)", n, C, ld);
  S2D_REQUIRE(ws_bytes >= s2d_rows_workspace_bytes(C), "s2d_bn_train_sums: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(ws);
  double* sums = partial + (size_t)kColBlocks * 2 * C;
  bn_stats_partial_kernel<<<kColBlocks, kColThreads, 0, st>>>(x, ld, n, C, partial);
  S2D_LAUNCH_CHECK();
  column_final_kernel<<<(2 * C * 32 + 255) / 256, 256, 0, st>>>(partial, kColBlocks, 2 * C, sums);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  return S2D_OK;
}

extern "C" int s2d_bn_train_finalize(const double* sums2c, const double* n_total_dev, int C, float eps, float momentum,
                                     const float* gamma, const float* beta, float* running_mean, float* running_var,
                                     float* mean, float* invstd, float* scale, float* shift, void* stream) {
  S2D_REQUIRE(sums2c && n_total_dev && mean && invstd && scale && shift && C >= 1, "s2d_bn_train_finalize: bad argument");
  bn_finalize_kernel<<<(C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      sums2c, 1.0, n_total_dev, C, eps, momentum, gamma, beta, running_mean, running_var, mean, invstd, scale, shift);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_rows_affine_act(const float* x, int ld, int n, int C, const float* scale, const float* shift,
                                   const float* residual, int res_ld, int act, int res_after_act, float* out, int out_ld,
                                   void* stream) {
  S2D_REQUIRE(n >= 0 && C >= 1 && ld >= C && out_ld >= C && (!residual || res_ld >= C), "s2d_rows_affine_act: bad sizes");
  S2D_REQUIRE(act >= S2D_ACT_NONE && act <= S2D_ACT_GELU, "s2d_rows_affine_act: unknown activation %d", act);
  if (n == 0) return S2D_OK;
  S2D_REQUIRE(x && out, "s2d_rows_affine_act: null pointer");
  if (vec4_ok(C, {ld, out_ld, residual ? res_ld : 0}, {x, out, residual, scale, shift}))
    rows_affine_act_vec4_kernel<<<grid_for((long long)n * (C / 4), 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ld, n, C, scale, shift, residual, res_ld, act, res_after_act, out, out_ld);
  else
    rows_affine_act_kernel<<<grid_for((long long)n * C, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ld, n, C, scale, shift, residual, res_ld, act, res_after_act, out, out_ld);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_rows_affine_act_bwd(const float* x, int ld, int n, int C, const float* scale, const float* shift,
                                       const float* residual, int res_ld, int act, int res_after_act, const float* dy,
                                       int dy_ld, float* dz, int dz_ld, float* dshift, void* ws, size_t ws_bytes, void* stream) {
  S2D_REQUIRE(x && dy && dz && ws, "s2d_rows_affine_act_bwd: null pointer");
  S2D_REQUIRE(n >= 1 && C >= 1 && ld >= C && dy_ld >= C && dz_ld >= C && (!residual || res_ld >= C),
              "s2d_rows_affine_act_bwd: bad sizes");
  S2D_REQUIRE(ws_bytes >= s2d_rows_workspace_bytes(C), "s2d_rows_affine_act_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(ws);
  double* sums = partial + (size_t)kColBlocks * 2 * C;
  rows_affine_act_bwd_kernel<<<kColBlocks, kColThreads, 0, st>>>(x, ld, n, C, scale, shift, residual, res_ld, act,
                                                                res_after_act, dy, dy_ld, dz, dz_ld, partial);
  S2D_LAUNCH_CHECK();
  column_final_kernel<<<(2 * C * 32 + 255) / 256, 256, 0, st>>>(partial, kColBlocks, 2 * C, sums);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  if (dshift) {                       // plain bias: d(bias) = column sums of dz
    sums_to_float_kernel<<<(C + 255) / 256, 256, 0, st>>>(sums, C, dshift);
    S2D_LAUNCH_CHECK();
    count_launches(1);
  }
  return S2D_OK;
}

// ws must be the workspace the preceding s2d_rows_affine_act_bwd call filled (it holds the column sums)
extern "C" int s2d_bn_train_bwd(const float* x, int ld, int n, int C, const float* dz, int dz_ld, const float* mean,
                                const float* invstd, const float* gamma, float* dx, int dx_ld, float* dgamma, float* dbeta,
                                void* ws, size_t ws_bytes, void* stream) {
  S2D_REQUIRE(x && dz && mean && invstd && dx && ws, "s2d_bn_train_bwd: null pointer");
  S2D_REQUIRE(n >= 1 && C >= 1 && ld >= C && dz_ld >= C && dx_ld >= C, "s2d_bn_train_bwd: bad sizes");
  S2D_REQUIRE(ws_bytes >= s2d_rows_workspace_bytes(C), "s2d_bn_train_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* sums = static_cast<double*>(ws) + (size_t)kColBlocks * 2 * C;
  float* coef = reinterpret_cast<float*>(sums + 2 * (size_t)C);
  bn_bwd_coeff_kernel<<<(C + 255) / 256, 256, 0, st>>>(sums, (double)n, nullptr, C, mean, invstd, gamma, dgamma, dbeta, coef);
  S2D_LAUNCH_CHECK();
  if (vec4_ok(C, {ld, dz_ld, dx_ld}, {x, dz, dx, coef}))
    bn_bwd_apply_vec4_kernel<<<grid_for((long long)n * (C / 4), 256, 148 * 16), 256, 0, st>>>(x, ld, n, C, dz, dz_ld, coef, dx,
                                                                                           dx_ld);
  else
    bn_bwd_apply_kernel<<<grid_for((long long)n * C, 256, 148 * 16), 256, 0, st>>>(x, ld, n, C, dz, dz_ld, coef, dx, dx_ld);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  return S2D_OK;
}

// SyncBatchNorm backward in two halves around the all-reduce of the column sums in the workspace: the affine gradients
// come from the LOCAL sums (data-parallel gradient averaging treats them like any other parameter gradient), dx from the
// GLOBAL sums and row count.
extern "C" int s2d_bn_train_bwd_params(const double* sums2c_local, int C, const float* mean, const float* invstd,
                                       float* dgamma, float* dbeta, void* stream) {
  S2D_REQUIRE(sums2c_local && mean && invstd && C >= 1, "s2d_bn_train_bwd_params: bad argument");
  bn_bwd_coeff_kernel<<<(C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(sums2c_local, 1.0, nullptr, C, mean,
                                                                                     invstd, nullptr, dgamma, dbeta, nullptr);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_bn_train_bwd_dx(const float* x, int ld, int n, int C, const float* dz, int dz_ld, const float* mean,
                                   const float* invstd, const float* gamma, const double* sums2c_global,
                                   const double* n_total_dev, float* dx, int dx_ld, void* ws, size_t ws_bytes, void* stream) {
  S2D_REQUIRE(ws && mean && invstd && sums2c_global && n_total_dev && (n == 0 || (x && dz && dx)),
              "s2d_bn_train_bwd_dx: null pointer");
  S2D_REQUIRE(n >= 0 && C >= 1 && ld >= C && dz_ld >= C && dx_ld >= C, "s2d_bn_train_bwd_dx: bad sizes");
  S2D_REQUIRE(ws_bytes >= s2d_rows_workspace_bytes(C), "s2d_bn_train_bwd_dx: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* coef = reinterpret_cast<float*>(static_cast<double*>(ws) + (size_t)kColBlocks * 2 * C + 2 * (size_t)C);
  bn_bwd_coeff_kernel<<<(C + 255) / 256, 256, 0, st>>>(sums2c_global, 1.0, n_total_dev, C, mean, invstd, gamma, nullptr, nullptr,
                                                      coef);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  if (n > 0) {
    if (vec4_ok(C, {ld, dz_ld, dx_ld}, {x, dz, dx, coef}))
      bn_bwd_apply_vec4_kernel<<<grid_for((long long)n * (C / 4), 256, 148 * 16), 256, 0, st>>>(x, ld, n, C, dz, dz_ld, coef,
                                                                                             dx, dx_ld);
    else
      bn_bwd_apply_kernel<<<grid_for((long long)n * C, 256, 148 * 16), 256, 0, st>>>(x, ld, n, C, dz, dz_ld, coef, dx, dx_ld);
    S2D_LAUNCH_CHECK();
    count_launches(1);
  }
  return S2D_OK;
}

extern "C" size_t s2d_layernorm_bwd_workspace_bytes(int B) {
  if (B < 1) return 0;
  return (size_t)B * kLnBlocks * 2 * sizeof(double) + (size_t)B * 4 * sizeof(float);
}

extern "C" int s2d_layernorm_chw_bwd(const float* x, const float* weight, int B, int C, int HW, float eps, const float* dy,
                                     float* dx, float* dweight, float* dbias, void* ws, size_t ws_bytes, void* stream) {
  S2D_REQUIRE(x && dy && dx && ws, "s2d_layernorm_chw_bwd: null pointer");
  S2D_REQUIRE(B >= 1 && C >= 1 && HW >= 1, "s2d_layernorm_chw_bwd: bad sizes");
  S2D_REQUIRE(ws_bytes >= s2d_layernorm_bwd_workspace_bytes(B), "s2d_layernorm_chw_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(ws);
  float* stats = reinterpret_cast<float*>(partial + (size_t)B * kLnBlocks * 2);
  float* gs = stats + 2 * (size_t)B;
  const long long per = (long long)C * HW;
  ln_stats_kernel<<<dim3(kLnBlocks, B), 256, 0, st>>>(x, C, HW, partial);
  S2D_LAUNCH_CHECK();
  ln_finalize_kernel<<<B, 32, 0, st>>>(partial, kLnBlocks, per, eps, stats);
  S2D_LAUNCH_CHECK();
  ln_bwd_sums_kernel<<<dim3(kLnBlocks, B), 256, 0, st>>>(x, dy, weight, C, HW, stats, partial);
  S2D_LAUNCH_CHECK();
  ln_bwd_finalize_kernel<<<B, 32, 0, st>>>(partial, kLnBlocks, per, gs);
  S2D_LAUNCH_CHECK();
  ln_bwd_apply_kernel<<<grid_for(per, 256, 148 * 16), 256, 0, st>>>(x, dy, weight, B, C, HW, stats, gs, dx, dweight, dbias);
  S2D_LAUNCH_CHECK();
  count_launches(5);
  return S2D_OK;
}

static int dw_chunks(int n) {
  int c = (n + 63) / 64;
  if (c > 148 * 2) c = 148 * 2;
  return c < 1 ? 1 : c;
}

extern "C" size_t s2d_dwconv2d_wgrad_workspace_bytes(int B, int H, int W, int C, int k) {
  if (B < 1 || H < 1 || W < 1 || C < 1 || k < 1) return 0;
  return (size_t)dw_chunks(B * H * W) * k * k * C * sizeof(float);
}

extern "C" int s2d_dwconv2d_wgrad(const float* x, const float* dy, int B, int H, int W, int C, int k, int pad, float* dweight,
                                  void* ws, size_t ws_bytes, void* stream) {
  S2D_REQUIRE(x && dy && dweight && ws, "s2d_dwconv2d_wgrad: null pointer");
  S2D_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 1 && k >= 1 && pad >= 0, "s2d_dwconv2d_wgrad: bad sizes");
  S2D_REQUIRE(ws_bytes >= s2d_dwconv2d_wgrad_workspace_bytes(B, H, W, C, k), "s2d_dwconv2d_wgrad: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = B * H * W, chunks = dw_chunks(n), rpc = (n + chunks - 1) / chunks;
  float* partial = static_cast<float*>(ws);
  dwconv_wgrad_kernel<<<dim3(chunks, k * k), 256, 0, st>>>(x, dy, B, H, W, C, k, pad, rpc, partial);
  S2D_LAUNCH_CHECK();
  dwconv_wgrad_reduce_kernel<<<(k * k * C + 255) / 256, 256, 0, st>>>(partial, chunks, k * k, C, dweight);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  return S2D_OK;
}

extern "C" size_t s2d_grad_norm_workspace_bytes(void) { return (size_t)148 * 4 * sizeof(double); }

// out[0] = L2 norm of g, out[1] = clip coefficient min(1, max_norm / (norm + 1e-6)) (torch.nn.utils.clip_grad_norm_)
extern "C" int s2d_grad_norm_clip(const float* g, long long n, float max_norm, float* out, void* ws, size_t ws_bytes,
                                  void* stream) {
  S2D_REQUIRE(g && out && ws && n >= 1, "s2d_grad_norm_clip: bad argument");
  S2D_REQUIRE(ws_bytes >= s2d_grad_norm_workspace_bytes(), "s2d_grad_norm_clip: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = grid_for(n, 256 * 8, 148 * 4);
  sumsq_partial_kernel<<<nblk, 256, 0, st>>>(g, n, static_cast<double*>(ws));
  S2D_LAUNCH_CHECK();
  clip_coef_kernel<<<1, 32, 0, st>>>(static_cast<double*>(ws), nblk, max_norm, out);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  return S2D_OK;
}

extern "C" int s2d_adam_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, const float* grad_scale, void* stream) {
  S2D_REQUIRE(p && g && exp_avg && exp_avg_sq && n >= 1 && step >= 1, "s2d_adam_step: bad argument");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_step_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, 1.f - weight_decay * lr, (float)bc1, (float)sqrt(bc2), grad_scale);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
