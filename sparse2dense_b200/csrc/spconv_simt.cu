// spconv_simt.cu -- fp32 output-stationary sparse convolution (gather-GEMM) on CUDA cores.
//
// Replaces spconv v1.x's per-offset  gather -> torch::mm -> scatter_add  (3 x 27 launches per
// layer, rulebook-pair traffic P*(Cin+2*Cout)*4 B) and the BatchNorm1d / ReLU / residual modules
// that det3d/models/backbones/scn.py:69-85,104-152 applies to .features afterwards, by ONE
// launch per layer:
//
//   out[o,:] = act( (sum_k in[tbl[k][o],:] @ W[k]) * scale + shift (+ residual[o,:]) )
//
// A CTA owns a tile of TM output rows and all Cout channels; it walks the K kernel offsets,
// gathers the neighbour rows of the tile into shared memory (float4 / coalesced per row),
// streams the matching Cin-chunk of W[k], and accumulates an RM x 4 register tile per thread.
// No atomics, no scatter: every output row is written exactly once, coalesced, so the result
// is deterministic.  Offsets for which no row of the tile has a neighbour are skipped
// (block-uniform).  This is the reference-faithful fp32 path (S2D_PRECISION_FP32); the
// tcgen05 TF32 path for Cin >= 32 lives in spconv_tc.cu.
#include "common.cuh"

namespace s2d {

constexpr int kMaxK = 27;

template <int CIN, int COUT>
struct SimtCfg {
  static constexpr int THREADS = 256;
  static constexpr int RN = 4;                       // couts per thread (one float4 of W)
  static constexpr int TX = COUT / RN;               // threads along cout
  static constexpr int TY = THREADS / TX;            // threads along rows
  static constexpr int RM = (COUT >= 64) ? 8 : 4;    // rows per thread
  static constexpr int TM = TY * RM;                 // rows per CTA
  static constexpr int KC = (CIN < 32) ? ((CIN + 3) / 4 * 4) : 32;  // Cin chunk staged per step
  static constexpr int NCHUNK = (CIN + KC - 1) / KC;
  static constexpr int LDA = KC + 1;                 // +1: conflict-free column reads
  static_assert(COUT % RN == 0 && THREADS % TX == 0, "bad tile");
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(256) spconv_simt_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                          const int* __restrict__ tbl, int tbl_stride, int n_out,
                                                          int K, const float* __restrict__ scale,
                                                          const float* __restrict__ shift,
                                                          const float* __restrict__ residual, int relu,
                                                          float* __restrict__ out) {
  using Cfg = SimtCfg<CIN, COUT>;
  constexpr int TM = Cfg::TM, KC = Cfg::KC, LDA = Cfg::LDA, RM = Cfg::RM, TX = Cfg::TX;
  __shared__ int s_nbr[kMaxK * TM];
  __shared__ float s_a[TM * LDA];
  __shared__ __align__(16) float s_b[KC * COUT];
  __shared__ unsigned s_kmask;

  const int tile0 = blockIdx.x * TM;
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;

  if (tid == 0) s_kmask = 0;
  __syncthreads();
  // neighbour rows of this tile for every offset; remember which offsets are populated at all
  for (int idx = tid; idx < K * TM; idx += Cfg::THREADS) {
    const int k = idx / TM, r = idx - k * TM;
    const int row = tile0 + r;
    const int j = row < n_out ? __ldg(tbl + (size_t)k * tbl_stride + row) : -1;
    s_nbr[idx] = j;
    if (j >= 0 && !((s_kmask >> k) & 1)) atomicOr(&s_kmask, 1u << k);
  }
  __syncthreads();
  const unsigned kmask = s_kmask;

  float acc[RM][4];
#pragma unroll
  for (int m = 0; m < RM; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[m][n] = 0.f;

  for (int k = 0; k < K; ++k) {
    if (!((kmask >> k) & 1)) continue;  // block-uniform
    const float* Wk = W + (size_t)k * CIN * COUT;
#pragma unroll 1
    for (int c0 = 0; c0 < CIN; c0 += KC) {
      // stage A: TM gathered rows x KC input channels
      if constexpr (CIN % 4 == 0) {
        constexpr int V = KC / 4;
        for (int idx = tid; idx < TM * V; idx += Cfg::THREADS) {
          const int r = idx / V, v = idx - r * V;
          const int j = s_nbr[k * TM + r];
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j >= 0) x = __ldg(reinterpret_cast<const float4*>(in + (size_t)j * CIN + c0) + v);
          float* d = s_a + r * LDA + v * 4;
          d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
        }
      } else {
        for (int idx = tid; idx < TM * KC; idx += Cfg::THREADS) {
          const int r = idx / KC, c = idx - r * KC;
          const int j = s_nbr[k * TM + r];
          s_a[r * LDA + c] = (j >= 0 && c0 + c < CIN) ? __ldg(in + (size_t)j * CIN + c0 + c) : 0.f;
        }
      }
      // stage B: rows c0..c0+KC of W[k], contiguous in memory
      for (int idx = tid; idx < KC * COUT / 4; idx += Cfg::THREADS) {
        const int c = (idx * 4) / COUT;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + c < CIN) w = __ldg(reinterpret_cast<const float4*>(Wk + (size_t)c0 * COUT) + idx);
        reinterpret_cast<float4*>(s_b)[idx] = w;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < KC; ++kk) {
        const float4 b = reinterpret_cast<const float4*>(s_b + kk * COUT)[tx];
#pragma unroll
        for (int m = 0; m < RM; ++m) {
          const float a = s_a[(ty * RM + m) * LDA + kk];
          acc[m][0] = fmaf(a, b.x, acc[m][0]);
          acc[m][1] = fmaf(a, b.y, acc[m][1]);
          acc[m][2] = fmaf(a, b.z, acc[m][2]);
          acc[m][3] = fmaf(a, b.w, acc[m][3]);
        }
      }
      __syncthreads();
    }
  }

  // fused epilogue: BN affine (+ residual) (+ ReLU), one float4 store per (row, tx)
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale) + tx);
  if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift) + tx);
#pragma unroll
  for (int m = 0; m < RM; ++m) {
    const int row = tile0 + ty * RM + m;
    if (row >= n_out) continue;
    float4 y;
    y.x = fmaf(acc[m][0], sc.x, sh.x);
    y.y = fmaf(acc[m][1], sc.y, sh.y);
    y.z = fmaf(acc[m][2], sc.z, sh.z);
    y.w = fmaf(acc[m][3], sc.w, sh.w);
    if (residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(residual + (size_t)row * COUT) + tx);
      y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
    }
    if (relu) {
      y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
    }
    reinterpret_cast<float4*>(out + (size_t)row * COUT)[tx] = y;
  }
}

// Any shape / stride / activation: one thread per (row, cout).  Correctness path for shapes outside the
// tuned kernels (and the tiny CenterHead output convs, Cout <= 3).
__global__ void __launch_bounds__(256) conv_generic_kernel(const __grid_constant__ s2d_conv_params P) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)P.n_out * P.Cout) return;
  const int row = (int)(idx / P.Cout), co = (int)(idx - (long long)row * P.Cout);
  float acc = 0.f;
  for (int k = 0; k < P.K; ++k) {
    const int j = __ldg(P.tbl + (size_t)k * P.tbl_stride + row);
    if (j < 0) continue;
    const float* x = P.in + (size_t)j * P.in_ld;
    const float* w = P.weights + (size_t)k * P.Cin * P.Cout + co;
    for (int ci = 0; ci < P.Cin; ++ci) acc = fmaf(__ldg(x + ci), __ldg(w + (size_t)ci * P.Cout), acc);
  }
  float y = fmaf(acc, P.scale ? P.scale[co] : 1.f, P.shift ? P.shift[co] : 0.f);
  const int orow = P.out_rows ? P.out_rows[row] : row;
  const float r = P.residual ? P.residual[(size_t)orow * P.res_ld + co] : 0.f;
  if (!P.res_after_act) y += r;
  if (P.act == S2D_ACT_RELU) y = fmaxf(y, 0.f);
  if (P.act == S2D_ACT_GELU) y = 0.5f * y * (1.f + erff(y * 0.70710678118654752440f));
  if (P.res_after_act) y += r;
  P.out[(size_t)orow * P.out_ld + co] = y;
}

template <int CIN, int COUT>
static int launch_simt(const s2d_conv_params& p, cudaStream_t st) {
  using Cfg = SimtCfg<CIN, COUT>;
  spconv_simt_kernel<CIN, COUT><<<div_up(p.n_out, Cfg::TM), Cfg::THREADS, 0, st>>>(
      p.in, p.weights, p.tbl, p.tbl_stride, p.n_out, p.K, p.scale, p.shift, p.residual, p.act == S2D_ACT_RELU, p.out);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

static int conv_fwd_fp32(const s2d_conv_params& p, cudaStream_t st) {
  const bool plain = p.in_ld == p.Cin && p.out_ld == p.Cout && (!p.residual || p.res_ld == p.Cout) && !p.out_rows &&
                     !p.res_after_act && (p.act == S2D_ACT_NONE || p.act == S2D_ACT_RELU);
  if (plain) {
#define S2D_SIMT_CASE(ci, co) \
  if (p.Cin == ci && p.Cout == co) return launch_simt<ci, co>(p, st)
    S2D_SIMT_CASE(5, 16);
    S2D_SIMT_CASE(16, 16);
    S2D_SIMT_CASE(16, 32);
    S2D_SIMT_CASE(32, 32);
    S2D_SIMT_CASE(32, 64);
    S2D_SIMT_CASE(64, 64);
    S2D_SIMT_CASE(64, 128);
    S2D_SIMT_CASE(128, 128);
#undef S2D_SIMT_CASE
  }
  conv_generic_kernel<<<div_up((long long)p.n_out * p.Cout, 256), 256, 0, st>>>(p);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

int conv_fwd_tf32(const s2d_conv_params& p, cudaStream_t st);  // spconv_tc.cu
int conv_fwd_bf2(const s2d_conv_params& p, cudaStream_t st);   // conv_bf2.cu

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_conv_fwd(const s2d_conv_params* params, void* stream) {
  S2D_REQUIRE(params, "s2d_conv_fwd: null params");
  const s2d_conv_params& p = *params;
  S2D_REQUIRE(p.n_in >= 0 && p.n_out >= 0 && p.Cin >= 1 && p.Cout >= 1, "s2d_conv_fwd: bad sizes");
  S2D_REQUIRE(p.K >= 1 && p.K <= kMaxK, "s2d_conv_fwd: K=%d outside [1,%d]", p.K, kMaxK);
  S2D_REQUIRE(p.tbl_stride >= p.n_out, "s2d_conv_fwd: tbl_stride %d < n_out %d", p.tbl_stride, p.n_out);
  const bool b2 = p.precision == S2D_PRECISION_BF16X2;   // reads in_split, may write out_split only
  S2D_REQUIRE(((b2 && !p.in) || p.in_ld >= p.Cin) && ((b2 && !p.out) || p.out_ld >= p.Cout) &&
                  (!p.residual || p.res_ld >= p.Cout),
              "s2d_conv_fwd: row stride smaller than the channel count");
  S2D_REQUIRE(p.act >= S2D_ACT_NONE && p.act <= S2D_ACT_GELU, "s2d_conv_fwd: unknown activation %d", p.act);
  if (p.n_out == 0) return S2D_OK;
  S2D_REQUIRE((p.in || (b2 && p.in_split)) && p.weights && p.tbl && (p.out || (b2 && p.out_split)),
              "s2d_conv_fwd: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (p.precision) {
    case S2D_PRECISION_FP32:
      return conv_fwd_fp32(p, st);
    case S2D_PRECISION_TF32:
    case S2D_PRECISION_TF32X3:
    case S2D_PRECISION_TF32_BF16C:
    case S2D_PRECISION_AUTO:
      return conv_fwd_tf32(p, st);
    case S2D_PRECISION_BF16X2:
      return conv_fwd_bf2(p, st);
    default:
      set_error("s2d_conv_fwd: unknown precision %d", p.precision);
      return S2D_ERR_INVALID;
  }
}

extern "C" int s2d_spconv_fwd(const float* in, int n_in, const float* W, const int* tbl, int tbl_stride, int n_out,
                              int Cin, int Cout, int K, const float* scale, const float* shift,
                              const float* residual, int relu, float* out, int precision, void* stream) {
  s2d_conv_params p;
  p.in = in; p.weights = W; p.tbl = tbl; p.scale = scale; p.shift = shift; p.residual = residual; p.out = out;
  p.out_rows = nullptr; p.in_ld = Cin; p.out_ld = Cout; p.res_ld = Cout; p.tbl_stride = tbl_stride; p.K = K;
  p.n_in = n_in; p.n_out = n_out; p.Cin = Cin; p.Cout = Cout; p.act = relu ? S2D_ACT_RELU : S2D_ACT_NONE;
  p.res_after_act = 0; p.precision = precision;
  p.in_split = nullptr; p.out_split = nullptr; p.tile_masks = nullptr; p.in_split_ld = 0; p.out_split_ld = 0;
  return s2d_conv_fwd(&p, stream);
}
