// dense2d.cu -- the non-GEMM pieces of the dense BEV stage (S2D neck / RPN / CenterHead) in NHWC rows:
// regular-grid neighbour tables that let the gather-GEMM kernels run Conv2d / ConvTranspose2d, layout
// transposes at the module boundary, the depthwise 7x7 conv and the [C,H,W] LayerNorm of the ConvNeXt blocks
// (det3d/models/necks/rpn.py:204-222), and the NHWC form of SparseConvTensor.dense() (scn.py:173-176).
// All HBM-bound, coalesced over the channel dimension.
#include "common.cuh"

namespace s2d {

// tbl[k][o] for Conv2d: output site o = (b, oy, ox), tap k = (ky, kx): input row at (oy*s - p + ky, ox*s - p + kx)
__global__ void __launch_bounds__(256) grid2d_table_kernel(int B, int H, int W, int Ho, int Wo, int kh, int kw, int s,
                                                           int p, int* __restrict__ tbl, int stride) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = B * Ho * Wo;
  if (o >= n) return;
  const int ox = o % Wo, oy = (o / Wo) % Ho, b = o / (Wo * Ho);
  int k = 0;
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx, ++k) {
      const int iy = oy * s - p + ky, ix = ox * s - p + kx;
      const bool ok = (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
      tbl[(size_t)k * stride + o] = ok ? (b * H + iy) * W + ix : -1;
    }
}

// ConvTranspose2d with stride 2 as four sub-pixel convolutions.  Class (py, px) owns the outputs
// (2y+py, 2x+px), y < H, x < W.  Output oy receives input iy through tap ky iff oy = 2*iy - p + ky, so the taps of
// the class are ky = ky0 + 2a with ky0 = (py + p) & 1, a < kh/2, and iy = (oy + p - ky) / 2.
// Tap index inside the class: kk = a * (kw/2) + c (the host slices the weights the same way).
__global__ void __launch_bounds__(256) grid2d_tconv_table_kernel(int B, int H, int W, int kh, int kw, int p, int py,
                                                                 int px, int* __restrict__ tbl, int stride,
                                                                 int* __restrict__ out_rows) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = B * H * W;
  if (o >= n) return;
  const int x = o % W, y = (o / W) % H, b = o / (W * H);
  const int Ho = 2 * H, Wo = 2 * W;   // (H-1)*2 - 2p + kh with (kh,p) = (4,1) or (2,0)
  const int oy = 2 * y + py, ox = 2 * x + px;
  out_rows[o] = (b * Ho + oy) * Wo + ox;
  const int ky0 = (py + p) & 1, kx0 = (px + p) & 1;
  int kk = 0;
  for (int a = 0; a < kh / 2; ++a)
    for (int c = 0; c < kw / 2; ++c, ++kk) {
      const int ty = oy + p - (ky0 + 2 * a), tx = ox + p - (kx0 + 2 * c);
      const int iy = ty >> 1, ix = tx >> 1;
      const bool ok = ty >= 0 && tx >= 0 && iy < H && ix < W;
      tbl[(size_t)kk * stride + o] = ok ? (b * H + iy) * W + ix : -1;
    }
}

// [B, C, HW] -> rows [B*HW, ld] (channel fastest), 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, int C, int HW,
                                                           float* __restrict__ out, int ld) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (c < C && p < HW) ? in[((size_t)b * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (p < HW && c < C) out[((size_t)b * HW + p) * ld + c] = tile[tx][i];
  }
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ in, int ld, int C, int HW,
                                                           float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    tile[i][tx] = (p < HW && c < C) ? in[((size_t)b * HW + p) * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    if (c < C && p < HW) out[((size_t)b * C + c) * HW + p] = tile[tx][i];
  }
}

// depthwise k x k conv (groups = C), NHWC rows; weight [C, k, k] (torch [C,1,k,k]), bias [C] nullable
__global__ void __launch_bounds__(256) dwconv2d_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                       const float* __restrict__ bias, int B, int H, int W, int C,
                                                       int k, int pad, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W * C) return;
  const int c = (int)(idx % C);
  const long long pix = idx / C;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
  float acc = bias ? bias[c] : 0.f;
  for (int ky = 0; ky < k; ++ky) {
    const int iy = y - pad + ky;
    if ((unsigned)iy >= (unsigned)H) continue;
    for (int kx = 0; kx < k; ++kx) {
      const int ix = x - pad + kx;
      if ((unsigned)ix >= (unsigned)W) continue;
      acc = fmaf(__ldg(in + ((size_t)(b * H + iy) * W + ix) * C + c), __ldg(w + ((size_t)c * k + ky) * k + kx), acc);
    }
  }
  out[idx] = acc;
}

// The same for C % 64 == 0 (the ConvNeXt blocks: 7 x 7, C = 256): block = 64 channels x 16 consecutive pixels; the
// 64 x k x k weights of the block sit in shared memory as [tap][channel] (the [C][k][k] layout read per thread costs a
// separate 32 B sector per lane and tap: 0.26 ms for a 47 x 47 x 256 map of 8 scenes); a thread owns four channels of one
// pixel: one 16 B input load and one conflict-free LDS.128 per tap.
constexpr int kDwMaxTaps = 49;
__global__ void __launch_bounds__(256) dwconv2d_c64_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int B, int H, int W, int C,
                                                           int k, int pad, float* __restrict__ out) {
  __shared__ __align__(16) float s_w[kDwMaxTaps][64];
  const int c0 = blockIdx.y * 64;
  const int taps = k * k;
  for (int i = threadIdx.x; i < taps * 64; i += 256) {
    const int c = i / taps, t = i - c * taps;                       // consecutive threads read consecutive floats of w
    s_w[t][c] = __ldg(w + (size_t)(c0 + c) * taps + t);
  }
  __syncthreads();
  const int cq = threadIdx.x & 15;
  const long long pix = (long long)blockIdx.x * 16 + (threadIdx.x >> 4);
  if (pix >= (long long)B * H * W) return;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
  const int c = c0 + 4 * cq;
  float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int ky = 0; ky < k; ++ky) {
    const int iy = y - pad + ky;
    if ((unsigned)iy >= (unsigned)H) continue;
    const float* row = in + ((size_t)(b * H + iy) * W) * C + c;
    for (int kx = 0; kx < k; ++kx) {
      const int ix = x - pad + kx;
      if ((unsigned)ix >= (unsigned)W) continue;
      const float4 v = __ldg(reinterpret_cast<const float4*>(row + (size_t)ix * C));
      const float4 ww = *reinterpret_cast<const float4*>(&s_w[ky * k + kx][4 * cq]);
      acc.x = fmaf(v.x, ww.x, acc.x); acc.y = fmaf(v.y, ww.y, acc.y);
      acc.z = fmaf(v.z, ww.z, acc.z); acc.w = fmaf(v.w, ww.w, acc.w);
    }
  }
  *reinterpret_cast<float4*>(out + (size_t)pix * C + c) = acc;
}

// LayerNorm over all C*H*W elements of a sample (nn.LayerNorm([C,H,W])), data in NHWC rows, affine in [C,H,W].
// Pass 1: per-block partial (sum, sumsq) in double, fixed order -> deterministic.  Pass 2 re-adds the partials.
constexpr int kLnBlocks = 64;   // partial blocks per sample
__global__ void __launch_bounds__(256) layernorm_stats_kernel(const float* __restrict__ in, long long per_sample,
                                                              double* __restrict__ partial) {
  const int b = blockIdx.y;
  const float* x = in + (size_t)b * per_sample;
  double s = 0.0, q = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = x[i];
    s += v; q += v * v;
  }
  __shared__ double sh[2][256];
  sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d) { sh[0][threadIdx.x] += sh[0][threadIdx.x + d]; sh[1][threadIdx.x] += sh[1][threadIdx.x + d]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + 0] = sh[0][0];
    partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = sh[1][0];
  }
}

__global__ void __launch_bounds__(256) layernorm_apply_kernel(const float* __restrict__ in,
                                                              const double* __restrict__ partial, int nblk,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int C, int HW, float eps,
                                                              float* __restrict__ out) {
  const int b = blockIdx.y;
  const long long per_sample = (long long)C * HW;
  double s = 0.0, q = 0.0;
  for (int i = 0; i < nblk; ++i) { s += partial[((size_t)b * nblk + i) * 2]; q += partial[((size_t)b * nblk + i) * 2 + 1]; }
  const double mean = s / (double)per_sample;
  const double var = q / (double)per_sample - mean * mean;   // biased, like torch
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float fmean = (float)mean;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int p = (int)(i / C);
    const size_t a = (size_t)c * HW + p;               // affine parameters are stored [C, H, W]
    const float v = (in[(size_t)b * per_sample + i] - fmean) * rstd;
    out[(size_t)b * per_sample + i] = fmaf(v, gamma ? gamma[a] : 1.f, beta ? beta[a] : 0.f);
  }
}

// dense() + view in NHWC: out[(b,y,x)][c*D + z] = feat[row][c]
__global__ void __launch_bounds__(256) dense_bev_nhwc_kernel(const float* __restrict__ feat,
                                                             const int4* __restrict__ coors, int n, int C, int D,
                                                             int H, int W, float* __restrict__ out, int ld) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const int4 c = coors[row];
  float* dst = out + ((size_t)(c.x * H + c.z) * W + c.w) * ld + c.y;
  for (int ch = lane; ch < C; ch += 32) dst[(size_t)ch * D] = __ldg(feat + (size_t)row * C + ch);
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_grid2d_table(int B, int H, int W, int kh, int kw, int stride, int pad, int* tbl, int tbl_stride,
                                void* stream) {
  S2D_REQUIRE(B >= 1 && H >= 1 && W >= 1 && kh >= 1 && kw >= 1 && stride >= 1 && pad >= 0 && kh * kw <= 27 && tbl,
              "s2d_grid2d_table: bad argument");
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  S2D_REQUIRE(Ho >= 1 && Wo >= 1 && tbl_stride >= B * Ho * Wo, "s2d_grid2d_table: bad output size / stride");
  grid2d_table_kernel<<<div_up((long long)B * Ho * Wo, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      B, H, W, Ho, Wo, kh, kw, stride, pad, tbl, tbl_stride);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_grid2d_tconv_table(int B, int H, int W, int kh, int kw, int pad, int py, int px, int* tbl,
                                      int tbl_stride, int* out_rows, void* stream) {
  S2D_REQUIRE(B >= 1 && H >= 1 && W >= 1 && tbl && out_rows && tbl_stride >= B * H * W,
              "s2d_grid2d_tconv_table: bad argument");
  S2D_REQUIRE((kh == 4 && kw == 4 && pad == 1) || (kh == 2 && kw == 2 && pad == 0),
              "s2d_grid2d_tconv_table: only stride-2 ConvTranspose2d with (k,p) = (4,1) or (2,0)");
  S2D_REQUIRE((py == 0 || py == 1) && (px == 0 || px == 1), "s2d_grid2d_tconv_table: parity must be 0 or 1");
  grid2d_tconv_table_kernel<<<div_up((long long)B * H * W, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      B, H, W, kh, kw, pad, py, px, tbl, tbl_stride, out_rows);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_nchw_to_nhwc(const float* in, int B, int C, int HW, float* out, int out_ld, void* stream) {
  S2D_REQUIRE(in && out && B >= 1 && C >= 1 && HW >= 1 && out_ld >= C, "s2d_nchw_to_nhwc: bad argument");
  nchw_to_nhwc_kernel<<<dim3(div_up(HW, 32), div_up(C, 32), B), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, C, HW,
                                                                                                           out, out_ld);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_nhwc_to_nchw(const float* in, int in_ld, int B, int C, int HW, float* out, void* stream) {
  S2D_REQUIRE(in && out && B >= 1 && C >= 1 && HW >= 1 && in_ld >= C, "s2d_nhwc_to_nchw: bad argument");
  nhwc_to_nchw_kernel<<<dim3(div_up(HW, 32), div_up(C, 32), B), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, in_ld,
                                                                                                           C, HW, out);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_dwconv2d(const float* in, const float* weight, const float* bias, int B, int H, int W, int C, int k,
                            int pad, float* out, void* stream) {
  S2D_REQUIRE(in && weight && out && B >= 1 && H >= 1 && W >= 1 && C >= 1 && k >= 1 && pad >= 0,
              "s2d_dwconv2d: bad argument");
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) |
                         reinterpret_cast<uintptr_t>(bias)) & 15) == 0;
  if (C % 64 == 0 && k * k <= kDwMaxTaps && aligned)     // same fmaf order per output as the generic kernel: bit-identical
    dwconv2d_c64_kernel<<<dim3(div_up((long long)B * H * W, 16), C / 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, weight, bias, B, H, W, C, k, pad, out);
  else
    dwconv2d_kernel<<<div_up((long long)B * H * W * C, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, weight, bias, B, H, W, C, k, pad, out);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" size_t s2d_layernorm_workspace_bytes(int B) { return (size_t)(B > 0 ? B : 0) * kLnBlocks * 2 * sizeof(double); }

extern "C" int s2d_layernorm_chw(const float* in, const float* gamma, const float* beta, int B, int C, int HW,
                                 float eps, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  S2D_REQUIRE(in && out && B >= 1 && C >= 1 && HW >= 1, "s2d_layernorm_chw: bad argument");
  if (!workspace || workspace_bytes < s2d_layernorm_workspace_bytes(B)) {
    set_error("s2d_layernorm_chw: workspace too small");
    return S2D_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  layernorm_stats_kernel<<<dim3(kLnBlocks, B), 256, 0, st>>>(in, (long long)C * HW, partial);
  layernorm_apply_kernel<<<dim3(kNumSMs, B), 256, 0, st>>>(in, partial, kLnBlocks, gamma, beta, C, HW, eps, out);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  return S2D_OK;
}

extern "C" int s2d_dense_bev_nhwc(const float* feat, const int* coors, int n_rows, int C, int batch, int D, int H,
                                  int W, float* out, int out_ld, void* stream) {
  S2D_REQUIRE(n_rows >= 0 && C >= 1 && batch >= 1 && D >= 1 && H >= 1 && W >= 1 && out && out_ld >= C * D,
              "s2d_dense_bev_nhwc: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  S2D_CUDA(cudaMemsetAsync(out, 0, (size_t)batch * H * W * out_ld * sizeof(float), st));
  if (n_rows == 0) return S2D_OK;
  S2D_REQUIRE(feat && coors, "s2d_dense_bev_nhwc: null argument");
  dense_bev_nhwc_kernel<<<div_up((long long)n_rows * 32, 256), 256, 0, st>>>(
      feat, reinterpret_cast<const int4*>(coors), n_rows, C, D, H, W, out, out_ld);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
