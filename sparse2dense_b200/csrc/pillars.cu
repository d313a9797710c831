// pillars.cu -- device pieces of the PointPillars + S2D variant (BASELINE configs[3]):
//   * s2d_pfn_fwd: PillarFeatureNet.forward with two PFNLayers fused into one kernel
//     (det3d/models/readers/pillar_encoder.py:41-56,114-154): decorate (cluster offset, pillar-centre offset),
//     zero the padded points, Linear(10->32)+BN1d+ReLU, max over the points, concat [x, max], Linear(64->64)+BN1d+ReLU, max.
//     The reference runs ~20 eager kernels over a [M,20,64] tensor; here a pillar never leaves shared memory.
//   * s2d_maxpool2d_rows / s2d_gather_rows: nn.MaxPool2d(2,2) and nn.Upsample(nearest) of PointPillarsScatter_S2D
//     (pillar_encoder.py:236,283,295) on NHWC rows.
//   * s2d_grid2d_tconv_table_s: sub-pixel tables of ConvTranspose2d(kernel k, stride s, padding (k-s)/2), which covers the
//     RPN deblocks with stride 4 (det3d/models/necks/rpn.py:82-95, us_layer_strides=[1,2,4]).
// The pillar scatter to the BEV canvas (pillar_encoder.py:337-371) is s2d_dense_bev_nhwc with D = 1.
#include "common.cuh"

namespace s2d {

constexpr int kPfnMaxPts = 32;

// one CTA (64 threads) per pillar, grid-stride; weights staged once per CTA in shared memory (transposed [k][c])
__global__ void __launch_bounds__(64) pfn2_kernel(const float* __restrict__ voxels, const int* __restrict__ num_points,
                                                  const int* __restrict__ coors, int M, int P, int F, float vx, float vy,
                                                  float x_off, float y_off, const float* __restrict__ W0,
                                                  const float* __restrict__ s0, const float* __restrict__ b0,
                                                  const float* __restrict__ W1, const float* __restrict__ s1,
                                                  const float* __restrict__ b1, float* __restrict__ out) {
  constexpr int C0 = 32, C1 = 64, FIN = 10;
  __shared__ float w0t[FIN][C0];
  __shared__ float w1t[C1][C1];          // [k][c]
  __shared__ float raw[kPfnMaxPts][5];
  __shared__ float dec[kPfnMaxPts][FIN];
  __shared__ __align__(16) float x0[kPfnMaxPts][C0];
  __shared__ float xmax[2][C0];
  __shared__ float mean3[3];
  const int t = threadIdx.x;
  for (int i = t; i < FIN * C0; i += 64) w0t[i % FIN][i / FIN] = W0[i];            // W0 [32,10] row-major
  for (int i = t; i < C1 * C1; i += 64) w1t[i % C1][i / C1] = W1[i];               // W1 [64,64] row-major
  const float sc0 = s0[t & 31], sh0 = b0[t & 31], sc1 = s1[t], sh1 = b1[t];
  __syncthreads();
  // this thread's column of the point half of W1 lives in registers for the whole kernel: the inner product of layer 1 then
  // needs only broadcast LDS.128 of the point's activations (0.25 shared-memory loads per FMA instead of 2)
  float w1r[C0];
#pragma unroll
  for (int k = 0; k < C0; ++k) w1r[k] = w1t[k][t];
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const int n = num_points[m];
    for (int i = t; i < P * F; i += 64) raw[i / F][i % F] = voxels[(size_t)m * P * F + i];
    __syncthreads();
    if (t < 3) {                                                   // sum over ALL rows (padded rows are zero) / n
      float s = 0.f;
      for (int p = 0; p < P; ++p) s += raw[p][t];
      mean3[t] = s / (float)n;
    }
    __syncthreads();
    const float cx = (float)coors[m * 4 + 3] * vx + x_off, cy = (float)coors[m * 4 + 2] * vy + y_off;
    for (int i = t; i < P * FIN; i += 64) {
      const int p = i / FIN, k = i % FIN;
      float v;
      if (k < 5) v = raw[p][k];
      else if (k < 8) v = raw[p][k - 5] - mean3[k - 5];
      else v = raw[p][k - 8] - (k == 8 ? cx : cy);
      dec[p][k] = p < n ? v : 0.f;                                 // get_paddings_indicator mask (:146-149)
    }
    __syncthreads();
    {                                                              // layer 0: channel c = t & 31, points split in halves
      const int c = t & 31, half = t >> 5;
      const int p0 = half * ((P + 1) / 2), p1 = half ? P : (P + 1) / 2;
      float mx = -INFINITY;
      for (int p = p0; p < p1; ++p) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < FIN; ++k) acc = fmaf(dec[p][k], w0t[k][c], acc);
        const float y = fmaxf(fmaf(acc, sc0, sh0), 0.f);
        x0[p][c] = y;
        mx = fmaxf(mx, y);
      }
      xmax[half][c] = mx;
    }
    __syncthreads();
    {                                                              // layer 1: channel c = t
      float tail = 0.f;                                            // the repeated-max half of the concat is the same for all points
#pragma unroll 8
      for (int k = 0; k < C0; ++k) tail = fmaf(fmaxf(xmax[0][k], xmax[1][k]), w1t[C0 + k][t], tail);
      float mx = -INFINITY;
      for (int p = 0; p < P; ++p) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < C0; k += 4) {                          // same ascending-k fmaf chain as before: bit-identical
          const float4 v = *reinterpret_cast<const float4*>(&x0[p][k]);
          acc = fmaf(v.x, w1r[k], acc); acc = fmaf(v.y, w1r[k + 1], acc);
          acc = fmaf(v.z, w1r[k + 2], acc); acc = fmaf(v.w, w1r[k + 3], acc);
        }
        mx = fmaxf(mx, fmaxf(fmaf(acc + tail, sc1, sh1), 0.f));
      }
      out[(size_t)m * C1 + t] = mx;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) maxpool2x2_rows_kernel(const float* __restrict__ in, int in_ld, int B, int H, int W,
                                                              int C, float* __restrict__ out, int out_ld) {
  const int Ho = H / 2, Wo = W / 2, C4 = C >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * Ho * Wo * C4) return;
  const int c = (int)(i % C4);
  const long long o = i / C4;
  const int x = (int)(o % Wo), y = (int)((o / Wo) % Ho), b = (int)(o / ((long long)Wo * Ho));
  const size_t r0 = ((size_t)b * H + 2 * y) * W + 2 * x;
  const float4 a = __ldg(reinterpret_cast<const float4*>(in + r0 * in_ld) + c);
  const float4 bq = __ldg(reinterpret_cast<const float4*>(in + (r0 + 1) * in_ld) + c);
  const float4 cq = __ldg(reinterpret_cast<const float4*>(in + (r0 + W) * in_ld) + c);
  const float4 d = __ldg(reinterpret_cast<const float4*>(in + (r0 + W + 1) * in_ld) + c);
  float4 r;
  r.x = fmaxf(fmaxf(a.x, bq.x), fmaxf(cq.x, d.x)); r.y = fmaxf(fmaxf(a.y, bq.y), fmaxf(cq.y, d.y));
  r.z = fmaxf(fmaxf(a.z, bq.z), fmaxf(cq.z, d.z)); r.w = fmaxf(fmaxf(a.w, bq.w), fmaxf(cq.w, d.w));
  reinterpret_cast<float4*>(out + (size_t)o * out_ld)[c] = r;
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ in, int in_ld,
                                                          const int* __restrict__ idx, long long n_out, int C,
                                                          float* __restrict__ out, int out_ld) {
  const int C4 = C >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out * C4) return;
  const long long o = i / C4;
  const int c = (int)(i % C4);
  const int j = __ldg(idx + o);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j >= 0) v = __ldg(reinterpret_cast<const float4*>(in + (size_t)j * in_ld) + c);
  reinterpret_cast<float4*>(out + (size_t)o * out_ld)[c] = v;
}

// Sub-pixel class (py, px) of ConvTranspose2d(k, stride s, pad p) with k - s == 2p: outputs (s*y + py, s*x + px) over the
// INPUT grid.  Output oy receives input iy through tap ky iff oy = s*iy - p + ky, so ky = ((py + p) mod s) + s*a.
__global__ void __launch_bounds__(256) grid2d_tconv_table_s_kernel(int B, int H, int W, int k, int s, int p, int py, int px,
                                                                   int* __restrict__ tbl, int stride,
                                                                   int* __restrict__ out_rows) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = B * H * W;
  if (o >= n) return;
  const int x = o % W, y = (o / W) % H, b = o / (W * H);
  const int Ho = s * H, Wo = s * W;
  const int oy = s * y + py, ox = s * x + px;
  out_rows[o] = (b * Ho + oy) * Wo + ox;
  const int ky0 = (py + p) % s, kx0 = (px + p) % s;
  int kk = 0;
  for (int a = 0; a < k / s; ++a)
    for (int c = 0; c < k / s; ++c, ++kk) {
      const int ty = oy + p - (ky0 + s * a), tx = ox + p - (kx0 + s * c);
      const bool ok = ty >= 0 && tx >= 0 && ty / s < H && tx / s < W;
      tbl[(size_t)kk * stride + o] = ok ? (b * H + ty / s) * W + tx / s : -1;
    }
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_pfn_fwd(const float* voxels, const int* num_points, const int* coors, int n_pillars, int max_points,
                           int F, float vx, float vy, float x_offset, float y_offset, const float* W0, const float* scale0,
                           const float* shift0, const float* W1, const float* scale1, const float* shift1, float* out,
                           void* stream) {
  S2D_REQUIRE(n_pillars >= 0, "s2d_pfn_fwd: bad pillar count");
  if (n_pillars == 0) return S2D_OK;
  S2D_REQUIRE(voxels && num_points && coors && W0 && scale0 && shift0 && W1 && scale1 && shift1 && out,
              "s2d_pfn_fwd: null argument");
  S2D_REQUIRE(F == 5 && max_points >= 1 && max_points <= kPfnMaxPts,
              "s2d_pfn_fwd: built for 5 point features and <= %d points per pillar (got F=%d, P=%d)", kPfnMaxPts, F, max_points);
  const int grid = n_pillars < kNumSMs * 16 ? n_pillars : kNumSMs * 16;
  pfn2_kernel<<<grid, 64, 0, static_cast<cudaStream_t>(stream)>>>(voxels, num_points, coors, n_pillars, max_points, F, vx, vy,
                                                                 x_offset, y_offset, W0, scale0, shift0, W1, scale1, shift1,
                                                                 out);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_maxpool2d_rows(const float* in, int in_ld, int B, int H, int W, int C, float* out, int out_ld,
                                  void* stream) {
  S2D_REQUIRE(in && out && B >= 1 && H >= 2 && W >= 2 && C >= 4 && C % 4 == 0 && in_ld % 4 == 0 && out_ld % 4 == 0,
              "s2d_maxpool2d_rows: bad argument");
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 4);
  maxpool2x2_rows_kernel<<<div_up(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, in_ld, B, H, W, C, out, out_ld);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_gather_rows(const float* in, int in_ld, const int* idx, long long n_out, int C, float* out, int out_ld,
                               void* stream) {
  S2D_REQUIRE(n_out >= 0 && C >= 4 && C % 4 == 0 && in_ld % 4 == 0 && out_ld % 4 == 0, "s2d_gather_rows: bad argument");
  if (n_out == 0) return S2D_OK;
  S2D_REQUIRE(in && idx && out, "s2d_gather_rows: null argument");
  gather_rows_kernel<<<div_up(n_out * (C / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, in_ld, idx, n_out, C,
                                                                                                out, out_ld);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_grid2d_tconv_table_s(int B, int H, int W, int k, int stride, int pad, int py, int px, int* tbl,
                                        int tbl_stride, int* out_rows, void* stream) {
  S2D_REQUIRE(B >= 1 && H >= 1 && W >= 1 && tbl && out_rows && tbl_stride >= B * H * W,
              "s2d_grid2d_tconv_table_s: bad argument");
  S2D_REQUIRE(stride >= 1 && k >= stride && k % stride == 0 && k - stride == 2 * pad,
              "s2d_grid2d_tconv_table_s: need k %% stride == 0 and k - stride == 2*pad (output = stride * input)");
  S2D_REQUIRE(py >= 0 && py < stride && px >= 0 && px < stride, "s2d_grid2d_tconv_table_s: class outside [0,stride)");
  grid2d_tconv_table_s_kernel<<<div_up((long long)B * H * W, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      B, H, W, k, stride, pad, py, px, tbl, tbl_stride, out_rows);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
