// losses.cu -- forward values of the distillation / detection losses of the training step,
// as single-pass, deterministic reductions (per-block partial sums in double, one final block):
//   * s2d_masked_mse:   sparse2dense_loss terms, det3d/torchie/trainer/trainer.py:783-789
//                       (F.mse_loss(F_S[F_D > 0], F_D[F_D > 0]) and the complement): the reference materialises four
//                       boolean-indexed temporaries of up to 36 MB/scene; here one read of each map.
//   * s2d_focal_loss:   FastFocalLoss.forward (det3d/models/losses/centernet_loss.py:27-54) == fastfocalloss
//                       (trainer.py:38-58), with CenterHead._sigmoid (center_head.py:246-248) and F.sigmoid of the
//                       teacher map (trainer.py:792) optionally fused.
//   * s2d_gather_reg_loss: RegLoss.forward (centernet_loss.py:6-25, L1) and distill_reg_loss (trainer.py:68-76, squared
//                       error, target gathered from the teacher map) incl. _transpose_and_gather_feat
//                       (det3d/core/utils/center_utils.py:66-80).
// The *_bwd entry points give the gradient w.r.t. the student's map (the teacher / targets are constants).
#include "common.cuh"

namespace s2d {

constexpr int kRedThreads = 256;
constexpr int kRedBlocks = 592;   // 4 per SM

template <int N>
__device__ __forceinline__ void block_reduce_store(double (&v)[N], double* __restrict__ partial) {
  __shared__ double sh[N][kRedThreads / 32];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double x = v[i];
    for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    if ((threadIdx.x & 31) == 0) sh[i][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double s = 0.0;
    for (int w = 0; w < kRedThreads / 32; ++w) s += sh[threadIdx.x][w];      // fixed order: deterministic
    partial[(size_t)blockIdx.x * N + threadIdx.x] = s;
  }
}

template <int N>
__global__ void final_reduce_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
  if (threadIdx.x < N) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * N + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// out[0] = sum_{d>0} (s-d)^2, out[1] = #{d>0}, out[2] = sum_{d<=0} (s-d)^2, out[3] = #{d<=0}
__global__ void __launch_bounds__(kRedThreads) masked_mse_kernel(const float4* __restrict__ fs, const float4* __restrict__ fd,
                                                                 long long n4, const float* __restrict__ fs_tail,
                                                                 const float* __restrict__ fd_tail, int tail,
                                                                 double* __restrict__ partial) {
  float sp = 0.f, sn = 0.f;
  unsigned cp = 0, cn = 0;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  int since = 0;
  auto one = [&](float s, float d) {
    const float e = s - d, q = e * e;
    if (d > 0.f) { sp += q; ++cp; } else { sn += q; ++cn; }
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(fs + i), b = __ldg(fd + i);
    one(a.x, b.x); one(a.y, b.y); one(a.z, b.z); one(a.w, b.w);
    if (++since == 64) {                         // flush the fp32 running sums into double every 256 elements
      acc[0] += sp; acc[2] += sn; sp = sn = 0.f; since = 0;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < tail) one(fs_tail[threadIdx.x], fd_tail[threadIdx.x]);
  acc[0] += sp; acc[2] += sn; acc[1] = (double)cp; acc[3] = (double)cn;
  block_reduce_store<4>(acc, partial);
}

// element (b, c, cell) of a map lives at base + b*sb + c*sc + cell*scell (NCHW: sc = HW, scell = 1; NHWC rows: sc = 1, scell = ld)
struct MapView { const float* p; long long sb, sc, scell; };
__device__ __forceinline__ float map_at(const MapView& m, int b, int c, int cell) {
  return __ldg(m.p + (long long)b * m.sb + (long long)c * m.sc + (long long)cell * m.scell);
}
__device__ __forceinline__ float sigmoidf_rn(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// out[0] = neg_loss sum, out[1] = pos_loss sum, out[2] = num_pos
__global__ void __launch_bounds__(kRedThreads) focal_loss_kernel(MapView out, MapView tgt, int B, int C, int HW,
                                                                 int out_logits, int tgt_logits,
                                                                 const long long* __restrict__ ind,
                                                                 const unsigned char* __restrict__ mask,
                                                                 const long long* __restrict__ cat, int M,
                                                                 double* __restrict__ partial) {
  double acc[3] = {0.0, 0.0, 0.0};
  const long long total = (long long)B * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cell = (int)(i % HW), c = (int)((i / HW) % C), b = (int)(i / ((long long)HW * C));
    float o = map_at(out, b, c, cell), t = map_at(tgt, b, c, cell);
    if (out_logits) o = fminf(fmaxf(sigmoidf_rn(o), 1e-4f), 1.f - 1e-4f);        // CenterHead._sigmoid
    if (tgt_logits) t = sigmoidf_rn(t);
    const float g1 = 1.f - t, g2 = g1 * g1;
    acc[0] += (double)(logf(1.f - o) * (o * o) * (g2 * g2));                       // log(1-out) * out^2 * (1-target)^4
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * M; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / M);
    const float m = mask[i] ? 1.f : 0.f;
    float p = map_at(out, b, (int)cat[i], (int)ind[i]);
    if (out_logits) p = fminf(fmaxf(sigmoidf_rn(p), 1e-4f), 1.f - 1e-4f);
    const float q = 1.f - p;
    acc[1] += (double)(logf(p) * (q * q) * m);
    acc[2] += (double)m;
  }
  block_reduce_store<3>(acc, partial);
}

// out[d] = sum_{b,m} err(pred*mask, target*mask), d < D; out[D] = sum(mask).  squared: 0 = L1 (RegLoss), 1 = squared error.
// target: either tgt_rows f32 [B, M, D] or a map (tgt_map.p != null) gathered at the same indices.
constexpr int kRegMaxD = 16;
__global__ void __launch_bounds__(kRedThreads) gather_reg_loss_kernel(MapView pred, const float* __restrict__ tgt_rows,
                                                                      MapView tgt_map, int B, int M, int D, int squared,
                                                                      const long long* __restrict__ ind,
                                                                      const unsigned char* __restrict__ mask,
                                                                      double* __restrict__ partial) {
  double acc[kRegMaxD + 1];
#pragma unroll
  for (int d = 0; d <= kRegMaxD; ++d) acc[d] = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * M; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / M);
    const float m = mask[i] ? 1.f : 0.f;
    const int cell = (int)ind[i];
#pragma unroll
    for (int d = 0; d < kRegMaxD; ++d) {
      if (d < D) {
        const float p = map_at(pred, b, d, cell) * m;
        const float t = (tgt_map.p ? map_at(tgt_map, b, d, cell) : tgt_rows[i * D + d]) * m;
        const float e = p - t;
        acc[d] += (double)(squared ? e * e : fabsf(e));
      }
    }
    acc[kRegMaxD] += (double)m;
  }
  block_reduce_store<kRegMaxD + 1>(acc, partial);
}

// ---- backward ------------------------------------------------------------------------------------------------------
struct GradView { float* p; long long sb, sc, scell; };
__device__ __forceinline__ float* grad_at(const GradView& m, int b, int c, int cell) {
  return m.p + (long long)b * m.sb + (long long)c * m.sc + (long long)cell * m.scell;
}

// d/ds [ w_pos * sum_{t>0}(s-t)^2 / n_pos + w_neg * sum_{t<=0}(s-t)^2 / n_neg ] * upstream
__global__ void masked_mse_bwd_kernel(const float* __restrict__ fs, const float* __restrict__ ft, long long n,
                                      const double* __restrict__ out4, float w_pos, float w_neg,
                                      const float* __restrict__ upstream, float* __restrict__ dfs) {
  const float up = upstream ? *upstream : 1.f;
  const float cp = out4[1] > 0.0 ? (float)(2.0 * w_pos / out4[1]) * up : 0.f;
  const float cn = out4[3] > 0.0 ? (float)(2.0 * w_neg / out4[3]) * up : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = fs[i], t = __ldg(ft + i);
    dfs[i] = (s - t) * (t > 0.f ? cp : cn);
  }
}

// loss = -(pos + neg) / num (num > 0)  or  -neg;  gradient w.r.t. the student map (probabilities, or logits when fused)
__global__ void focal_loss_bwd_map_kernel(MapView out, MapView tgt, int B, int C, int HW, int out_logits, int tgt_logits,
                                          const double* __restrict__ out3, const float* __restrict__ upstream, GradView g) {
  const float up = upstream ? *upstream : 1.f;
  const float coef = -up / (out3[2] > 0.0 ? (float)out3[2] : 1.f);
  const long long total = (long long)B * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cell = (int)(i % HW), c = (int)((i / HW) % C), b = (int)(i / ((long long)HW * C));
    float o = map_at(out, b, c, cell), t = map_at(tgt, b, c, cell);
    float chain = 1.f;
    if (out_logits) {
      const float sg = sigmoidf_rn(o);
      o = fminf(fmaxf(sg, 1e-4f), 1.f - 1e-4f);
      chain = (sg >= 1e-4f && sg <= 1.f - 1e-4f) ? sg * (1.f - sg) : 0.f;       // clamp passes the gradient inside only
    }
    if (tgt_logits) t = sigmoidf_rn(t);
    const float g1 = 1.f - t, g4 = (g1 * g1) * (g1 * g1);
    const float d = g4 * (2.f * o * logf(1.f - o) - o * o / (1.f - o));
    *grad_at(g, b, c, cell) = coef * d * chain;
  }
}

__global__ void focal_loss_bwd_peaks_kernel(MapView out, int B, int out_logits, const long long* __restrict__ ind,
                                            const unsigned char* __restrict__ mask, const long long* __restrict__ cat, int M,
                                            const double* __restrict__ out3, const float* __restrict__ upstream, GradView g) {
  if (!(out3[2] > 0.0)) return;                       // num_pos == 0: the loss is -neg only
  const float up = upstream ? *upstream : 1.f;
  const float coef = -up / (float)out3[2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * M; i += (long long)gridDim.x * blockDim.x) {
    if (!mask[i]) continue;
    const int b = (int)(i / M), c = (int)cat[i], cell = (int)ind[i];
    float p = map_at(out, b, c, cell), chain = 1.f;
    if (out_logits) {
      const float sg = sigmoidf_rn(p);
      p = fminf(fmaxf(sg, 1e-4f), 1.f - 1e-4f);
      chain = (sg >= 1e-4f && sg <= 1.f - 1e-4f) ? sg * (1.f - sg) : 0.f;
    }
    const float q = 1.f - p;
    const float d = q * q / p - 2.f * q * logf(p);
    atomicAdd(grad_at(g, b, c, cell), coef * d * chain);
  }
}

__global__ void gather_reg_loss_bwd_kernel(MapView pred, const float* __restrict__ tgt_rows, MapView tgt_map, int B, int M,
                                           int D, int squared, const long long* __restrict__ ind,
                                           const unsigned char* __restrict__ mask, const double* __restrict__ sums,
                                           const float* __restrict__ upstream, GradView g) {
  const float inv = 1.f / ((float)sums[kRegMaxD] + 1e-4f);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)B * M * D; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / D;
    const int d = (int)(e - i * D);
    if (!mask[i]) continue;
    const int b = (int)(i / M), cell = (int)ind[i];
    const float p = map_at(pred, b, d, cell);
    const float t = tgt_map.p ? map_at(tgt_map, b, d, cell) : tgt_rows[i * D + d];
    const float err = p - t;
    const float dd = squared ? 2.f * err : (err > 0.f ? 1.f : (err < 0.f ? -1.f : 0.f));
    atomicAdd(grad_at(g, b, d, cell), dd * inv * (upstream ? upstream[d] : 1.f));
  }
}

// ---- PCR losses: KD_VoxelNet.mask_offset_loss (det3d/models/detectors/voxelnet.py:171-185) -------------------------------
// The reference densifies the reconstruction voxels into gt [N,5,D,H,W] (136-543 MB at B = 4), builds a same-sized grid of
// voxel centres and evaluates  BCEWithLogits(gen_mask, gt.sum(1) != 0, pos_weight = #neg / #pos)  and
// L1(gen_offset[sel], (gt[:, :3] - grid * mask)[sel]),  sel = that difference != 0.  Only occupied voxels have a non-zero
// gt, so here: one dense pass over the mask logits (softplus sum) plus one pass over the M occupied voxels.
// Predictions are rows in (b, y, x, z) order (z fastest): row = ((b*H + y)*W + x)*D + z.
__device__ __forceinline__ float softplusf(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

struct PcrGrid { int D, H, W; float sx, mx, hx, sy, my, hy, sz, mz, hz; };   // centre = idx*s - m + h, in that order (fp32)

__global__ void __launch_bounds__(kRedThreads) pcr_softplus_sum_kernel(const float* __restrict__ logits, long long n,
                                                                       double* __restrict__ partial) {
  double acc[1] = {0.0};
  float f = 0.f;
  int since = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    f += softplusf(__ldg(logits + i));
    if (++since == 64) { acc[0] += (double)f; f = 0.f; since = 0; }
  }
  acc[0] += (double)f;
  block_reduce_store<1>(acc, partial);
}

__device__ __forceinline__ bool pcr_voxel(const PcrGrid& g, const int* __restrict__ coors, const float* __restrict__ gt,
                                          long long v, long long& row, float (&t3)[3]) {
  const int b = coors[v * 4], z = coors[v * 4 + 1], y = coors[v * 4 + 2], x = coors[v * 4 + 3];
  const float* f = gt + v * 5;
  const float sum = (((f[0] + f[1]) + f[2]) + f[3]) + f[4];
  row = (((long long)b * g.H + y) * g.W + x) * g.D + z;
  if (sum == 0.f) return false;
  t3[0] = __fsub_rn(f[0], __fadd_rn(__fsub_rn(__fmul_rn((float)x, g.sx), g.mx), g.hx));
  t3[1] = __fsub_rn(f[1], __fadd_rn(__fsub_rn(__fmul_rn((float)y, g.sy), g.my), g.hy));
  t3[2] = __fsub_rn(f[2], __fadd_rn(__fsub_rn(__fmul_rn((float)z, g.sz), g.mz), g.hz));
  return true;
}

// partial sums: {softplus(-x) over pos, softplus(x) over pos, #pos, sum |offset - t3| over sel, #sel}
__global__ void __launch_bounds__(kRedThreads) pcr_voxel_loss_kernel(PcrGrid g, const float* __restrict__ logits,
                                                                     const float* __restrict__ offset,
                                                                     const int* __restrict__ coors, const float* __restrict__ gt,
                                                                     long long M, double* __restrict__ partial) {
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < M; v += (long long)gridDim.x * blockDim.x) {
    long long row;
    float t3[3];
    if (!pcr_voxel(g, coors, gt, v, row, t3)) continue;
    const float x = __ldg(logits + row);
    acc[0] += (double)softplusf(-x);
    acc[1] += (double)softplusf(x);
    acc[2] += 1.0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (t3[c] != 0.f) { acc[3] += (double)fabsf(__ldg(offset + row * 3 + c) - t3[c]); acc[4] += 1.0; }
  }
  block_reduce_store<5>(acc, partial);
}

// sums6 = {softplus over all, softplus(-x) pos, softplus(x) pos, #pos, l1 sum, #sel}
__global__ void pcr_mask_bwd_dense_kernel(const float* __restrict__ logits, long long n, const float* __restrict__ up,
                                          float* __restrict__ d_logits) {
  const float c = (up ? up[0] : 1.f) / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = __ldg(logits + i);
    d_logits[i] = c * __fdiv_rn(1.f, 1.f + expf(-x));
  }
}

__global__ void pcr_voxel_bwd_kernel(PcrGrid g, const float* __restrict__ logits, const float* __restrict__ offset,
                                     const int* __restrict__ coors, const float* __restrict__ gt, long long M, long long n,
                                     const double* __restrict__ sums6, const float* __restrict__ up,
                                     float* __restrict__ d_logits, float* __restrict__ d_offset) {
  const float npos = (float)sums6[3];
  const float beta = ((float)n - npos) / npos;
  const float cm = (up ? up[0] : 1.f) / (float)n;
  const float co = sums6[5] > 0.0 ? (up ? up[1] : 1.f) / (float)sums6[5] : 0.f;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < M; v += (long long)gridDim.x * blockDim.x) {
    long long row;
    float t3[3];
    if (!pcr_voxel(g, coors, gt, v, row, t3)) continue;
    const float x = __ldg(logits + row);
    d_logits[row] = -cm * beta * __fdiv_rn(1.f, 1.f + expf(x));                 // replaces the negative-cell gradient
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (t3[c] == 0.f) continue;
      const float e = __ldg(offset + row * 3 + c) - t3[c];
      d_offset[row * 3 + c] = co * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f));
    }
  }
}

}  // namespace s2d

using namespace s2d;

extern "C" size_t s2d_loss_workspace_bytes(void) { return (size_t)kRedBlocks * (kRegMaxD + 1) * sizeof(double); }

extern "C" int s2d_masked_mse(const float* f_student, const float* f_teacher, long long n, double* out4, void* workspace,
                              size_t workspace_bytes, void* stream) {
  S2D_REQUIRE(f_student && f_teacher && out4 && workspace && n >= 0, "s2d_masked_mse: bad argument");
  S2D_REQUIRE(workspace_bytes >= s2d_loss_workspace_bytes(), "s2d_masked_mse: workspace too small");
  S2D_REQUIRE(((reinterpret_cast<uintptr_t>(f_student) | reinterpret_cast<uintptr_t>(f_teacher)) & 15) == 0,
              "s2d_masked_mse: maps must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  const long long n4 = n / 4;
  masked_mse_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(reinterpret_cast<const float4*>(f_student),
                                                       reinterpret_cast<const float4*>(f_teacher), n4, f_student + 4 * n4,
                                                       f_teacher + 4 * n4, (int)(n - 4 * n4), partial);
  final_reduce_kernel<4><<<1, 32, 0, st>>>(partial, kRedBlocks, out4);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  return S2D_OK;
}

extern "C" int s2d_focal_loss(const float* out, long long out_sb, long long out_sc, long long out_scell, int out_is_logits,
                              const float* target, long long tgt_sb, long long tgt_sc, long long tgt_scell,
                              int target_is_logits, int B, int C, int HW, const long long* ind, const unsigned char* mask,
                              const long long* cat, int M, double* out3, void* workspace, size_t workspace_bytes,
                              void* stream) {
  S2D_REQUIRE(out && target && ind && mask && cat && out3 && workspace, "s2d_focal_loss: null argument");
  S2D_REQUIRE(B >= 1 && C >= 1 && HW >= 1 && M >= 0, "s2d_focal_loss: bad sizes");
  S2D_REQUIRE(workspace_bytes >= s2d_loss_workspace_bytes(), "s2d_focal_loss: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  focal_loss_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(MapView{out, out_sb, out_sc, out_scell},
                                                       MapView{target, tgt_sb, tgt_sc, tgt_scell}, B, C, HW, out_is_logits,
                                                       target_is_logits, ind, mask, cat, M, partial);
  final_reduce_kernel<3><<<1, 32, 0, st>>>(partial, kRedBlocks, out3);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  return S2D_OK;
}

extern "C" int s2d_gather_reg_loss(const float* pred, long long pred_sb, long long pred_sc, long long pred_scell,
                                   const float* target_rows, const float* target_map, long long tgt_sb, long long tgt_sc,
                                   long long tgt_scell, int B, int M, int D, int squared, const long long* ind,
                                   const unsigned char* mask, double* out, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  S2D_REQUIRE(pred && (target_rows || target_map) && ind && mask && out && workspace, "s2d_gather_reg_loss: null argument");
  S2D_REQUIRE(B >= 1 && M >= 0 && D >= 1 && D <= kRegMaxD, "s2d_gather_reg_loss: D outside [1,%d]", kRegMaxD);
  S2D_REQUIRE(workspace_bytes >= s2d_loss_workspace_bytes(), "s2d_gather_reg_loss: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  gather_reg_loss_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(MapView{pred, pred_sb, pred_sc, pred_scell}, target_rows,
                                                            MapView{target_map, tgt_sb, tgt_sc, tgt_scell}, B, M, D,
                                                            squared, ind, mask, partial);
  final_reduce_kernel<kRegMaxD + 1><<<1, 32, 0, st>>>(partial, kRedBlocks, out);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  return S2D_OK;
}

extern "C" int s2d_masked_mse_bwd(const float* f_student, const float* f_teacher, long long n, const double* out4, float w_pos,
                                  float w_neg, const float* upstream, float* d_student, void* stream) {
  S2D_REQUIRE(f_student && f_teacher && out4 && d_student && n >= 0, "s2d_masked_mse_bwd: bad argument");
  if (n == 0) return S2D_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  masked_mse_bwd_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(f_student, f_teacher, n, out4, w_pos, w_neg,
                                                                                  upstream, d_student);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_focal_loss_bwd(const float* out, long long out_sb, long long out_sc, long long out_scell, int out_is_logits,
                                  const float* target, long long tgt_sb, long long tgt_sc, long long tgt_scell,
                                  int target_is_logits, int B, int C, int HW, const long long* ind, const unsigned char* mask,
                                  const long long* cat, int M, const double* out3, const float* upstream, float* d_out,
                                  long long d_sb, long long d_sc, long long d_scell, void* stream) {
  S2D_REQUIRE(out && target && ind && mask && cat && out3 && d_out, "s2d_focal_loss_bwd: null argument");
  S2D_REQUIRE(B >= 1 && C >= 1 && HW >= 1 && M >= 0, "s2d_focal_loss_bwd: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GradView g{d_out, d_sb, d_sc, d_scell};
  focal_loss_bwd_map_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(MapView{out, out_sb, out_sc, out_scell},
                                                               MapView{target, tgt_sb, tgt_sc, tgt_scell}, B, C, HW,
                                                               out_is_logits, target_is_logits, out3, upstream, g);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  if (M > 0) {
    focal_loss_bwd_peaks_kernel<<<((long long)B * M + 255) / 256, 256, 0, st>>>(MapView{out, out_sb, out_sc, out_scell}, B,
                                                                               out_is_logits, ind, mask, cat, M, out3,
                                                                               upstream, g);
    S2D_LAUNCH_CHECK();
    count_launches(1);
  }
  return S2D_OK;
}

extern "C" int s2d_gather_reg_loss_bwd(const float* pred, long long pred_sb, long long pred_sc, long long pred_scell,
                                       const float* target_rows, const float* target_map, long long tgt_sb, long long tgt_sc,
                                       long long tgt_scell, int B, int M, int D, int squared, const long long* ind,
                                       const unsigned char* mask, const double* sums, const float* upstream, float* d_pred,
                                       long long d_numel, long long d_sb, long long d_sc, long long d_scell, void* stream) {
  S2D_REQUIRE(pred && (target_rows || target_map) && ind && mask && sums && d_pred, "s2d_gather_reg_loss_bwd: null argument");
  S2D_REQUIRE(B >= 1 && M >= 0 && D >= 1 && D <= kRegMaxD && d_numel >= 0, "s2d_gather_reg_loss_bwd: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  S2D_CUDA(cudaMemsetAsync(d_pred, 0, (size_t)d_numel * sizeof(float), st));
  if (M > 0) {
    gather_reg_loss_bwd_kernel<<<((long long)B * M * D + 255) / 256, 256, 0, st>>>(
        MapView{pred, pred_sb, pred_sc, pred_scell}, target_rows, MapView{target_map, tgt_sb, tgt_sc, tgt_scell}, B, M, D, squared,
        ind, mask, sums, upstream, GradView{d_pred, d_sb, d_sc, d_scell});
    S2D_LAUNCH_CHECK();
    count_launches(1);
  }
  return S2D_OK;
}

static PcrGrid pcr_grid(int D, int H, int W, const float* c9) {
  return PcrGrid{D, H, W, c9[0], c9[1], c9[2], c9[3], c9[4], c9[5], c9[6], c9[7], c9[8]};
}

extern "C" int s2d_pcr_loss(const float* mask_logits, const float* offset, int B, int D, int H, int W, const int* coors,
                            const float* gt_feats, long long M, const float* centre9, double* sums6, void* workspace,
                            size_t workspace_bytes, void* stream) {
  S2D_REQUIRE(mask_logits && offset && centre9 && sums6 && workspace && (M == 0 || (coors && gt_feats)),
              "s2d_pcr_loss: null argument");
  S2D_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1 && M >= 0, "s2d_pcr_loss: bad sizes");
  S2D_REQUIRE(workspace_bytes >= s2d_loss_workspace_bytes(), "s2d_pcr_loss: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  const long long n = (long long)B * D * H * W;
  pcr_softplus_sum_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(mask_logits, n, partial);
  final_reduce_kernel<1><<<1, 32, 0, st>>>(partial, kRedBlocks, sums6);
  pcr_voxel_loss_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(pcr_grid(D, H, W, centre9), mask_logits, offset, coors, gt_feats, M,
                                                           partial);
  final_reduce_kernel<5><<<1, 32, 0, st>>>(partial, kRedBlocks, sums6 + 1);
  S2D_LAUNCH_CHECK();
  count_launches(4);
  return S2D_OK;
}

extern "C" int s2d_pcr_loss_bwd(const float* mask_logits, const float* offset, int B, int D, int H, int W, const int* coors,
                                const float* gt_feats, long long M, const float* centre9, const double* sums6,
                                const float* upstream2, float* d_mask_logits, float* d_offset, void* stream) {
  S2D_REQUIRE(mask_logits && offset && centre9 && sums6 && d_mask_logits && d_offset && (M == 0 || (coors && gt_feats)),
              "s2d_pcr_loss_bwd: null argument");
  S2D_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1 && M >= 0, "s2d_pcr_loss_bwd: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)B * D * H * W;
  S2D_CUDA(cudaMemsetAsync(d_offset, 0, (size_t)n * 3 * sizeof(float), st));
  pcr_mask_bwd_dense_kernel<<<148 * 16, 256, 0, st>>>(mask_logits, n, upstream2, d_mask_logits);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  if (M > 0) {
    long long blocks = (M + 255) / 256;
    pcr_voxel_bwd_kernel<<<(int)blocks, 256, 0, st>>>(pcr_grid(D, H, W, centre9), mask_logits, offset, coors, gt_feats, M, n,
                                                     sums6, upstream2, d_mask_logits, d_offset);
    S2D_LAUNCH_CHECK();
    count_launches(1);
  }
  return S2D_OK;
}
