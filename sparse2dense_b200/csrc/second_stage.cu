// second_stage.cu -- the two-stage refinement of CenterPoint, device side:
//   * s2d_bev_box_features: TwoStageDetector.get_box_center (det3d/models/detectors/two_stage.py:49-76: box centre +
//     the four face mid-points from center_to_corner_box2d, det3d/core/bbox/box_torch_ops.py:386-406 / rotation_2d
//     :347-360) fused with BEVFeatureExtractor.forward (det3d/models/second_stage/bird_eye_view.py:24-40) and
//     bilinear_interpolate_torch (det3d/core/utils/center_utils.py:93-122) and the zero-padded
//     reorder_first_stage_pred_and_feature (two_stage.py:78-119): one launch writes the [B*P, 5*C] RoI feature
//     matrix straight from the NHWC BEV rows -- the reference first makes an NHWC copy of the 512-channel map.
//   * s2d_roi_refine: RoIHeadTemplate.generate_predicted_boxes (det3d/models/roi_heads/roi_head_template.py:153-183,
//     rotate_points_along_z box_torch_ops.py:326-344) + TwoStageDetector.post_process (two_stage.py:121-151).
// The RoI MLP itself (roi_head.py:70-106, 1x1 Conv1d + BN1d + ReLU) runs on the gather-GEMM kernels with an
// identity table (s2d_conv_fwd, K = 1).
#include "common.cuh"

namespace s2d {

__device__ __forceinline__ float rn_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float rn_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float rn_sub(float a, float b) { return __fsub_rn(a, b); }

// block = (sample b, roi slot p, point q); threads over channels (float4)
__global__ void __launch_bounds__(128) bev_box_features_kernel(const float* __restrict__ bev, int bev_ld, int H, int W,
                                                               int C, const float* __restrict__ boxes,
                                                               const int* __restrict__ n_boxes, int P, int num_point,
                                                               float pc_x, float pc_y, float vx, float vy,
                                                               float out_stride, float* __restrict__ out) {
  const int q = blockIdx.x, p = blockIdx.y, b = blockIdx.z;
  float* dst = out + ((size_t)(b * P + p) * num_point + q) * C;
  const int C4 = C >> 2;
  if (p >= n_boxes[b]) {                       // padded RoI slot: zero features (two_stage.py:92-95)
    for (int c = threadIdx.x; c < C4; c += blockDim.x) reinterpret_cast<float4*>(dst)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float* bx = boxes + (size_t)(b * P + p) * 7;
  float px = bx[0], py = bx[1];
  if (q > 0) {
    // corners x0y0, x0y1, x1y1, x1y0 of dims * (+-0.5), rotated: x' = x cos + y sin, y' = -x sin + y cos
    const float hx = rn_mul(bx[3], 0.5f), hy = rn_mul(bx[4], 0.5f);
    const float cs = cosf(bx[6]), sn = sinf(bx[6]);
    const float cxs[4] = {-hx, -hx, hx, hx}, cys[4] = {-hy, hy, hy, -hy};
    float X[4], Y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      X[k] = rn_add(rn_add(rn_mul(cxs[k], cs), rn_mul(cys[k], sn)), px);
      Y[k] = rn_add(rn_add(rn_mul(cxs[k], -sn), rn_mul(cys[k], cs)), py);
    }
    // front (c0,c1), back (c2,c3), left (c0,c3), right (c1,c2)
    const int i0 = (q == 1 || q == 3) ? 0 : (q == 2 ? 2 : 1);
    const int i1 = (q == 1) ? 1 : (q == 4 ? 2 : 3);
    px = __fdiv_rn(rn_add(X[i0], X[i1]), 2.f);
    py = __fdiv_rn(rn_add(Y[i0], Y[i1]), 2.f);
  }
  // absl_to_relative: (v - pc_start) / voxel / out_stride, two divisions
  const float x = __fdiv_rn(__fdiv_rn(rn_sub(px, pc_x), vx), out_stride);
  const float y = __fdiv_rn(__fdiv_rn(rn_sub(py, pc_y), vy), out_stride);
  // bilinear_interpolate_torch: indices are clamped BEFORE the weights are formed (weights need not sum to 1)
  long long x0 = (long long)floorf(x), y0 = (long long)floorf(y);
  long long x1 = x0 + 1, y1 = y0 + 1;
  x0 = min(max(x0, 0ll), (long long)W - 1); x1 = min(max(x1, 0ll), (long long)W - 1);
  y0 = min(max(y0, 0ll), (long long)H - 1); y1 = min(max(y1, 0ll), (long long)H - 1);
  const float wa = rn_mul(rn_sub((float)x1, x), rn_sub((float)y1, y));
  const float wb = rn_mul(rn_sub((float)x1, x), rn_sub(y, (float)y0));
  const float wc = rn_mul(rn_sub(x, (float)x0), rn_sub((float)y1, y));
  const float wd = rn_mul(rn_sub(x, (float)x0), rn_sub(y, (float)y0));
  const float* base = bev + (size_t)b * H * W * bev_ld;
  const float4* Ia = reinterpret_cast<const float4*>(base + (size_t)(y0 * W + x0) * bev_ld);
  const float4* Ib = reinterpret_cast<const float4*>(base + (size_t)(y1 * W + x0) * bev_ld);
  const float4* Ic = reinterpret_cast<const float4*>(base + (size_t)(y0 * W + x1) * bev_ld);
  const float4* Id = reinterpret_cast<const float4*>(base + (size_t)(y1 * W + x1) * bev_ld);
  for (int c = threadIdx.x; c < C4; c += blockDim.x) {
    const float4 a = __ldg(Ia + c), bb = __ldg(Ib + c), cc = __ldg(Ic + c), d = __ldg(Id + c);
    float4 r;   // ((Ia*wa + Ib*wb) + Ic*wc) + Id*wd, products and sums rounded separately like the eager ops
    r.x = rn_add(rn_add(rn_add(rn_mul(a.x, wa), rn_mul(bb.x, wb)), rn_mul(cc.x, wc)), rn_mul(d.x, wd));
    r.y = rn_add(rn_add(rn_add(rn_mul(a.y, wa), rn_mul(bb.y, wb)), rn_mul(cc.y, wc)), rn_mul(d.y, wd));
    r.z = rn_add(rn_add(rn_add(rn_mul(a.z, wa), rn_mul(bb.z, wb)), rn_mul(cc.z, wc)), rn_mul(d.z, wd));
    r.w = rn_add(rn_add(rn_add(rn_mul(a.w, wa), rn_mul(bb.w, wb)), rn_mul(cc.w, wc)), rn_mul(d.w, wd));
    reinterpret_cast<float4*>(dst)[c] = r;
  }
}

__global__ void __launch_bounds__(128) roi_refine_kernel(const float* __restrict__ rois,
                                                         const float* __restrict__ roi_scores,
                                                         const int* __restrict__ n_boxes, int P,
                                                         const float* __restrict__ rcnn_cls, int cls_ld,
                                                         const float* __restrict__ rcnn_reg, int reg_ld,
                                                         float* __restrict__ out_boxes, float* __restrict__ out_scores) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const size_t i = (size_t)b * P + p;
  float* ob = out_boxes + i * 7;
  if (p >= n_boxes[b]) {
#pragma unroll
    for (int q = 0; q < 7; ++q) ob[q] = 0.f;
    out_scores[i] = 0.f;
    return;
  }
  const float* roi = rois + i * 7;
  const float* reg = rcnn_reg + i * reg_ld;
  const float ry = roi[6];
  const float cs = cosf(ry), sn = sinf(ry);
  // (residual + roi with xyz zeroed) rotated about z by the roi yaw, then translated by the roi centre
  const float x = reg[0], y = reg[1], z = reg[2];
  ob[0] = rn_add(rn_add(rn_mul(x, cs), rn_mul(y, sn)), roi[0]);
  ob[1] = rn_add(rn_add(rn_mul(x, -sn), rn_mul(y, cs)), roi[1]);
  ob[2] = rn_add(z, roi[2]);
  ob[3] = rn_add(reg[3], roi[3]);
  ob[4] = rn_add(reg[4], roi[4]);
  ob[5] = rn_add(reg[5], roi[5]);
  ob[6] = rn_add(reg[6], ry);
  const float sg = __fdiv_rn(1.f, rn_add(1.f, expf(-rcnn_cls[i * cls_ld])));
  out_scores[i] = sqrtf(rn_mul(sg, roi_scores[i]));          // two_stage.py:134
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_bev_box_features(const float* bev, int bev_ld, int batch, int H, int W, int C, const float* boxes,
                                    const int* n_boxes, int max_boxes, int num_point, const float* pc_start_host,
                                    const float* voxel_size_host, float out_stride, float* out, void* stream) {
  S2D_REQUIRE(bev && boxes && n_boxes && out && pc_start_host && voxel_size_host, "s2d_bev_box_features: null argument");
  S2D_REQUIRE(batch >= 1 && H >= 1 && W >= 1 && max_boxes >= 1, "s2d_bev_box_features: bad sizes");
  S2D_REQUIRE(C % 4 == 0 && bev_ld % 4 == 0 && bev_ld >= C, "s2d_bev_box_features: channels / row stride must be multiples of 4");
  S2D_REQUIRE(num_point == 1 || num_point == 5, "s2d_bev_box_features: num_point must be 1 or 5 (two_stage.py:52-75)");
  bev_box_features_kernel<<<dim3(num_point, max_boxes, batch), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      bev, bev_ld, H, W, C, boxes, n_boxes, max_boxes, num_point, pc_start_host[0], pc_start_host[1], voxel_size_host[0],
      voxel_size_host[1], out_stride, out);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" int s2d_roi_refine(const float* rois, const float* roi_scores, const int* n_boxes, int batch, int max_boxes,
                              const float* rcnn_cls, int cls_ld, const float* rcnn_reg, int reg_ld, float* out_boxes,
                              float* out_scores, void* stream) {
  S2D_REQUIRE(rois && roi_scores && n_boxes && rcnn_cls && rcnn_reg && out_boxes && out_scores,
              "s2d_roi_refine: null argument");
  S2D_REQUIRE(batch >= 1 && max_boxes >= 1 && cls_ld >= 1 && reg_ld >= 7, "s2d_roi_refine: bad sizes");
  roi_refine_kernel<<<dim3(div_up(max_boxes, 128), batch), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      rois, roi_scores, n_boxes, max_boxes, rcnn_cls, cls_ld, rcnn_reg, reg_ld, out_boxes, out_scores);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
