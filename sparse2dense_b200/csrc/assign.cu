// assign.cu -- CenterPoint training targets on the device:
// AssignLabel.__call__ (det3d/datasets/pipelines/preprocess.py:489-653) with gaussian_radius / gaussian2D /
// draw_umich_gaussian (det3d/core/utils/center_utils.py:18-64) and limit_period (det3d/core/bbox/box_np_ops.py:360-361)
// for a whole batch in one launch: one CTA per (object, sample) computes the object's cell, radius and regression row and
// splats its gaussian window into the class heat map with an integer atomicMax (non-negative floats order like their bit
// patterns, so the result does not depend on the order of the objects).  The reference does this per scene in numpy inside
// the DataLoader workers and ships the maps over PCIe.
//
// Arithmetic follows the reference expression by expression: float32 for sizes / centres / radius in the written order
// (no FMA contraction), float64 for the gaussian (np.ogrid of python floats) rounded once to float32.
#include "common.cuh"

namespace s2d {

struct AssignArgs {
  const float* boxes;        // [B, M, 9]  x y z w l h vx vy rot, task order (class-major)
  const int* classes;        // [B, M]     1-based class inside the task
  const int* num_objs;       // [B]
  int B, M, num_cls, H, W, out_size_factor, min_radius;
  float pc_x, pc_y, vs_x, vs_y;
  double overlap;
  float* hm;                 // [B, num_cls, H, W]
  float* anno_box;           // [B, M, 10]
  long long* ind;            // [B, M]
  unsigned char* mask;       // [B, M]
  long long* cat;            // [B, M]
  float* boxes_and_cls;      // [B, M, 10] or null
  int cls_offset;            // added to the class in boxes_and_cls (merge_multi_group_label)
};

__device__ __forceinline__ float gaussian_radius_f32(float height, float width, double mo) {
  // center_utils.py:18-39, float32 scalars, operation order as written
  const float one_m = (float)(1.0 - mo), one_p = (float)(1.0 + mo);
  const float b1 = __fadd_rn(height, width);
  const float c1 = __fdiv_rn(__fmul_rn(__fmul_rn(width, height), one_m), one_p);
  const float sq1 = __fsqrt_rn(__fsub_rn(__fmul_rn(b1, b1), __fmul_rn(4.f, c1)));
  const float r1 = __fdiv_rn(__fadd_rn(b1, sq1), 2.f);
  const float b2 = __fmul_rn(2.f, __fadd_rn(height, width));
  const float c2 = __fmul_rn(__fmul_rn(one_m, width), height);
  const float sq2 = __fsqrt_rn(__fsub_rn(__fmul_rn(b2, b2), __fmul_rn(16.f, c2)));
  const float r2 = __fdiv_rn(__fadd_rn(b2, sq2), 2.f);
  const float a3 = (float)(4.0 * mo);
  const float b3 = __fmul_rn((float)(-2.0 * mo), __fadd_rn(height, width));
  const float c3 = __fmul_rn(__fmul_rn((float)(mo - 1.0), width), height);
  const float sq3 = __fsqrt_rn(__fsub_rn(__fmul_rn(b3, b3), __fmul_rn(__fmul_rn(4.f, a3), c3)));
  const float r3 = __fdiv_rn(__fadd_rn(b3, sq3), 2.f);
  return fminf(r1, fminf(r2, r3));
}

__global__ void __launch_bounds__(128) assign_label_kernel(AssignArgs A) {
  const int k = blockIdx.x, b = blockIdx.y;
  __shared__ int s_draw, s_cx, s_cy, s_r, s_cls;
  const size_t o = (size_t)b * A.M + k;
  if (threadIdx.x == 0) {
    s_draw = 0;
    if (k < min(A.num_objs[b], A.M)) {
      const float* bx = A.boxes + o * 9;
      const float period = 6.2831855f;                                       // float32(2*pi)
      const float rot = __fsub_rn(bx[8], __fmul_rn(floorf(__fadd_rn(__fdiv_rn(bx[8], period), 0.5f)), period));
      const int cls = A.classes[o];
      if (A.boxes_and_cls) {                                                 // x y z w l h rot vx vy class
        float* g = A.boxes_and_cls + o * 10;
        g[0] = bx[0]; g[1] = bx[1]; g[2] = bx[2]; g[3] = bx[3]; g[4] = bx[4]; g[5] = bx[5];
        g[6] = rot; g[7] = bx[6]; g[8] = bx[7]; g[9] = (float)(cls + A.cls_offset);
      }
      const float osf = (float)A.out_size_factor;
      const float w = __fdiv_rn(__fdiv_rn(bx[3], A.vs_x), osf), l = __fdiv_rn(__fdiv_rn(bx[4], A.vs_y), osf);
      if (w > 0.f && l > 0.f) {
        const int radius = max(A.min_radius, (int)gaussian_radius_f32(l, w, A.overlap));
        const float cx = __fdiv_rn(__fdiv_rn(__fsub_rn(bx[0], A.pc_x), A.vs_x), osf);
        const float cy = __fdiv_rn(__fdiv_rn(__fsub_rn(bx[1], A.pc_y), A.vs_y), osf);
        const int ix = (int)cx, iy = (int)cy;                                // astype(np.int32): truncation
        if (ix >= 0 && ix < A.W && iy >= 0 && iy < A.H) {
          s_draw = 1; s_cx = ix; s_cy = iy; s_r = radius; s_cls = cls - 1;
          A.cat[o] = cls - 1;
          A.ind[o] = (long long)iy * A.W + ix;
          A.mask[o] = 1;
          float* a = A.anno_box + o * 10;
          a[0] = __fsub_rn(cx, (float)ix);
          a[1] = __fsub_rn(cy, (float)iy);
          a[2] = bx[2];
          a[3] = (float)log((double)bx[3]); a[4] = (float)log((double)bx[4]); a[5] = (float)log((double)bx[5]);
          a[6] = bx[6]; a[7] = bx[7];
          a[8] = (float)sin((double)rot); a[9] = (float)cos((double)rot);
        }
      }
    }
  }
  __syncthreads();
  if (!s_draw) return;
  const int r = s_r, cx = s_cx, cy = s_cy;
  const int left = min(cx, r), right = min(A.W - cx, r + 1), top = min(cy, r), bottom = min(A.H - cy, r + 1);
  const int ww = left + right, hh = top + bottom;
  const double sigma = (double)(2 * r + 1) / 6.0;
  const double denom = 2.0 * sigma * sigma;
  int* hm = reinterpret_cast<int*>(A.hm + ((size_t)b * A.num_cls + s_cls) * A.H * A.W);
  for (int e = threadIdx.x; e < ww * hh; e += blockDim.x) {
    const int dy = e / ww - top, dx = e - (e / ww) * ww - left;
    const float g = (float)exp(-((double)(dx * dx) + (double)(dy * dy)) / denom);
    atomicMax(hm + (size_t)(cy + dy) * A.W + (cx + dx), __float_as_int(g));
  }
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_assign_label(const float* gt_boxes, const int* gt_classes, const int* num_objs, int B, int max_objs,
                                int num_cls, int H, int W, float pc_x, float pc_y, float voxel_x, float voxel_y,
                                int out_size_factor, double gaussian_overlap, int min_radius, int cls_offset, float* hm,
                                float* anno_box, long long* ind, unsigned char* mask, long long* cat, float* gt_boxes_and_cls,
                                void* stream) {
  S2D_REQUIRE(gt_boxes && gt_classes && num_objs && hm && anno_box && ind && mask && cat, "s2d_assign_label: null pointer");
  S2D_REQUIRE(B >= 1 && max_objs >= 1 && num_cls >= 1 && H >= 1 && W >= 1 && out_size_factor >= 1 && min_radius >= 0 &&
                  voxel_x > 0.f && voxel_y > 0.f, "s2d_assign_label: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bm = (size_t)B * max_objs;
  S2D_CUDA(cudaMemsetAsync(hm, 0, (size_t)B * num_cls * H * W * sizeof(float), st));
  S2D_CUDA(cudaMemsetAsync(anno_box, 0, bm * 10 * sizeof(float), st));
  S2D_CUDA(cudaMemsetAsync(ind, 0, bm * sizeof(long long), st));
  S2D_CUDA(cudaMemsetAsync(mask, 0, bm, st));
  S2D_CUDA(cudaMemsetAsync(cat, 0, bm * sizeof(long long), st));
  if (gt_boxes_and_cls) S2D_CUDA(cudaMemsetAsync(gt_boxes_and_cls, 0, bm * 10 * sizeof(float), st));
  AssignArgs A{gt_boxes, gt_classes, num_objs, B, max_objs, num_cls, H, W, out_size_factor, min_radius,
               pc_x, pc_y, voxel_x, voxel_y, gaussian_overlap, hm, anno_box, ind, mask, cat, gt_boxes_and_cls, cls_offset};
  assign_label_kernel<<<dim3(max_objs, B), 128, 0, st>>>(A);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
