// oracle/_ref: the reference's OWN CPU rotated-IoU code, compiled unmodified from where it lies
// (det3d/ops/iou3d_nms/src/iou3d_cpu.cpp; included by path, never copied) with a two-function C
// surface so the restatement in s2d_oracle.c can be checked against it.  Test infrastructure only.
#include REF_IOU3D_CPU_SOURCE

extern "C" float ref_iou_bev(const float* a, const float* b) { return iou_bev(a, b); }
extern "C" void ref_iou_bev_matrix(const float* a, int na, const float* b, int nb, float* out) {
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) out[(size_t)i * nb + j] = iou_bev(a + 7 * i, b + 7 * j);
}
