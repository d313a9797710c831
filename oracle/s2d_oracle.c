/*
 * s2d_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product path (sparse2dense_b200/) never does.
 *
 * What is restated, and from where (paths relative to the reference checkout):
 *   orc_points_to_voxel   det3d/ops/point_cloud/point_cloud_ops.py:7-55,112-184
 *                         (serial first-appearance voxelizer, reverse_index=True).
 *                         PINNED: bit-exact against the reference's own numba code on
 *                         the fixtures in tests/golden/voxelize_*.npz.
 *   orc_voxel_mean        det3d/models/readers/voxel_encoder.py:17-24
 *   orc_rulebook_subm /   the rulebook ("indice pairs") of spconv v1.x @ 7342772 as used
 *   orc_rulebook_sparse   by det3d/models/backbones/scn.py:16-39,104-152.  spconv is an
 *   orc_spconv_fwd        un-vendored third-party dependency (docs/INSTALL.md:12,65-71);
 *                         its published algorithm is restated here:
 *                         out[o] = sum_k W[k]^T in[j(o,k)].  PARITY UNPINNED: the reference
 *                         holds no test or golden vector at this boundary; the restatement
 *                         is cross-checked against a dense torch F.conv3d formulation
 *                         (tests/test_oracle.py) and small committed fixtures.
 *   orc_bn_act            eval-mode BatchNorm1d + residual + ReLU over active rows,
 *                         det3d/models/backbones/scn.py:69-85,104-152
 *   orc_dense_bev         SparseConvTensor.dense() + view(N, C*D, H, W), scn.py:173-176
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (see oracle/Makefile).  No
 * -ffast-math: the voxel coordinate must be the separately rounded fp32 (p - lo) / vs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------ */
/* voxelizer: point_cloud_ops.py:7-55 (kernel) and :112-184 (wrapper)                     */
/* ------------------------------------------------------------------------------------ */

/* grid_size = round((hi - lo) / vs) in fp32, round-half-even (np.round) -- :24-29, :143-144 */
void orc_grid_size(const float* voxel_size, const float* coors_range, int* grid_xyz) {
  for (int j = 0; j < 3; ++j) {
    volatile float d = coors_range[3 + j] - coors_range[j];
    volatile float g = d / voxel_size[j];
    grid_xyz[j] = (int)rintf(g);
  }
}

/*
 * points [N,F] f32 -> voxels [max_voxels,max_points,F] (caller-zeroed), coors [max_voxels,3]
 * (z,y,x), num_points [max_voxels] (caller-zeroed).  Returns voxel_num.
 * coor_to_voxelidx is the dense int32 map of shape (Dz,Dy,Dx) filled with -1 by the caller
 * (:145-150) or NULL, in which case it is allocated here.
 */
int orc_points_to_voxel(const float* points, int N, int F, const float* voxel_size,
                        const float* coors_range, int max_points, int max_voxels,
                        int* coor_to_voxelidx, float* voxels, int* coors, int* num_points) {
  int grid[3];
  orc_grid_size(voxel_size, coors_range, grid);
  const int64_t Dx = grid[0], Dy = grid[1], Dz = grid[2];
  int* map = coor_to_voxelidx;
  if (!map) {
    map = (int*)malloc(sizeof(int) * (size_t)(Dx * Dy * Dz));
    if (!map) return -1;
    memset(map, 0xff, sizeof(int) * (size_t)(Dx * Dy * Dz));
  }
  int voxel_num = 0;
  for (int i = 0; i < N; ++i) {                               /* :33 */
    int coor[3];
    int failed = 0;
    for (int j = 0; j < 3; ++j) {                             /* :35-40 */
      volatile float d = points[(size_t)i * F + j] - coors_range[j];
      volatile float q = d / voxel_size[j];
      float c = floorf(q);
      if (c < 0.0f || c >= (float)grid[j]) { failed = 1; break; }
      coor[2 - j] = (int)c;                                   /* reversed: (z,y,x) */
    }
    if (failed) continue;
    int64_t lin = ((int64_t)coor[0] * Dy + coor[1]) * Dx + coor[2];
    int voxelidx = map[lin];                                  /* :43 */
    if (voxelidx == -1) {                                     /* :44-50 */
      voxelidx = voxel_num;
      if (voxel_num >= max_voxels) continue;                  /* later-appearing voxels are dropped */
      voxel_num += 1;
      map[lin] = voxelidx;
      coors[voxelidx * 3 + 0] = coor[0];
      coors[voxelidx * 3 + 1] = coor[1];
      coors[voxelidx * 3 + 2] = coor[2];
    }
    int num = num_points[voxelidx];                           /* :51-54 */
    if (num < max_points) {
      memcpy(voxels + ((size_t)voxelidx * max_points + num) * F, points + (size_t)i * F,
             sizeof(float) * F);
      num_points[voxelidx] = num + 1;
    }
  }
  if (!coor_to_voxelidx) free(map);
  return voxel_num;
}

/* voxel_encoder.py:17-24 -- sum over the padded point slots / num_points */
void orc_voxel_mean(const float* voxels, const int* num_points, int M, int P, int F, int C,
                    float* out) {
  for (int v = 0; v < M; ++v)
    for (int c = 0; c < C; ++c) {
      float s = 0.f;
      for (int p = 0; p < P; ++p) s += voxels[((size_t)v * P + p) * F + c];
      out[(size_t)v * C + c] = s / (float)num_points[v];
    }
}

/* ------------------------------------------------------------------------------------ */
/* coordinate hash (CPU): 64-bit linear voxel index -> row                                */
/* ------------------------------------------------------------------------------------ */
typedef struct {
  uint64_t* keys;
  int* vals;
  uint64_t mask;
} orc_hash;

static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

static int hash_init(orc_hash* h, int64_t n) {
  uint64_t cap = 16;
  while (cap < (uint64_t)(2 * n + 1)) cap <<= 1;
  h->keys = (uint64_t*)malloc(cap * sizeof(uint64_t));
  h->vals = (int*)malloc(cap * sizeof(int));
  if (!h->keys || !h->vals) return -1;
  memset(h->keys, 0xff, cap * sizeof(uint64_t));
  h->mask = cap - 1;
  return 0;
}
static void hash_free(orc_hash* h) { free(h->keys); free(h->vals); }
static void hash_put(orc_hash* h, uint64_t key, int val) {
  uint64_t s = mix64(key) & h->mask;
  while (h->keys[s] != ~0ULL && h->keys[s] != key) s = (s + 1) & h->mask;
  h->keys[s] = key; h->vals[s] = val;
}
static int hash_get(const orc_hash* h, uint64_t key) {
  uint64_t s = mix64(key) & h->mask;
  while (h->keys[s] != ~0ULL) {
    if (h->keys[s] == key) return h->vals[s];
    s = (s + 1) & h->mask;
  }
  return -1;
}

static inline uint64_t lin4(int b, int z, int y, int x, const int* shape) {
  return (((uint64_t)b * shape[0] + z) * shape[1] + y) * shape[2] + x;
}

/* ------------------------------------------------------------------------------------ */
/* rulebooks -- spconv v1.x semantics; tables are k-major [K][stride]  */
/* ------------------------------------------------------------------------------------ */

/*
 * SubMConv3d: output set == input set, same row order; neighbour of row i under kernel
 * offset k=(kz,ky,kx) (row-major) is the active voxel at x_i + (k - ksize/2) * dilation.
 * tbl[k*stride + i] = input row or -1.  Returns the number of (in,out) pairs.
 */
int64_t orc_rulebook_subm(const int* coors, int N, const int* shape, const int* ksize,
                          const int* dilation, int* tbl, int tbl_stride) {
  orc_hash h;
  if (hash_init(&h, N)) return -1;
  for (int i = 0; i < N; ++i) {
    const int* c = coors + (size_t)i * 4;
    hash_put(&h, lin4(c[0], c[1], c[2], c[3], shape), i);
  }
  int64_t pairs = 0;
#pragma omp parallel for reduction(+ : pairs) schedule(static)
  for (int i = 0; i < N; ++i) {
    const int* c = coors + (size_t)i * 4;
    int k = 0;
    for (int kz = 0; kz < ksize[0]; ++kz)
      for (int ky = 0; ky < ksize[1]; ++ky)
        for (int kx = 0; kx < ksize[2]; ++kx, ++k) {
          int z = c[1] + (kz - ksize[0] / 2) * dilation[0];
          int y = c[2] + (ky - ksize[1] / 2) * dilation[1];
          int x = c[3] + (kx - ksize[2] / 2) * dilation[2];
          int j = -1;
          if (z >= 0 && z < shape[0] && y >= 0 && y < shape[1] && x >= 0 && x < shape[2])
            j = hash_get(&h, lin4(c[0], z, y, x, shape));
          tbl[(size_t)k * tbl_stride + i] = j;
          pairs += (j >= 0);
        }
  }
  hash_free(&h);
  return pairs;
}

void orc_conv_out_shape(const int* shape_in, const int* ksize, const int* stride, const int* pad,
                        const int* dilation, int* shape_out) {
  for (int a = 0; a < 3; ++a)
    shape_out[a] = (shape_in[a] + 2 * pad[a] - dilation[a] * (ksize[a] - 1) - 1) / stride[a] + 1;
}

static int cmp_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return (x > y) - (x < y);
}

/*
 * SparseConv3d (regular, strided): output site o is active iff some active input x and kernel
 * offset k satisfy x = o*s - p + k*d.  Canonical output order = ascending flattened
 * (b,z,y,x) index.  out_coors [cap,4]; tbl[k*tbl_stride + o] = input row or -1.
 * *pairs_out receives the pair count.  Returns n_out, or -2 if cap is too small.
 */
int orc_rulebook_sparse(const int* coors, int N, const int* shape_in, const int* ksize,
                        const int* stride, const int* pad, const int* dilation, int* out_coors,
                        int cap, int* tbl, int tbl_stride, int64_t* pairs_out) {
  int so[3];
  orc_conv_out_shape(shape_in, ksize, stride, pad, dilation, so);
  const int K = ksize[0] * ksize[1] * ksize[2];
  uint64_t* cand = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)N * K + 8);
  if (!cand) return -1;
  size_t nc = 0;
  for (int i = 0; i < N; ++i) {
    const int* c = coors + (size_t)i * 4;
    for (int kz = 0; kz < ksize[0]; ++kz) {
      int tz = c[1] + pad[0] - kz * dilation[0];
      if (tz < 0 || tz % stride[0]) continue;
      int oz = tz / stride[0];
      if (oz >= so[0]) continue;
      for (int ky = 0; ky < ksize[1]; ++ky) {
        int ty = c[2] + pad[1] - ky * dilation[1];
        if (ty < 0 || ty % stride[1]) continue;
        int oy = ty / stride[1];
        if (oy >= so[1]) continue;
        for (int kx = 0; kx < ksize[2]; ++kx) {
          int tx = c[3] + pad[2] - kx * dilation[2];
          if (tx < 0 || tx % stride[2]) continue;
          int ox = tx / stride[2];
          if (ox >= so[2]) continue;
          cand[nc++] = lin4(c[0], oz, oy, ox, so);
        }
      }
    }
  }
  qsort(cand, nc, sizeof(uint64_t), cmp_u64);
  size_t n_out = 0;
  for (size_t i = 0; i < nc; ++i)
    if (i == 0 || cand[i] != cand[i - 1]) cand[n_out++] = cand[i];
  if ((int64_t)n_out > cap) { free(cand); return -2; }
  for (size_t o = 0; o < n_out; ++o) {
    uint64_t l = cand[o];
    int x = (int)(l % so[2]); l /= so[2];
    int y = (int)(l % so[1]); l /= so[1];
    int z = (int)(l % so[0]); l /= so[0];
    out_coors[o * 4 + 0] = (int)l; out_coors[o * 4 + 1] = z;
    out_coors[o * 4 + 2] = y;      out_coors[o * 4 + 3] = x;
  }
  free(cand);

  orc_hash h;
  if (hash_init(&h, N)) return -1;
  for (int i = 0; i < N; ++i) {
    const int* c = coors + (size_t)i * 4;
    hash_put(&h, lin4(c[0], c[1], c[2], c[3], shape_in), i);
  }
  int64_t pairs = 0;
#pragma omp parallel for reduction(+ : pairs) schedule(static)
  for (int64_t o = 0; o < (int64_t)n_out; ++o) {
    const int* c = out_coors + (size_t)o * 4;
    int k = 0;
    for (int kz = 0; kz < ksize[0]; ++kz)
      for (int ky = 0; ky < ksize[1]; ++ky)
        for (int kx = 0; kx < ksize[2]; ++kx, ++k) {
          int z = c[1] * stride[0] - pad[0] + kz * dilation[0];
          int y = c[2] * stride[1] - pad[1] + ky * dilation[1];
          int x = c[3] * stride[2] - pad[2] + kx * dilation[2];
          int j = -1;
          if (z >= 0 && z < shape_in[0] && y >= 0 && y < shape_in[1] && x >= 0 && x < shape_in[2])
            j = hash_get(&h, lin4(c[0], z, y, x, shape_in));
          tbl[(size_t)k * tbl_stride + o] = j;
          pairs += (j >= 0);
        }
  }
  hash_free(&h);
  if (pairs_out) *pairs_out = pairs;
  return (int)n_out;
}

/* ------------------------------------------------------------------------------------ */
/* sparse convolution forward + BN/residual/ReLU + densify                                */
/* ------------------------------------------------------------------------------------ */

/*
 * out[o, :] = sum_k in[tbl[k][o], :] @ W[k]   with W [K, Cin, Cout] (= spconv's
 * [kD,kH,kW,Cin,Cout] flattened; cross-correlation, no flip).  fp32 products; the running
 * sum is fp32 (wide=0, what spconv's fp32 SGEMM + scatter-add does up to ordering) or
 * fp64 (wide=1, used as the tighter truth when judging reduced-precision kernels).
 */
void orc_spconv_fwd(const float* in, const float* W, const int* tbl, int tbl_stride, int N_out,
                    int Cin, int Cout, int K, float* out, int wide) {
#pragma omp parallel
  {
    double* accd = (double*)malloc(sizeof(double) * Cout);
    float* accf = (float*)malloc(sizeof(float) * Cout);
#pragma omp for schedule(dynamic, 256)
    for (int o = 0; o < N_out; ++o) {
      for (int c = 0; c < Cout; ++c) { accd[c] = 0.0; accf[c] = 0.f; }
      for (int k = 0; k < K; ++k) {
        int j = tbl[(size_t)k * tbl_stride + o];
        if (j < 0) continue;
        const float* x = in + (size_t)j * Cin;
        const float* w = W + (size_t)k * Cin * Cout;
        if (wide) {
          for (int ci = 0; ci < Cin; ++ci) {
            double xv = x[ci];
            const float* wr = w + (size_t)ci * Cout;
            for (int c = 0; c < Cout; ++c) accd[c] += xv * (double)wr[c];
          }
        } else {
          for (int ci = 0; ci < Cin; ++ci) {
            float xv = x[ci];
            const float* wr = w + (size_t)ci * Cout;
            for (int c = 0; c < Cout; ++c) accf[c] += xv * wr[c];
          }
        }
      }
      float* y = out + (size_t)o * Cout;
      if (wide) for (int c = 0; c < Cout; ++c) y[c] = (float)accd[c];
      else      for (int c = 0; c < Cout; ++c) y[c] = accf[c];
    }
    free(accd); free(accf);
  }
}

/* y = x*scale + shift (+ residual), optional ReLU; in place.  scn.py:69-85 */
void orc_bn_act(float* x, int N, int C, const float* scale, const float* shift,
                const float* residual, int relu) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; ++i)
    for (int c = 0; c < C; ++c) {
      float v = x[(size_t)i * C + c] * scale[c] + shift[c];
      if (residual) v += residual[(size_t)i * C + c];
      if (relu && v < 0.f) v = 0.f;
      x[(size_t)i * C + c] = v;
    }
}

/* dense() + view: bev[b, c*D + d, y, x] = feat[row(b,d,y,x), c]; bev caller-zeroed.  scn.py:173-176 */
void orc_dense_bev(const float* feat, const int* coors, int N, int C, int D, int H, int W,
                   float* bev) {
  for (int i = 0; i < N; ++i) {
    const int* c = coors + (size_t)i * 4;
    for (int ch = 0; ch < C; ++ch)
      bev[((((size_t)c[0] * C + ch) * D + c[1]) * H + c[2]) * W + c[3]] = feat[(size_t)i * C + ch];
  }
}

/* ==========================================================================================
 * Rotated-BEV IoU and greedy NMS  (CenterHead.predict -> rotate_nms_pcdet -> nms_gpu).
 * Restates det3d/ops/iou3d_nms/src/iou3d_nms_kernel.cu:
 *   cross / check_rect_cross (:36-50), check_in_box2d (:52-62, MARGIN 1e-2), intersection (:64-93,
 *   EPS 1e-8), rotate_around_center (:95-99), point_cmp (:101-103), box_overlap (:105-226),
 *   iou_bev (:228-235), and the host sweep of iou3d_nms.cpp:90-136 (keep box i iff no kept box
 *   j < i has IoU(j, i) > thresh; boxes arrive sorted by descending score).
 * PINNED: checked against the reference's own CPU twin (det3d/ops/iou3d_nms/src/iou3d_cpu.cpp,
 * compiled unmodified into oracle/_ref by oracle/Makefile) and against tests/golden/iou_bev_*.npz
 * generated from it.  All arithmetic in fp32, no contraction (-ffp-contract=off).
 * ========================================================================================== */
typedef struct { float x, y; } orc_pt;

static inline float orc_cross2(orc_pt a, orc_pt b) { return a.x * b.y - a.y * b.x; }
static inline float orc_cross3(orc_pt p1, orc_pt p2, orc_pt p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
static inline float orc_minf(float a, float b) { return a > b ? b : a; }
static inline float orc_maxf(float a, float b) { return a > b ? a : b; }

static int orc_rect_cross(orc_pt p1, orc_pt p2, orc_pt q1, orc_pt q2) {
  return orc_minf(p1.x, p2.x) <= orc_maxf(q1.x, q2.x) && orc_minf(q1.x, q2.x) <= orc_maxf(p1.x, p2.x) &&
         orc_minf(p1.y, p2.y) <= orc_maxf(q1.y, q2.y) && orc_minf(q1.y, q2.y) <= orc_maxf(p1.y, p2.y);
}

static int orc_in_box2d(const float* box, orc_pt p) {
  const float MARGIN = 1e-2f;
  const float ac = cosf(-box[6]), as = sinf(-box[6]);
  const float rx = (p.x - box[0]) * ac + (p.y - box[1]) * (-as);
  const float ry = (p.x - box[0]) * as + (p.y - box[1]) * ac;
  return fabsf(rx) < box[3] / 2 + MARGIN && fabsf(ry) < box[4] / 2 + MARGIN;
}

static int orc_intersection(orc_pt p1, orc_pt p0, orc_pt q1, orc_pt q0, orc_pt* ans) {
  const float EPS = 1e-8f;
  if (!orc_rect_cross(p0, p1, q0, q1)) return 0;
  const float s1 = orc_cross3(q0, p1, p0), s2 = orc_cross3(p1, q1, p0);
  const float s3 = orc_cross3(p0, q1, q0), s4 = orc_cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  const float s5 = orc_cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > EPS) {
    ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const float D = a0 * b1 - a1 * b0;
    ans->x = (b0 * c1 - b1 * c0) / D;
    ans->y = (a1 * c0 - a0 * c1) / D;
  }
  return 1;
}

static void orc_corners(const float* box, orc_pt* c) {
  const float hx = box[3] / 2, hy = box[4] / 2;
  const float x1 = box[0] - hx, y1 = box[1] - hy, x2 = box[0] + hx, y2 = box[1] + hy;
  const float ac = cosf(box[6]), as = sinf(box[6]);
  const float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
  for (int k = 0; k < 4; ++k) {
    c[k].x = (px[k] - box[0]) * ac + (py[k] - box[1]) * (-as) + box[0];
    c[k].y = (px[k] - box[0]) * as + (py[k] - box[1]) * ac + box[1];
  }
  c[4] = c[0];
}

float orc_box_overlap(const float* a, const float* b) {
  orc_pt ca[5], cb[5], pts[16], ctr = {0.f, 0.f};
  orc_corners(a, ca);
  orc_corners(b, cb);
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (orc_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], &pts[cnt])) {
        ctr.x = ctr.x + pts[cnt].x; ctr.y = ctr.y + pts[cnt].y;
        ++cnt;
      }
  for (int k = 0; k < 4; ++k) {
    if (orc_in_box2d(a, cb[k])) { ctr.x = ctr.x + cb[k].x; ctr.y = ctr.y + cb[k].y; pts[cnt++] = cb[k]; }
    if (orc_in_box2d(b, ca[k])) { ctr.x = ctr.x + ca[k].x; ctr.y = ctr.y + ca[k].y; pts[cnt++] = ca[k]; }
  }
  ctr.x /= cnt; ctr.y /= cnt;                          /* cnt == 0: NaN centre, empty loops below */
  for (int j = 0; j < cnt - 1; ++j)                    /* bubble sort by angle about the centroid */
    for (int i = 0; i < cnt - j - 1; ++i)
      if (atan2f(pts[i].y - ctr.y, pts[i].x - ctr.x) > atan2f(pts[i + 1].y - ctr.y, pts[i + 1].x - ctr.x)) {
        orc_pt t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
      }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    orc_pt u = {pts[k].x - pts[0].x, pts[k].y - pts[0].y}, v = {pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y};
    area += orc_cross2(u, v);
  }
  return (float)(fabs(area) / 2.0);
}

float orc_iou_bev(const float* a, const float* b) {
  const float sa = a[3] * a[4], sb = b[3] * b[4];
  const float ov = orc_box_overlap(a, b);
  return ov / fmaxf(sa + sb - ov, 1e-8f);
}

/* ious [na, nb] */
void orc_iou_bev_matrix(const float* a, int na, const float* b, int nb, float* ious) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) ious[(size_t)i * nb + j] = orc_iou_bev(a + 7 * i, b + 7 * j);
}

/* nms_gpu (iou3d_nms.cpp:90-136): boxes [n,7] sorted by descending score -> keep indices; returns the count.
 * min_margin (nullable): min |IoU - thresh| over the pairs the sweep evaluated (test diagnostics). */
int orc_nms_sorted(const float* boxes, int n, float thresh, int64_t* keep, float* min_margin) {
  unsigned char* removed = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  int nk = 0;
  float margin = 1e30f;
  for (int i = 0; i < n; ++i) {
    if (removed[i]) continue;
    keep[nk++] = i;
    for (int j = i + 1; j < n; ++j) {
      const float iou = orc_iou_bev(boxes + 7 * i, boxes + 7 * j);
      const float m = fabsf(iou - thresh);
      if (m < margin) margin = m;
      if (iou > thresh) removed[j] = 1;
    }
  }
  free(removed);
  if (min_margin) *min_margin = margin;
  return nk;
}
