// detect.cu -- CenterHead.predict on the device: decode, candidate selection, rotated-BEV NMS, gather.
//
// Replaces (reference, per sample and per task): ~25 eager torch kernels of center_head.py:342-419 (permute,
// sigmoid / exp / atan2, meshgrid decode), the boolean-mask indexing + torch.sort of post_processing (:450-495) and
// rotate_nms_pcdet (box_torch_ops.py:449-470), and iou3d_nms_cuda.nms_gpu (iou3d_nms.cpp:90-136), which
// cudaMallocs a mask, copies it to the host and sweeps it on the CPU.  Here the whole chain is four launches for
// the whole batch and never leaves the device.
#include "common.cuh"

namespace s2d {

constexpr int kNmsMaxBoxes = 4096;          // nms_pre_max_size of every Waymo config; 64 mask words per row
constexpr int kNmsTile = 64;

// ---------------------------------------------------------------------------------------------
// decode  (center_head.py:342-401 + the masks of :456-465)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) centerhead_decode_kernel(const __grid_constant__ s2d_decode_params P,
                                                                float* __restrict__ boxes, float* __restrict__ scores,
                                                                int* __restrict__ labels,
                                                                unsigned long long* __restrict__ keys) {
  const int HW = P.H * P.W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)P.B * HW) return;
  const int cell = (int)(i % HW);
  const int row = cell / P.W, col = cell - row * P.W;
  // scores, labels = torch.max(sigmoid(hm), dim=-1): first maximum wins
  const float* hm = P.hm + i * P.ld_hm;
  float best = 0.f;
  int lab = 0;
  for (int c = 0; c < P.num_cls; ++c) {
    const float s = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-hm[c])));
    if (c == 0 || s > best) { best = s; lab = c; }
  }
  const float* reg = P.reg + i * P.ld_reg;
  const float* dim = P.dim + i * P.ld_dim;
  const float* rot = P.rot + i * P.ld_rot;
  // xs = (col + reg_x) * out_size_factor * voxel_x + pc_x, every step rounded to fp32 like the eager torch ops
  const float x = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn((float)col, reg[0]), P.out_size_factor), P.voxel_x), P.pc_x);
  const float y = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn((float)row, reg[1]), P.out_size_factor), P.voxel_y), P.pc_y);
  const float z = P.height[i * P.ld_height];
  float* b = boxes + i * 7;
  b[0] = x; b[1] = y; b[2] = z;
  b[3] = expf(dim[0]); b[4] = expf(dim[1]); b[5] = expf(dim[2]);
  b[6] = atan2f(rot[0], rot[1]);                                   // atan2(sin, cos)
  scores[i] = best;
  labels[i] = lab;
  const bool ok = best > P.score_threshold && x >= P.range[0] && y >= P.range[1] && z >= P.range[2] &&
                  x <= P.range[3] && y <= P.range[4] && z <= P.range[5];
  // sort key: score descending, then cell ascending (sigmoid > 0, so the float bits order like the values)
  keys[i] = ok ? ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)cell)
               : 0ull;
}

// ---------------------------------------------------------------------------------------------
// top-K selection + sort, one CTA per sample: order[b][0..count) = cells by descending key
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) topk_sort_kernel(const unsigned long long* __restrict__ keys, int cells,
                                                         int pre_max, int* __restrict__ order,
                                                         int* __restrict__ counts) {
  __shared__ unsigned long long s_keys[kNmsMaxBoxes];
  __shared__ int s_cnt;
  __shared__ int s_red[32];
  const unsigned long long* k = keys + (size_t)blockIdx.x * cells;
  const int tid = threadIdx.x;

  auto block_count_ge = [&](unsigned long long thr) {
    int c = 0;
    for (int i0 = 0; i0 < cells; i0 += 8 * 1024) {               // eight independent loads in flight per thread: the loop is
      unsigned long long v[8];                                   // latency bound (one CTA per sample), not bandwidth bound
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * 1024 + tid;
        v[u] = i < cells ? __ldg(k + i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) c += (v[u] >= thr && v[u] != 0ull) ? 1 : 0;
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = c;
    __syncthreads();
    int tot = 0;
    for (int w = 0; w < 32; ++w) tot += s_red[w];
    __syncthreads();                                     // nobody reuses shared memory before every thread has its total
    return tot;
  };

  // the K-th largest key (keys are unique: the cell index is part of the key); 1 = "every valid key".
  // Radix select, one byte per pass from the top: histogram of the byte among the keys that match the prefix found so far
  // (warp-aggregated shared-memory atomics: the high bytes of a score take few distinct values), then walk the bins from 255
  // down to the one that contains the pre_max-th largest key.  8 passes over the keys (the first version tested one BIT per
  // pass: 65 passes, 0.6 ms for the 219 k cells of a pillar map).
  unsigned long long thr = 1ull;
  const int n_valid = block_count_ge(1ull);
  if (n_valid > pre_max) {
    __shared__ int s_hist[256];
    __shared__ int s_pick[2];
    unsigned long long prefix = 0ull, mask = 0ull;
    int remaining = pre_max;
    for (int shift = 56; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += 1024) s_hist[i] = 0;
      __syncthreads();
      for (int i0 = 0; i0 < cells; i0 += 8 * 1024) {
        unsigned long long v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 1024 + tid;
          v[u] = i < cells ? __ldg(k + i) : 0ull;                // 0 never matches a non-zero prefix; with prefix 0 it only
        }                                                        // lands in bins below the wanted one (n_valid > pre_max)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool in = i0 + u * 1024 + tid < cells && (v[u] & mask) == prefix;
          const unsigned digit = (unsigned)(v[u] >> shift) & 255u;
          const unsigned act = __ballot_sync(0xffffffffu, in);
          if (in) {
            const unsigned peers = __match_any_sync(act, digit);
            if ((tid & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[digit], __popc(peers));
          }
        }
      }
      __syncthreads();
      if (tid == 0) {
        int cum = 0, b = 255;
        for (; b > 0; --b) {
          if (cum + s_hist[b] >= remaining) break;
          cum += s_hist[b];
        }
        s_pick[0] = b;
        s_pick[1] = remaining - cum;                       // rank of the wanted key inside bin b
      }
      __syncthreads();
      prefix |= (unsigned long long)s_pick[0] << shift;
      mask |= 255ull << shift;
      remaining = s_pick[1];
      __syncthreads();
    }
    thr = prefix;
  }
  const int n_sel = n_valid > pre_max ? pre_max : n_valid;
  for (int i = tid; i < kNmsMaxBoxes; i += 1024) s_keys[i] = 0ull;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  for (int i0 = 0; i0 < cells; i0 += 8 * 1024) {
    unsigned long long v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * 1024 + tid;
      v[u] = i < cells ? __ldg(k + i) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (v[u] >= thr && v[u] != 0ull) s_keys[atomicAdd(&s_cnt, 1)] = v[u];
  }
  __syncthreads();
  // bitonic sort, descending (zeros sink to the end)
  for (int size = 2; size <= kNmsMaxBoxes; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < kNmsMaxBoxes / 2; t += 1024) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = s_keys[lo], b = s_keys[hi];
        if (desc ? (a < b) : (a > b)) { s_keys[lo] = b; s_keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  int* o = order + (size_t)blockIdx.x * pre_max;
  for (int i = tid; i < pre_max; i += 1024)
    o[i] = i < n_sel ? (int)(0xFFFFFFFFu - (unsigned)(s_keys[i] & 0xFFFFFFFFull)) : -1;
  if (tid == 0) counts[blockIdx.x] = n_sel;
}

// ---------------------------------------------------------------------------------------------
// rotated-BEV overlap, the geometry of iou3d_nms_kernel.cu:36-235 restated.  Products and sums use the
// explicit round-to-nearest intrinsics so that no FMA contraction changes a sign test; the CPU oracle
// (and the reference's CPU twin iou3d_cpu.cpp) evaluate the same expressions without contraction.
// ---------------------------------------------------------------------------------------------
struct P2 { float x, y; };
__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float cross2(P2 a, P2 b) { return fs(fm(a.x, b.y), fm(a.y, b.x)); }
__device__ __forceinline__ float cross3(P2 p1, P2 p2, P2 p0) {
  return fs(fm(fs(p1.x, p0.x), fs(p2.y, p0.y)), fm(fs(p2.x, p0.x), fs(p1.y, p0.y)));
}
__device__ __forceinline__ bool rect_cross(P2 p1, P2 p2, P2 q1, P2 q2) {
  return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
         fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}
__device__ __forceinline__ bool in_box2d(const float* box, P2 p) {
  const float ac = cosf(-box[6]), as = sinf(-box[6]);
  const float dx = fs(p.x, box[0]), dy = fs(p.y, box[1]);
  const float rx = fa(fm(dx, ac), fm(dy, -as));
  const float ry = fa(fm(dx, as), fm(dy, ac));
  return fabsf(rx) < fa(fd(box[3], 2.f), 1e-2f) && fabsf(ry) < fa(fd(box[4], 2.f), 1e-2f);
}
__device__ __forceinline__ bool seg_intersection(P2 p1, P2 p0, P2 q1, P2 q0, P2& ans) {
  if (!rect_cross(p0, p1, q0, q1)) return false;
  const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
  if (!(fm(s1, s2) > 0.f && fm(s3, s4) > 0.f)) return false;
  const float s5 = cross3(q1, p1, p0);
  if (fabsf(fs(s5, s1)) > 1e-8f) {
    ans.x = fd(fs(fm(s5, q0.x), fm(s1, q1.x)), fs(s5, s1));
    ans.y = fd(fs(fm(s5, q0.y), fm(s1, q1.y)), fs(s5, s1));
  } else {
    const float a0 = fs(p0.y, p1.y), b0 = fs(p1.x, p0.x), c0 = fs(fm(p0.x, p1.y), fm(p1.x, p0.y));
    const float a1 = fs(q0.y, q1.y), b1 = fs(q1.x, q0.x), c1 = fs(fm(q0.x, q1.y), fm(q1.x, q0.y));
    const float D = fs(fm(a0, b1), fm(a1, b0));
    ans.x = fd(fs(fm(b0, c1), fm(b1, c0)), D);
    ans.y = fd(fs(fm(a1, c0), fm(a0, c1)), D);
  }
  return true;
}
__device__ __forceinline__ void box_corners(const float* box, P2* c) {
  const float hx = fd(box[3], 2.f), hy = fd(box[4], 2.f);
  const float x1 = fs(box[0], hx), y1 = fs(box[1], hy), x2 = fa(box[0], hx), y2 = fa(box[1], hy);
  const float ac = cosf(box[6]), as = sinf(box[6]);
  const float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float dx = fs(px[k], box[0]), dy = fs(py[k], box[1]);
    c[k].x = fa(fa(fm(dx, ac), fm(dy, -as)), box[0]);
    c[k].y = fa(fa(fm(dx, as), fm(dy, ac)), box[1]);
  }
  c[4] = c[0];
}
__device__ float box_overlap_bev(const float* a, const float* b) {
  P2 ca[5], cb[5], pts[16];
  box_corners(a, ca);
  box_corners(b, cb);
  P2 ctr = {0.f, 0.f};
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      P2 q;
      if (seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], q)) {
        pts[cnt++] = q;
        ctr.x = fa(ctr.x, q.x); ctr.y = fa(ctr.y, q.y);
      }
    }
  for (int k = 0; k < 4; ++k) {
    if (in_box2d(a, cb[k])) { ctr.x = fa(ctr.x, cb[k].x); ctr.y = fa(ctr.y, cb[k].y); pts[cnt++] = cb[k]; }
    if (in_box2d(b, ca[k])) { ctr.x = fa(ctr.x, ca[k].x); ctr.y = fa(ctr.y, ca[k].y); pts[cnt++] = ca[k]; }
  }
  if (cnt < 3) return 0.f;                        // fewer than 3 vertices: the fan below is empty or degenerate-zero
  ctr.x = fd(ctr.x, (float)cnt); ctr.y = fd(ctr.y, (float)cnt);
  float ang[16];
  for (int i = 0; i < cnt; ++i) ang[i] = atan2f(fs(pts[i].y, ctr.y), fs(pts[i].x, ctr.x));
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (ang[i] > ang[i + 1]) {
        const P2 t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
        const float s = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = s;
      }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    const P2 u = {fs(pts[k].x, pts[0].x), fs(pts[k].y, pts[0].y)};
    const P2 v = {fs(pts[k + 1].x, pts[0].x), fs(pts[k + 1].y, pts[0].y)};
    area = fa(area, cross2(u, v));
  }
  return fabsf(area) * 0.5f;
}
__device__ __forceinline__ float iou_bev_dev(const float* a, const float* b) {
  const float sa = fm(a[3], a[4]), sb = fm(b[3], b[4]);
  const float ov = box_overlap_bev(a, b);
  return fd(ov, fmaxf(fs(fa(sa, sb), ov), 1e-8f));
}

// mask[b][i][cb] bit j set <=> IoU(box i, box 64*cb + j) > thresh, for 64*cb + j > i  (iou3d_nms_kernel.cu:267-311)
// 256 threads per 64 x 64 tile: thread = (row r, column quarter q); the IoU of a pair is ~1.5 k divergent instructions, so
// the kernel is latency bound and four times the threads per tile is worth more than anything else (0.64 -> ms at batch 8).
__global__ void __launch_bounds__(4 * kNmsTile) nms_mask_kernel(const float* __restrict__ boxes, long long sample_stride,
                                                                const int* __restrict__ order, int order_stride,
                                                                const int* __restrict__ counts, int n_fixed,
                                                                float thresh, int words,
                                                                unsigned long long* __restrict__ mask, int blk_lo,
                                                                const int* __restrict__ done) {
  const int b = blockIdx.z, rb = blockIdx.y, cbk = blockIdx.x;
  const int n = counts ? counts[b] : n_fixed;
  if (cbk < rb || rb * kNmsTile >= n || cbk * kNmsTile >= n) return;      // the sweep reads only words >= the row block
  if (cbk < blk_lo || (done && done[b])) return;                          // an earlier stage computed it / the sweep has finished
  const float* bx = boxes + (size_t)b * sample_stride * 7;
  const int* ord = order ? order + (size_t)b * order_stride : nullptr;
  __shared__ float s_box[kNmsTile * 7];
  __shared__ unsigned long long s_bits[3][kNmsTile];
  const int col_size = min(n - cbk * kNmsTile, kNmsTile), row_size = min(n - rb * kNmsTile, kNmsTile);
  const int t = threadIdx.x & (kNmsTile - 1), q = threadIdx.x / kNmsTile;
  if (q == 0 && t < col_size) {
    const int j = cbk * kNmsTile + t;
    const float* src = bx + (size_t)(ord ? ord[j] : j) * 7;
#pragma unroll
    for (int c = 0; c < 7; ++c) s_box[t * 7 + c] = src[c];
  }
  __syncthreads();
  unsigned long long bits = 0ull;
  if (t < row_size) {
    const int i = rb * kNmsTile + t;
    const float* src = bx + (size_t)(ord ? ord[i] : i) * 7;
    float cur[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) cur[c] = src[c];
    // Exact early-out: when the circumscribed circles (plus the 1e-2 corner margin of check_in_box2d) are disjoint,
    // no edges cross and no corner passes the inside test, so the reference overlap is exactly 0.
    const float rad = 0.5f * sqrtf(cur[3] * cur[3] + cur[4] * cur[4]) + 0.05f;
    const int j_lo = max(16 * q, (rb == cbk) ? t + 1 : 0), j_hi = min(16 * q + 16, col_size);
    for (int j = j_lo; j < j_hi; ++j) {
      const float* o = s_box + j * 7;
      const float dx = o[0] - cur[0], dy = o[1] - cur[1];
      const float reach = rad + 0.5f * sqrtf(o[3] * o[3] + o[4] * o[4]);
      if (dx * dx + dy * dy > reach * reach) continue;
      if (iou_bev_dev(cur, o) > thresh) bits |= 1ull << j;
    }
  }
  if (q > 0) s_bits[q - 1][t] = bits;
  __syncthreads();
  if (q == 0 && t < row_size)
    mask[((size_t)b * kNmsMaxBoxes + rb * kNmsTile + t) * words + cbk] = bits | s_bits[0][t] | s_bits[1][t] | s_bits[2][t];
}

// Greedy sweep of iou3d_nms.cpp:118-131 by one warp per sample (the running `remv` mask lives in two registers per
// lane), stopped after post_max kept boxes (rotate_nms_pcdet slices [:post_max_size]); then the kept detections
// are gathered.  keep[b][k] = position in the sorted order.
//
// Staged form (s2d_centerhead_select): the reference sweep stops at the post_max-th kept box, so only the pairs (i, j) with
// i < j <= i_stop are ever consulted -- with 4096 candidates and 500 kept that is a few percent of the IoU triangle.  Stage s
// computes the mask blocks below `limit_s` candidates that earlier stages have not, then sweeps the first limit_s candidates
// from scratch; when the sweep ends inside the stage (post_max kept, or all n candidates seen) it raises done[b] and the
// later stages' blocks exit at once.  The last stage has limit = pre_max, i.e. it is the unstaged algorithm: keep sets are
// identical by construction.
__global__ void __launch_bounds__(32) nms_sweep_kernel(const unsigned long long* __restrict__ mask, int words,
                                                       const int* __restrict__ counts, int n_fixed, int post_max,
                                                       int* __restrict__ keep, int* __restrict__ n_keep, int limit,
                                                       int* __restrict__ done) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int n_all = counts ? counts[b] : n_fixed;
  if (done && done[b]) return;
  const int n = min(n_all, limit);
  const int live_words = min(words, (n + kNmsTile - 1) / kNmsTile);   // words the stages so far have computed
  const unsigned long long* m = mask + (size_t)b * kNmsMaxBoxes * words;
  int* kp = keep + (size_t)b * post_max;
  unsigned long long r0 = 0ull, r1 = 0ull;      // remv words lane and lane + 32
  int nk = 0;
  // The sweep is one dependent chain (a row is needed as soon as its box is kept), so the mask rows of the next kAhead
  // candidates are loaded speculatively: a kept box then finds its row in registers instead of paying an L2 round trip.
  constexpr int kAhead = 4;
  unsigned long long p0[kAhead], p1[kAhead];
  auto load_row = [&](int i, unsigned long long& a, unsigned long long& c) {
    a = c = 0ull;
    if (i < n) {
      const unsigned long long* row = m + (size_t)i * words;
      const int nb = i >> 6;
      if (lane >= nb && lane < live_words) a = row[lane];
      if (lane + 32 >= nb && lane + 32 < live_words) c = row[lane + 32];
    }
  };
#pragma unroll
  for (int u = 0; u < kAhead; ++u) load_row(u, p0[u], p1[u]);
  for (int i = 0; i < n && nk < post_max; ++i) {
    const unsigned long long a = p0[0], c = p1[0];
#pragma unroll
    for (int u = 0; u + 1 < kAhead; ++u) { p0[u] = p0[u + 1]; p1[u] = p1[u + 1]; }
    load_row(i + kAhead, p0[kAhead - 1], p1[kAhead - 1]);
    const int nb = i >> 6;
    const unsigned long long w = __shfl_sync(0xffffffffu, nb < 32 ? r0 : r1, nb & 31);
    if ((w >> (i & 63)) & 1ull) continue;       // warp-uniform
    if (lane == 0) kp[nk] = i;
    ++nk;
    r0 |= a;
    r1 |= c;
  }
  if (lane == 0) {
    n_keep[b] = nk;
    if (done) done[b] = (nk >= post_max || n >= n_all) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(128) gather_detections_kernel(const float* __restrict__ boxes,
                                                                const float* __restrict__ scores,
                                                                const int* __restrict__ labels, int cells,
                                                                const int* __restrict__ order, int pre_max,
                                                                const int* __restrict__ keep,
                                                                const int* __restrict__ n_keep, int post_max,
                                                                float* __restrict__ out_boxes,
                                                                float* __restrict__ out_scores,
                                                                int* __restrict__ out_labels,
                                                                int* __restrict__ out_cells) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= post_max) return;
  const size_t o = (size_t)b * post_max + k;
  if (k < n_keep[b]) {
    const int cell = order[(size_t)b * pre_max + keep[o]];
    const size_t src = (size_t)b * cells + cell;
#pragma unroll
    for (int q = 0; q < 7; ++q) out_boxes[o * 7 + q] = boxes[src * 7 + q];
    out_scores[o] = scores[src];
    out_labels[o] = labels[src];
    if (out_cells) out_cells[o] = cell;
  } else {
#pragma unroll
    for (int q = 0; q < 7; ++q) out_boxes[o * 7 + q] = 0.f;
    out_scores[o] = 0.f;
    out_labels[o] = -1;
    if (out_cells) out_cells[o] = -1;
  }
}

struct SelectWs {
  int* order; int* counts; int* keep; int* done; unsigned long long* mask; size_t total;
};
static SelectWs carve_select(void* ws, int batch, int pre_max, int post_max) {
  Carver c(ws);
  SelectWs w;
  w.order = c.take<int>((size_t)batch * pre_max);
  w.counts = c.take<int>(batch);
  w.keep = c.take<int>((size_t)batch * post_max);
  w.done = c.take<int>(batch);
  w.mask = c.take<unsigned long long>((size_t)batch * kNmsMaxBoxes * (kNmsMaxBoxes / kNmsTile));
  w.total = c.off;
  return w;
}

}  // namespace s2d

using namespace s2d;

extern "C" int s2d_centerhead_decode(const s2d_decode_params* params, float* boxes, float* scores, int* labels,
                                     unsigned long long* keys, void* stream) {
  S2D_REQUIRE(params && boxes && scores && labels && keys, "s2d_centerhead_decode: null argument");
  const s2d_decode_params& p = *params;
  S2D_REQUIRE(p.reg && p.height && p.dim && p.rot && p.hm, "s2d_centerhead_decode: null head map");
  S2D_REQUIRE(p.B >= 0 && p.H >= 1 && p.W >= 1 && p.num_cls >= 1, "s2d_centerhead_decode: bad sizes");
  const long long n = (long long)p.B * p.H * p.W;
  if (n == 0) return S2D_OK;
  centerhead_decode_kernel<<<div_up(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, boxes, scores, labels, keys);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" size_t s2d_centerhead_select_workspace_bytes(int batch, int pre_max, int post_max) {
  if (batch < 1 || pre_max < 1 || post_max < 1) return 0;
  return carve_select(nullptr, batch, pre_max, post_max).total;
}

extern "C" int s2d_centerhead_select(const unsigned long long* keys, const float* boxes, const float* scores,
                                     const int* labels, int batch, int cells, int pre_max, float iou_threshold,
                                     int post_max, float* out_boxes, float* out_scores, int* out_labels, int* out_cells,
                                     int* n_out, void* workspace, size_t workspace_bytes, void* stream) {
  S2D_REQUIRE(keys && boxes && scores && labels && out_boxes && out_scores && out_labels && n_out && workspace,
              "s2d_centerhead_select: null argument");
  S2D_REQUIRE(batch >= 1 && cells >= 1 && post_max >= 1, "s2d_centerhead_select: bad sizes");
  S2D_REQUIRE(pre_max >= 1 && pre_max <= kNmsMaxBoxes, "s2d_centerhead_select: nms_pre_max_size %d outside [1,%d]",
              pre_max, kNmsMaxBoxes);
  const SelectWs w = carve_select(workspace, batch, pre_max, post_max);
  if (workspace_bytes < w.total) {
    set_error("s2d_centerhead_select: workspace %zu < %zu bytes", workspace_bytes, w.total);
    return S2D_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int words = kNmsMaxBoxes / kNmsTile;
  topk_sort_kernel<<<batch, 1024, 0, st>>>(keys, cells, pre_max, w.order, w.counts);
  S2D_CUDA(cudaMemsetAsync(w.done, 0, sizeof(int) * batch, st));
  int launches = 2;
  int lo = 0;
  for (int limit : {post_max + post_max / 4 + kNmsTile, 3 * post_max + kNmsTile, pre_max}) {
    limit = min(limit, pre_max);
    const int blocks = div_up(limit, kNmsTile);
    if (blocks <= lo) continue;
    nms_mask_kernel<<<dim3(blocks, blocks, batch), 4 * kNmsTile, 0, st>>>(boxes, cells, w.order, pre_max, w.counts, 0,
                                                                      iou_threshold, words, w.mask, lo, w.done);
    nms_sweep_kernel<<<batch, 32, 0, st>>>(w.mask, words, w.counts, 0, post_max, w.keep, n_out, blocks * kNmsTile, w.done);
    lo = blocks;
    launches += 2;
  }
  gather_detections_kernel<<<dim3(div_up(post_max, 128), batch), 128, 0, st>>>(
      boxes, scores, labels, cells, w.order, pre_max, w.keep, n_out, post_max, out_boxes, out_scores, out_labels,
      out_cells);
  S2D_LAUNCH_CHECK();
  count_launches(launches);
  return S2D_OK;
}

extern "C" size_t s2d_nms_workspace_bytes(int n_boxes) {
  if (n_boxes < 0 || n_boxes > kNmsMaxBoxes) return 0;
  return (size_t)kNmsMaxBoxes * (kNmsMaxBoxes / kNmsTile) * sizeof(unsigned long long);
}

extern "C" int s2d_nms_sorted(const float* boxes, int n_boxes, float iou_threshold, int* keep, int* n_keep,
                              void* workspace, size_t workspace_bytes, void* stream) {
  S2D_REQUIRE(keep && n_keep, "s2d_nms_sorted: null argument");
  S2D_REQUIRE(n_boxes >= 0 && n_boxes <= kNmsMaxBoxes, "s2d_nms_sorted: %d boxes outside [0,%d]", n_boxes, kNmsMaxBoxes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_boxes == 0) {
    S2D_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int), st));
    return S2D_OK;
  }
  S2D_REQUIRE(boxes && workspace, "s2d_nms_sorted: null argument");
  if (workspace_bytes < s2d_nms_workspace_bytes(n_boxes)) {
    set_error("s2d_nms_sorted: workspace too small");
    return S2D_ERR_WORKSPACE;
  }
  const int words = kNmsMaxBoxes / kNmsTile;
  unsigned long long* mask = static_cast<unsigned long long*>(workspace);
  const int blocks = div_up(n_boxes, kNmsTile);
  nms_mask_kernel<<<dim3(blocks, blocks, 1), 4 * kNmsTile, 0, st>>>(boxes, 0, nullptr, 0, nullptr, n_boxes, iou_threshold,
                                                                words, mask, 0, nullptr);
  nms_sweep_kernel<<<1, 32, 0, st>>>(mask, words, nullptr, n_boxes, n_boxes, keep, n_keep, n_boxes, nullptr);
  S2D_LAUNCH_CHECK();
  count_launches(2);
  return S2D_OK;
}

extern "C" int s2d_iou_bev(const float* boxes_a, int n_a, const float* boxes_b, int n_b, float* ious, void* stream);
namespace s2d {
__global__ void __launch_bounds__(256) iou_bev_kernel(const float* __restrict__ a, int na, const float* __restrict__ b,
                                                      int nb, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)na * nb) return;
  out[i] = iou_bev_dev(a + (i / nb) * 7, b + (i % nb) * 7);
}
}  // namespace s2d
extern "C" int s2d_iou_bev(const float* boxes_a, int n_a, const float* boxes_b, int n_b, float* ious, void* stream) {
  S2D_REQUIRE(n_a >= 0 && n_b >= 0, "s2d_iou_bev: bad sizes");
  if ((long long)n_a * n_b == 0) return S2D_OK;
  S2D_REQUIRE(boxes_a && boxes_b && ious, "s2d_iou_bev: null argument");
  iou_bev_kernel<<<div_up((long long)n_a * n_b, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(boxes_a, n_a, boxes_b,
                                                                                              n_b, ious);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
