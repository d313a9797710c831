// voxelize.cu -- deterministic hash-based voxel builder for sm_100a.
//
// Replaces the serial numba loop of det3d/ops/point_cloud/point_cloud_ops.py:7-55 (wrapper
// :112-184) bit-exactly, for a whole batch of scenes in one pass.  The reference loop is order
// dependent; its result is restated as order statistics so it can run in parallel:
//
//   voxel id        = rank of the voxel's FIRST point among all first points of the scene
//                     (ids follow first appearance, :44-50)
//   max_voxels cap  = voxels of rank >= max_voxels are dropped with all their points (:46-47),
//                     voxels created earlier keep filling (:51-54)
//   slot p of voxel = the voxel's p-th smallest point index, p < max_points (:51-54)
//
// Stages (all HBM/L2-latency bound integer work, one thread per point unless noted):
//   1 vox_insert   fp32 (p - lo) / vs -> floor -> key; hash find-or-insert (atomicCAS);
//                  atomicMin of the point index into the slot's `first`
//   2 scan         exclusive scan of "is first point of its voxel" -> ranks
//   3 vox_offsets  per-scene counts capped at max_voxels -> output row offsets (1 thread)
//   4 vox_assign   first points write (b,z,y,x) and slot 0 of their output row
//   5 vox_kth      (max_points-1 launches) p-th smallest index per voxel by atomicMin
//   6 vox_gather   one thread per output row: copy the kept points, count, mean
#include <limits.h>

#include "common.cuh"

namespace s2d {

constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kInf = 0x7f7f7f7f;  // what memset(0x7f) leaves in an int

struct VoxParams {
  int off[S2D_MAX_BATCH + 1];
  float lo[3], vs[3];
  int grid[3];  // x, y, z
  int batch, n_points, F, max_points, max_voxels;
  unsigned int hash_mask;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

__device__ __forceinline__ int scene_of(const VoxParams& P, int i) {
  int b = 0;
  while (b + 1 < P.batch && i >= P.off[b + 1]) ++b;
  return b;
}

// Stage 1.  pslot[i] = hash slot of point i's voxel or -1 (outside the range).
__global__ void __launch_bounds__(256) vox_insert_kernel(const float* __restrict__ points,
                                                         const __grid_constant__ VoxParams P,
                                                         unsigned long long* __restrict__ keys,
                                                         int* __restrict__ first, int* __restrict__ pslot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_points) return;
  const float* p = points + (size_t)i * P.F;
  int c[3];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    // separately rounded fp32 subtract, divide, floor -- point_cloud_ops.py:36
    const float q = floorf(__fdiv_rn(__fsub_rn(__ldg(p + j), P.lo[j]), P.vs[j]));
    ok = ok && (q >= 0.0f) && (q < (float)P.grid[j]);  // NaN fails both (undefined in the reference)
    c[j] = (int)q;
  }
  if (!ok) { pslot[i] = -1; return; }
  const int b = scene_of(P, i);
  const unsigned long long key =
      (((unsigned long long)b * P.grid[2] + c[2]) * P.grid[1] + c[1]) * P.grid[0] + c[0];
  unsigned int s = (unsigned int)mix64(key) & P.hash_mask;
  while (true) {
    const unsigned long long prev = atomicCAS(keys + s, kEmptyKey, key);
    if (prev == kEmptyKey || prev == key) break;
    s = (s + 1) & P.hash_mask;
  }
  atomicMin(first + s, i);
  pslot[i] = (int)s;
}

struct IsFirst {
  const int* pslot;
  const int* first;
  __device__ __forceinline__ int operator()(long long i) const {
    const int s = pslot[i];
    return s >= 0 && first[s] == (int)i;
  }
};

// Stage 3.  voxel_offsets[b] = sum_{b'<b} min(count_b', max_voxels).
__global__ void vox_offsets_kernel(const __grid_constant__ VoxParams P, const int* __restrict__ scan,
                                   int* __restrict__ voxel_offsets) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int run = 0;
  for (int b = 0; b < P.batch; ++b) {
    voxel_offsets[b] = run;
    const int cnt = scan[P.off[b + 1]] - scan[P.off[b]];
    run += min(cnt, P.max_voxels);
  }
  voxel_offsets[P.batch] = run;
}

// Stage 4.  vrow[slot] = output row of the voxel or -1 when it lost to the max_voxels cap.
__global__ void __launch_bounds__(256) vox_assign_kernel(const __grid_constant__ VoxParams P,
                                                         const unsigned long long* __restrict__ keys,
                                                         const int* __restrict__ first,
                                                         const int* __restrict__ pslot,
                                                         const int* __restrict__ scan,
                                                         const int* __restrict__ voxel_offsets,
                                                         int* __restrict__ vrow, int* __restrict__ coors,
                                                         int* __restrict__ ptidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_points) return;
  const int s = pslot[i];
  if (s < 0 || first[s] != i) return;
  const int b = scene_of(P, i);
  const int rank = scan[i] - scan[P.off[b]];
  if (rank >= P.max_voxels) { vrow[s] = -1; return; }
  const int row = voxel_offsets[b] + rank;
  vrow[s] = row;
  unsigned long long key = keys[s];
  const int x = (int)(key % P.grid[0]); key /= P.grid[0];
  const int y = (int)(key % P.grid[1]); key /= P.grid[1];
  const int z = (int)(key % P.grid[2]);
  reinterpret_cast<int4*>(coors)[row] = make_int4(b, z, y, x);  // reversed order, :40
  ptidx[(size_t)row * P.max_points] = i;
}

// Stage 5, pass p: ptidx[row][p] = min{ i in voxel : i > ptidx[row][p-1] }.
__global__ void __launch_bounds__(256) vox_kth_kernel(const __grid_constant__ VoxParams P, int p,
                                                      const int* __restrict__ pslot,
                                                      const int* __restrict__ vrow, int* __restrict__ ptidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_points) return;
  const int s = pslot[i];
  if (s < 0) return;
  const int row = vrow[s];
  if (row < 0) return;
  int* t = ptidx + (size_t)row * P.max_points;
  const int prev = t[p - 1];
  if (prev == kInf || i <= prev) return;
  atomicMin(t + p, i);
}

// Stage 6.  One thread per output row.
__global__ void __launch_bounds__(128) vox_gather_kernel(const float* __restrict__ points,
                                                         const __grid_constant__ VoxParams P,
                                                         const int* __restrict__ voxel_offsets,
                                                         const int* __restrict__ ptidx,
                                                         float* __restrict__ voxels, int* __restrict__ num_points,
                                                         float* __restrict__ mean, int mean_channels) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= voxel_offsets[P.batch]) return;
  const int* t = ptidx + (size_t)row * P.max_points;
  int num = 0;
  constexpr int kMaxC = 16;
  float acc[kMaxC];
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) acc[c] = 0.f;
  for (int p = 0; p < P.max_points; ++p) {
    const int i = t[p];
    const bool have = i != kInf;
    num += have;
    const float* src = points + (size_t)(have ? i : 0) * P.F;
    float* dst = voxels ? voxels + ((size_t)row * P.max_points + p) * P.F : nullptr;
    for (int c = 0; c < P.F; ++c) {
      const float v = have ? __ldg(src + c) : 0.f;
      if (dst) dst[c] = v;
      if (c < kMaxC) acc[c] += v;  // slot order, like features.sum(dim=1) over the padded slots
    }
  }
  num_points[row] = num;
  if (mean) {
    const float inv_n = (float)num;
    for (int c = 0; c < mean_channels; ++c) mean[(size_t)row * mean_channels + c] = __fdiv_rn(acc[c], inv_n);
  }
}

__global__ void __launch_bounds__(256) voxel_mean_kernel(const float* __restrict__ voxels,
                                                         const int* __restrict__ num_points, int n, int P,
                                                         int F, int C, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * C) return;
  const int v = idx / C, c = idx - v * C;
  float s = 0.f;
  for (int p = 0; p < P; ++p) s += __ldg(voxels + ((size_t)v * P + p) * F + c);
  out[idx] = __fdiv_rn(s, (float)num_points[v]);
}

struct VoxWorkspace {
  unsigned long long* keys;
  int *first, *vrow, *pslot, *scan, *sums, *ptidx;
  size_t hash_cap, bytes;
};

static VoxWorkspace vox_workspace(void* base, int n_points, int batch, int max_points, int max_voxels) {
  VoxWorkspace w;
  size_t cap = 1024;
  while (cap < 2 * (size_t)(n_points > 0 ? n_points : 1)) cap <<= 1;
  w.hash_cap = cap;
  Carver c(base);
  w.keys = c.take<unsigned long long>(cap);
  w.first = c.take<int>(cap);
  w.vrow = c.take<int>(cap);
  w.pslot = c.take<int>((size_t)n_points + 1);
  w.scan = c.take<int>((size_t)n_points + 2);
  w.sums = c.take<int>((size_t)scan_num_blocks(n_points) + 2);
  w.ptidx = c.take<int>((size_t)batch * max_voxels * max_points + 1);
  w.bytes = c.off;
  return w;
}

static void grid_size_host(const float* range, const float* vs, int* grid) {
  // fp32 round-half-even of (hi - lo) / vs -- point_cloud_ops.py:24-29,143-144
  for (int j = 0; j < 3; ++j) {
    volatile float d = range[3 + j] - range[j];
    volatile float g = d / vs[j];
    grid[j] = (int)rintf(g);
  }
}

}  // namespace s2d

using namespace s2d;

extern "C" size_t s2d_voxelize_workspace_bytes(int n_points, int batch, int max_points, int max_voxels) {
  if (n_points < 0 || batch <= 0 || max_points <= 0 || max_voxels <= 0) return 0;
  return vox_workspace(nullptr, n_points, batch, max_points, max_voxels).bytes;
}

extern "C" int s2d_voxelize(const float* points, const int* scene_offsets_host, int n_points, int batch, int F,
                            const float* range_host, const float* vsize_host, int max_points, int max_voxels,
                            float* voxels, int* coors, int* num_points, float* mean, int mean_channels,
                            int* voxel_offsets, void* workspace, size_t workspace_bytes, void* stream) {
  S2D_REQUIRE(batch >= 1 && batch <= S2D_MAX_BATCH, "s2d_voxelize: batch %d outside [1,%d]", batch, S2D_MAX_BATCH);
  S2D_REQUIRE(n_points >= 0 && F >= 3, "s2d_voxelize: need n_points >= 0 and F >= 3 (got %d, %d)", n_points, F);
  S2D_REQUIRE(max_points >= 1 && max_voxels >= 1, "s2d_voxelize: max_points/max_voxels must be positive");
  S2D_REQUIRE(scene_offsets_host && range_host && vsize_host && coors && num_points && voxel_offsets,
              "s2d_voxelize: null argument");
  S2D_REQUIRE(n_points == 0 || points, "s2d_voxelize: null points");
  S2D_REQUIRE(!mean || (mean_channels >= 1 && mean_channels <= F && mean_channels <= 16),
              "s2d_voxelize: mean_channels %d outside [1,min(F,16)]", mean_channels);
  S2D_REQUIRE(scene_offsets_host[0] == 0 && scene_offsets_host[batch] == n_points,
              "s2d_voxelize: scene_offsets must start at 0 and end at n_points");
  for (int b = 0; b < batch; ++b)
    S2D_REQUIRE(scene_offsets_host[b + 1] >= scene_offsets_host[b], "s2d_voxelize: scene_offsets not monotone");
  VoxWorkspace w = vox_workspace(workspace, n_points, batch, max_points, max_voxels);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("s2d_voxelize: workspace %zu B < required %zu B", workspace_bytes, w.bytes);
    return S2D_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  VoxParams P;
  for (int b = 0; b <= batch; ++b) P.off[b] = scene_offsets_host[b];
  for (int j = 0; j < 3; ++j) { P.lo[j] = range_host[j]; P.vs[j] = vsize_host[j]; }
  grid_size_host(range_host, vsize_host, P.grid);
  S2D_REQUIRE(P.grid[0] > 0 && P.grid[1] > 0 && P.grid[2] > 0, "s2d_voxelize: empty voxel grid");
  P.batch = batch; P.n_points = n_points; P.F = F; P.max_points = max_points; P.max_voxels = max_voxels;
  P.hash_mask = (unsigned int)(w.hash_cap - 1);

  S2D_CUDA(cudaMemsetAsync(w.keys, 0xff, w.hash_cap * sizeof(unsigned long long), st));
  S2D_CUDA(cudaMemsetAsync(w.first, 0x7f, w.hash_cap * sizeof(int), st));
  const size_t rows_bound = (size_t)((long long)batch * max_voxels < n_points ? (long long)batch * max_voxels : n_points);
  S2D_CUDA(cudaMemsetAsync(w.ptidx, 0x7f, (rows_bound * max_points + 1) * sizeof(int), st));
  const int nb = div_up(n_points > 0 ? n_points : 1, 256);
  if (n_points > 0) vox_insert_kernel<<<nb, 256, 0, st>>>(points, P, w.keys, w.first, w.pslot);
  int rc = exclusive_scan(IsFirst{w.pslot, w.first}, n_points, w.scan, w.sums, st);
  if (rc) return rc;
  vox_offsets_kernel<<<1, 32, 0, st>>>(P, w.scan, voxel_offsets);
  if (n_points > 0) {
    vox_assign_kernel<<<nb, 256, 0, st>>>(P, w.keys, w.first, w.pslot, w.scan, voxel_offsets, w.vrow, coors,
                                          w.ptidx);
    for (int p = 1; p < max_points; ++p) vox_kth_kernel<<<nb, 256, 0, st>>>(P, p, w.pslot, w.vrow, w.ptidx);
    const int rows_cap = (int)((long long)batch * max_voxels < n_points ? (long long)batch * max_voxels : n_points);
    vox_gather_kernel<<<div_up(rows_cap, 128), 128, 0, st>>>(points, P, voxel_offsets, w.ptidx, voxels, num_points,
                                                            mean, mean_channels);
  }
  S2D_LAUNCH_CHECK();
  count_launches(4 + (n_points > 0 ? 2 + max_points : 0));
  return S2D_OK;
}

extern "C" int s2d_voxel_mean(const float* voxels, const int* num_points, int n_voxels, int max_points, int F,
                              int channels, float* out, void* stream) {
  S2D_REQUIRE(n_voxels >= 0 && max_points >= 1 && channels >= 1 && channels <= F, "s2d_voxel_mean: bad shape");
  if (n_voxels == 0) return S2D_OK;
  S2D_REQUIRE(voxels && num_points && out, "s2d_voxel_mean: null argument");
  voxel_mean_kernel<<<div_up((long long)n_voxels * channels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      voxels, num_points, n_voxels, max_points, F, channels, out);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
