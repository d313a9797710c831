// rulebook.cu -- coordinate indices and rulebooks ("indice pairs") for sparse 3-D convolution.
//
// Replaces the rulebook builders of spconv v1.x (create_submconv_indice_pair /
// create_conv_indice_pair_p1,p2 -- an un-vendored dependency of the reference, used through
// spconv.SubMConv3d / spconv.SparseConv3d at det3d/models/backbones/scn.py:16-39,104-152).
//
// B200 design: no hash, no sort.  A sparse tensor's coordinate set is an occupancy BITMAP over
// the flattened (b,z,y,x) grid plus a per-word popcount prefix; rank(lin) = prefix[word] +
// popc(bits below) enumerates active voxels in ascending flattened order, which is also the
// canonical output order of a strided conv.  The bitmap of a Waymo batch is
// a few MB (12 MB per scene at the 41x1504x1504 input grid), i.e. L2 resident, and the 27
// neighbour probes of one voxel touch <= 9 words.  Results are deterministic and independent
// of thread scheduling, so rulebooks can be compared bit for bit with the CPU oracle.
//
// All kernels are integer, HBM/L2-latency bound, one thread per row (or per bitmap word).
#include "common.cuh"
#include "grouping.cuh"

namespace s2d {

struct ShapeP {
  int batch, D, H, W;
  __host__ __device__ long long volume() const { return (long long)batch * D * H * W; }
  __device__ __forceinline__ long long lin(int b, int z, int y, int x) const {
    return (((long long)b * D + z) * H + y) * W + x;
  }
};

struct ConvP {
  int k[3], s[3], p[3], d[3];
};

GridIndexLayout grid_index_layout(int batch, const int* shape, int n_rows_capacity) {
  GridIndexLayout L;
  const long long vol = (long long)batch * shape[0] * shape[1] * shape[2];
  L.n_words = (vol + 31) / 32;
  size_t off = 0;
  L.words_off = off;  off += align_up((size_t)(L.n_words + 1) * 4, 256);
  L.prefix_off = off; off += align_up((size_t)(L.n_words + 2) * 4, 256);
  L.sums_off = off;   off += align_up((size_t)(scan_num_blocks(L.n_words) + 2) * 4, 256);
  L.perm_off = off;   off += align_up((size_t)(n_rows_capacity > 0 ? n_rows_capacity : 1) * 4, 256);
  L.total = off;
  return L;
}

GridIndexPtrs grid_index_ptrs(const void* index, const GridIndexLayout& L) {
  char* b = const_cast<char*>(static_cast<const char*>(index));
  GridIndexPtrs p;
  p.words = reinterpret_cast<uint32_t*>(b + L.words_off);
  p.prefix = reinterpret_cast<int*>(b + L.prefix_off);
  p.perm = reinterpret_cast<int*>(b + L.perm_off);
  p.sums = reinterpret_cast<int*>(b + L.sums_off);
  p.n_words = L.n_words;
  return p;
}

__global__ void scan_block_sums_kernel(int* block_sums, int nb) {
  // single block of 1024 threads; thread t owns a contiguous chunk
  const int per = (nb + 1023) / 1024;
  const int lo = min(threadIdx.x * per, nb), hi = min(lo + per, nb);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += block_sums[i];
  int total;
  const int inc = block_inclusive_scan<1024>(s, &total);
  int run = inc - s;
  for (int i = lo; i < hi; ++i) {
    const int v = block_sums[i];
    block_sums[i] = run;
    run += v;
  }
  if (threadIdx.x == 0) block_sums[nb] = total;
}

struct PopcWord {
  const uint32_t* words;
  __device__ __forceinline__ int operator()(long long i) const { return __popc(words[i]); }
};

__device__ __forceinline__ int row_count(int n_bound, const int* n_dev) {
  return n_dev ? min(*n_dev, n_bound) : n_bound;
}

__device__ __forceinline__ void set_bit(uint32_t* words, long long lin) {
  const uint32_t bit = 1u << (lin & 31);
  uint32_t* w = words + (lin >> 5);
  if (!(*w & bit)) atomicOr(w, bit);
}

__global__ void __launch_bounds__(256) mark_rows_kernel(const int4* __restrict__ coors, int n_bound,
                                                        const int* __restrict__ n_dev, ShapeP S,
                                                        uint32_t* __restrict__ words) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= row_count(n_bound, n_dev)) return;
  const int4 c = coors[i];  // (b,z,y,x)
  if ((unsigned)c.x >= (unsigned)S.batch || (unsigned)c.y >= (unsigned)S.D || (unsigned)c.z >= (unsigned)S.H ||
      (unsigned)c.w >= (unsigned)S.W)
    return;  // out-of-grid rows are never found by a lookup
  set_bit(words, S.lin(c.x, c.y, c.z, c.w));
}

__global__ void __launch_bounds__(256) fill_perm_kernel(const int4* __restrict__ coors, int n_bound,
                                                        const int* __restrict__ n_dev, ShapeP S,
                                                        GridIndexView idx, int* __restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= row_count(n_bound, n_dev)) return;
  const int4 c = coors[i];
  if ((unsigned)c.x >= (unsigned)S.batch || (unsigned)c.y >= (unsigned)S.D || (unsigned)c.z >= (unsigned)S.H ||
      (unsigned)c.w >= (unsigned)S.W)
    return;
  const int r = idx.rank(S.lin(c.x, c.y, c.z, c.w));
  if (r >= 0) perm[r] = i;
}

// Strided conv, phase 1: mark every output site reached by an active input.
//   o = (x + p - k*d) / s  when divisible and inside the output grid   (x = o*s - p + k*d)
__global__ void __launch_bounds__(256) mark_outputs_kernel(const int4* __restrict__ coors, int n_bound,
                                                           const int* __restrict__ n_dev, ConvP C, ShapeP So,
                                                           uint32_t* __restrict__ words) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= row_count(n_bound, n_dev)) return;
  const int4 c = coors[i];
  for (int kz = 0; kz < C.k[0]; ++kz) {
    const int tz = c.y + C.p[0] - kz * C.d[0];
    if (tz < 0 || tz % C.s[0]) continue;
    const int oz = tz / C.s[0];
    if (oz >= So.D) continue;
    for (int ky = 0; ky < C.k[1]; ++ky) {
      const int ty = c.z + C.p[1] - ky * C.d[1];
      if (ty < 0 || ty % C.s[1]) continue;
      const int oy = ty / C.s[1];
      if (oy >= So.H) continue;
      for (int kx = 0; kx < C.k[2]; ++kx) {
        const int tx = c.w + C.p[2] - kx * C.d[2];
        if (tx < 0 || tx % C.s[2]) continue;
        const int ox = tx / C.s[2];
        if (ox >= So.W) continue;
        set_bit(words, So.lin(c.x, oz, oy, ox));
      }
    }
  }
}

// Strided conv, phase 2: enumerate the set bits in ascending order -> output coordinates.
__global__ void __launch_bounds__(256) emit_coords_kernel(const uint32_t* __restrict__ words,
                                                          const int* __restrict__ prefix, long long n_words,
                                                          ShapeP So, int4* __restrict__ out_coors, int capacity,
                                                          int* __restrict__ perm, int* __restrict__ n_out) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w == 0) *n_out = prefix[n_words];
  if (w >= n_words) return;
  uint32_t bits = words[w];
  int row = prefix[w];
  while (bits) {
    const int j = __ffs(bits) - 1;
    bits &= bits - 1;
    if (row < capacity) {
      long long l = w * 32 + j;
      const int x = (int)(l % So.W); l /= So.W;
      const int y = (int)(l % So.H); l /= So.H;
      const int z = (int)(l % So.D); l /= So.D;
      out_coors[row] = make_int4((int)l, z, y, x);
      perm[row] = row;
    }
    ++row;
  }
}

__device__ __forceinline__ void add_pairs(unsigned long long* n_pairs, int local) {
  if (!n_pairs) return;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_pairs, (unsigned long long)local);
}

// SubM table: tbl[k][i] = row of the active voxel at x_i + (k - ksize/2)*dilation, or -1.
__global__ void __launch_bounds__(256) subm_table_kernel(const int4* __restrict__ coors, int n, ShapeP S, ConvP C,
                                                         GridIndexView idx, int* __restrict__ tbl, int stride,
                                                         unsigned long long* n_pairs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int local = 0;
  if (i < n) {
    const int4 c = coors[i];
    int k = 0;
    for (int kz = 0; kz < C.k[0]; ++kz) {
      const int z = c.y + (kz - C.k[0] / 2) * C.d[0];
      for (int ky = 0; ky < C.k[1]; ++ky) {
        const int y = c.z + (ky - C.k[1] / 2) * C.d[1];
        for (int kx = 0; kx < C.k[2]; ++kx, ++k) {
          const int x = c.w + (kx - C.k[2] / 2) * C.d[2];
          int j = -1;
          if ((unsigned)z < (unsigned)S.D && (unsigned)y < (unsigned)S.H && (unsigned)x < (unsigned)S.W)
            j = idx.lookup(S.lin(c.x, z, y, x));
          tbl[(size_t)k * stride + i] = j;
          local += j >= 0;
        }
      }
    }
  }
  add_pairs(n_pairs, local);
}

// ---- 3 x 3 x 3 rulebooks built directly in GROUPED row order (conv_bf2.cu "Row grouping") ----------------------------------
// Input position of offset k for output voxel c: c * s - p + k * d (a submanifold layer is s = 1, p = d).
// Pass 1: the 9-bit key of a row needs only the occupancy bits of its 27 neighbours (no rank, no permutation lookup);
// pass 2 (after the counting sort of grouping.cuh): thread p builds the table column of row perm[p] and the block's
// ballots give the live-offset mask of the 128-row tile.  The scan-order table is never written.
__global__ void __launch_bounds__(kGrpRows) subm_keys_hist_kernel(const int4* __restrict__ coors, int n, ShapeP S, ConvP C,
                                                                  GridIndexView idx, unsigned short* __restrict__ keys,
                                                                  int* __restrict__ counts) {
  __shared__ __align__(16) unsigned short s_cnt[32][kGrpBuckets];
  const int row = blockIdx.x * kGrpRows + threadIdx.x;
  unsigned key = 0;
  if (row < n) {
    const int4 c = coors[row];
    int k = 0;
    for (int kz = 0; kz < 3; ++kz) {
      const int z = c.y * C.s[0] - C.p[0] + kz * C.d[0];
      for (int ky = 0; ky < 3; ++ky) {
        const int y = c.z * C.s[1] - C.p[1] + ky * C.d[1];
        for (int kx = 0; kx < 3; ++kx, ++k) {
          const int x = c.w * C.s[2] - C.p[2] + kx * C.d[2];
          if ((unsigned)z < (unsigned)S.D && (unsigned)y < (unsigned)S.H && (unsigned)x < (unsigned)S.W) {
            const long long lin = S.lin(c.x, z, y, x);
            if ((__ldg(idx.words + (lin >> 5)) >> (lin & 31)) & 1u) key |= 1u << (k / 3);
          }
        }
      }
    }
    keys[row] = (unsigned short)key;
  }
  group_block_counts(key, row < n, s_cnt);
  if (threadIdx.x < kGrpBuckets) {
    int total = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) total += s_cnt[w][threadIdx.x];
    counts[(size_t)blockIdx.x * kGrpBuckets + threadIdx.x] = total;
  }
}

__global__ void __launch_bounds__(128) subm_table_grouped_kernel(const int4* __restrict__ coors, int n, ShapeP S, ConvP C,
                                                                 GridIndexView idx, const int* __restrict__ perm,
                                                                 int* __restrict__ tbl, int stride, int* __restrict__ masks) {
  __shared__ unsigned s_mask;
  if (threadIdx.x == 0) s_mask = 0u;
  __syncthreads();
  const int p = blockIdx.x * 128 + threadIdx.x;
  int4 c = make_int4(0, 0, 0, 0);
  if (p < n) c = coors[__ldg(perm + p)];
  unsigned m = 0;
  int k = 0;
  for (int kz = 0; kz < 3; ++kz) {
    const int z = c.y * C.s[0] - C.p[0] + kz * C.d[0];
    for (int ky = 0; ky < 3; ++ky) {
      const int y = c.z * C.s[1] - C.p[1] + ky * C.d[1];
      for (int kx = 0; kx < 3; ++kx, ++k) {
        const int x = c.w * C.s[2] - C.p[2] + kx * C.d[2];
        int j = -1;
        if (p < n && (unsigned)z < (unsigned)S.D && (unsigned)y < (unsigned)S.H && (unsigned)x < (unsigned)S.W)
          j = idx.lookup(S.lin(c.x, z, y, x));
        if (p < n) tbl[(size_t)k * stride + p] = j;
        if (__ballot_sync(0xffffffffu, j >= 0)) m |= 1u << k;
      }
    }
  }
  if ((threadIdx.x & 31) == 0 && m) atomicOr(&s_mask, m);
  __syncthreads();
  if (threadIdx.x == 0) masks[blockIdx.x] = (int)s_mask;
}

// Strided table: tbl[k][o] = row of the active input at o*s - p + k*d, or -1.
__global__ void __launch_bounds__(256) sparse_table_kernel(const int4* __restrict__ out_coors, int n_out,
                                                           ShapeP Si, ConvP C, GridIndexView idx_in,
                                                           int* __restrict__ tbl, int stride,
                                                           unsigned long long* n_pairs) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  int local = 0;
  if (o < n_out) {
    const int4 c = out_coors[o];
    int k = 0;
    for (int kz = 0; kz < C.k[0]; ++kz) {
      const int z = c.y * C.s[0] - C.p[0] + kz * C.d[0];
      for (int ky = 0; ky < C.k[1]; ++ky) {
        const int y = c.z * C.s[1] - C.p[1] + ky * C.d[1];
        for (int kx = 0; kx < C.k[2]; ++kx, ++k) {
          const int x = c.w * C.s[2] - C.p[2] + kx * C.d[2];
          int j = -1;
          if ((unsigned)z < (unsigned)Si.D && (unsigned)y < (unsigned)Si.H && (unsigned)x < (unsigned)Si.W)
            j = idx_in.lookup(Si.lin(c.x, z, y, x));
          tbl[(size_t)k * stride + o] = j;
          local += j >= 0;
        }
      }
    }
  }
  add_pairs(n_pairs, local);
}

// dense() + view(N, C*D, H, W): one warp moves one row; the zero fill is a memset before it.
__global__ void __launch_bounds__(256) dense_bev_kernel(const float* __restrict__ feat,
                                                        const int4* __restrict__ coors, int n, int C, int D, int H,
                                                        int W, float* __restrict__ bev) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const int4 c = coors[row];
  const size_t plane = (size_t)H * W;
  float* dst = bev + ((size_t)c.x * C * D + c.y) * plane + (size_t)c.z * W + c.w;
  for (int ch = lane; ch < C; ch += 32) dst[(size_t)ch * D * plane] = __ldg(feat + (size_t)row * C + ch);
}

// Output-stationary form of dense() + view: cell map (row index or -1 per (b,z,y,x)), then one warp per 32 consecutive
// x cells moves 32-channel chunks through a padded shared-memory tile, so every global access is a full 128 B line
// (the row-stationary kernel above writes 4-byte pieces into C*D different planes) and empty cells are written as
// zeros in the same pass (no separate memset of the 36 MB / scene map).
__global__ void __launch_bounds__(256) bev_cell_map_kernel(const int4* __restrict__ coors, int n, int D, int H, int W,
                                                           int* __restrict__ cell_row) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const int4 c = coors[row];
  cell_row[(((size_t)c.x * D + c.y) * H + c.z) * W + c.w] = row;
}

__global__ void __launch_bounds__(256) dense_bev_tiled_kernel(const float* __restrict__ feat,
                                                              const int* __restrict__ cell_row, int C, int batch, int D,
                                                              int H, int W, float* __restrict__ bev) {
  __shared__ float tile[8][32][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xt = (W + 31) / 32;
  const long long task = (long long)blockIdx.x * 8 + warp;              // (b, z, y, x tile)
  if (task >= (long long)batch * D * H * xt) return;
  const int x0 = (int)(task % xt) * 32;
  const int y = (int)((task / xt) % H), z = (int)((task / ((long long)xt * H)) % D), b = (int)(task / ((long long)xt * H * D));
  const int x = x0 + lane;
  const int my_row = x < W ? __ldg(cell_row + (((size_t)b * D + z) * H + y) * W + x) : -1;
  const size_t plane = (size_t)H * W;
  float (*t)[33] = tile[warp];
  for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {                                       // cell i of the tile: 32 consecutive channels
      const int r = __shfl_sync(0xffffffffu, my_row, i);
      t[i][lane] = (r >= 0 && c0 + lane < C) ? __ldg(feat + (size_t)r * C + c0 + lane) : 0.f;
    }
    __syncwarp();
    if (x < W) {
#pragma unroll 4
      for (int ch = 0; ch < 32; ++ch)
        if (c0 + ch < C) bev[((size_t)b * C * D + (size_t)(c0 + ch) * D + z) * plane + (size_t)y * W + x] = t[lane][ch];
    }
    __syncwarp();
  }
}

static int check_shape(int batch, const int* shape, const char* who) {
  S2D_REQUIRE(shape && batch >= 1 && shape[0] > 0 && shape[1] > 0 && shape[2] > 0, "%s: bad batch/shape", who);
  S2D_REQUIRE((long long)batch * shape[0] * shape[1] * shape[2] < (1ll << 36), "%s: grid too large", who);
  return S2D_OK;
}

static int load_conv(ConvP& C, const int* k, const int* s, const int* p, const int* d, const char* who) {
  S2D_REQUIRE(k && d, "%s: null ksize/dilation", who);
  for (int a = 0; a < 3; ++a) {
    C.k[a] = k[a]; C.s[a] = s ? s[a] : 1; C.p[a] = p ? p[a] : 0; C.d[a] = d[a];
    S2D_REQUIRE(C.k[a] >= 1 && C.s[a] >= 1 && C.p[a] >= 0 && C.d[a] >= 1, "%s: bad conv geometry", who);
  }
  return S2D_OK;
}

static int build_prefix(GridIndexPtrs& I, cudaStream_t st) {
  return exclusive_scan(PopcWord{I.words}, I.n_words, I.prefix, I.sums, st);
}

}  // namespace s2d

using namespace s2d;

extern "C" size_t s2d_grid_index_bytes(int batch, const int* shape_host, int n_rows_capacity) {
  if (!shape_host || batch < 1 || n_rows_capacity < 0) return 0;
  return grid_index_layout(batch, shape_host, n_rows_capacity).total;
}

extern "C" int s2d_grid_index_build(const int* coors, int n_rows, const int* n_rows_dev, int batch,
                                    const int* shape_host, void* index, size_t index_bytes, void* stream) {
  int rc = check_shape(batch, shape_host, "s2d_grid_index_build");
  if (rc) return rc;
  S2D_REQUIRE(n_rows >= 0 && index && (n_rows == 0 || coors), "s2d_grid_index_build: null argument");
  const GridIndexLayout L = grid_index_layout(batch, shape_host, n_rows);
  if (index_bytes < L.total) {
    set_error("s2d_grid_index_build: index memory %zu B < required %zu B", index_bytes, L.total);
    return S2D_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GridIndexPtrs I = grid_index_ptrs(index, L);
  const ShapeP S{batch, shape_host[0], shape_host[1], shape_host[2]};
  S2D_CUDA(cudaMemsetAsync(I.words, 0, (size_t)(L.n_words + 1) * 4, st));
  const int nb = div_up(n_rows > 0 ? n_rows : 1, 256);
  if (n_rows > 0)
    mark_rows_kernel<<<nb, 256, 0, st>>>(reinterpret_cast<const int4*>(coors), n_rows, n_rows_dev, S, I.words);
  rc = build_prefix(I, st);
  if (rc) return rc;
  if (n_rows > 0)
    fill_perm_kernel<<<nb, 256, 0, st>>>(reinterpret_cast<const int4*>(coors), n_rows, n_rows_dev, S, I.view(),
                                         I.perm);
  S2D_LAUNCH_CHECK();
  count_launches(3 + (n_rows > 0 ? 2 : 0));
  return S2D_OK;
}

extern "C" int s2d_rulebook_subm(const int* coors, int n_rows, int batch, const int* shape_host,
                                 const int* ksize_host, const int* dilation_host, const void* index, int* tbl, int tbl_stride, unsigned long long* n_pairs,
                                 void* stream) {
  int rc = check_shape(batch, shape_host, "s2d_rulebook_subm");
  if (rc) return rc;
  ConvP C;
  rc = load_conv(C, ksize_host, nullptr, nullptr, dilation_host, "s2d_rulebook_subm");
  if (rc) return rc;
  S2D_REQUIRE(n_rows >= 0 && tbl_stride >= n_rows, "s2d_rulebook_subm: tbl_stride %d < n_rows %d", tbl_stride, n_rows);
  if (n_rows == 0) return S2D_OK;
  S2D_REQUIRE(coors && index && tbl, "s2d_rulebook_subm: null argument");
  const GridIndexLayout L = grid_index_layout(batch, shape_host, 0);
  const GridIndexPtrs I = grid_index_ptrs(index, L);
  const ShapeP S{batch, shape_host[0], shape_host[1], shape_host[2]};
  subm_table_kernel<<<div_up(n_rows, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const int4*>(coors), n_rows, S, C, I.view(), tbl, tbl_stride, n_pairs);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" size_t s2d_rulebook_subm_grouped_workspace_bytes(int n_rows) { return n_rows < 0 ? 0 : group_workspace_bytes(n_rows); }

namespace s2d {
// shared by the two grouped builders: coors = OUTPUT coordinates, index / shape = INPUT tensor
static int build_grouped(const int* coors, int n_rows, int batch, const int* shape_in_host, const ConvP& C, const void* index,
                         int* perm, int* tbl, int tbl_stride, int* tile_masks, void* workspace, cudaStream_t st) {
  const GridIndexLayout L = grid_index_layout(batch, shape_in_host, 0);
  const GridIndexPtrs I = grid_index_ptrs(index, L);
  const ShapeP S{batch, shape_in_host[0], shape_in_host[1], shape_in_host[2]};
  const int nblk = div_up(n_rows, kGrpRows);
  int* counts = static_cast<int*>(workspace);
  int* tails = counts + (size_t)nblk * kGrpBuckets;
  unsigned short* keys = reinterpret_cast<unsigned short*>(tails + kGrpSegs * kGrpBuckets);
  const int4* c4 = reinterpret_cast<const int4*>(coors);
  subm_keys_hist_kernel<<<nblk, kGrpRows, 0, st>>>(c4, n_rows, S, C, I.view(), keys, counts);
  group_scan(counts, nblk, tails, st);
  group_scatter_kernel<<<nblk, kGrpRows, 0, st>>>(keys, n_rows, nblk, counts, tails, perm);
  subm_table_grouped_kernel<<<div_up(n_rows, 128), 128, 0, st>>>(c4, n_rows, S, C, I.view(), perm, tbl, tbl_stride, tile_masks);
  S2D_LAUNCH_CHECK();
  count_launches(5);
  return S2D_OK;
}
}  // namespace s2d

extern "C" int s2d_rulebook_subm_grouped(const int* coors, int n_rows, int batch, const int* shape_host,
                                         const int* dilation_host, const void* index, int* perm, int* tbl, int tbl_stride,
                                         int* tile_masks, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_shape(batch, shape_host, "s2d_rulebook_subm_grouped");
  if (rc) return rc;
  const int k3[3] = {3, 3, 3};
  ConvP C;
  rc = load_conv(C, k3, nullptr, nullptr, dilation_host, "s2d_rulebook_subm_grouped");
  if (rc) return rc;
  for (int a = 0; a < 3; ++a) { C.s[a] = 1; C.p[a] = C.d[a]; }          // centred kernel: c - d + k * d
  S2D_REQUIRE(n_rows >= 0 && tbl_stride >= n_rows, "s2d_rulebook_subm_grouped: tbl_stride %d < n_rows %d", tbl_stride, n_rows);
  if (n_rows == 0) return S2D_OK;
  S2D_REQUIRE(coors && index && perm && tbl && tile_masks && workspace, "s2d_rulebook_subm_grouped: null argument");
  S2D_REQUIRE(workspace_bytes >= group_workspace_bytes(n_rows), "s2d_rulebook_subm_grouped: workspace too small");
  return build_grouped(coors, n_rows, batch, shape_host, C, index, perm, tbl, tbl_stride, tile_masks, workspace,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int s2d_rulebook_sparse_grouped(const int* out_coors, int n_out, int batch, const int* shape_in_host,
                                           const int* stride_host, const int* pad_host, const int* dilation_host,
                                           const void* index_in, int* perm, int* tbl, int tbl_stride, int* tile_masks,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_shape(batch, shape_in_host, "s2d_rulebook_sparse_grouped");
  if (rc) return rc;
  const int k3[3] = {3, 3, 3};
  ConvP C;
  rc = load_conv(C, k3, stride_host, pad_host, dilation_host, "s2d_rulebook_sparse_grouped");
  if (rc) return rc;
  S2D_REQUIRE(n_out >= 0 && tbl_stride >= n_out, "s2d_rulebook_sparse_grouped: tbl_stride %d < n_out %d", tbl_stride, n_out);
  if (n_out == 0) return S2D_OK;
  S2D_REQUIRE(out_coors && index_in && perm && tbl && tile_masks && workspace, "s2d_rulebook_sparse_grouped: null argument");
  S2D_REQUIRE(workspace_bytes >= group_workspace_bytes(n_out), "s2d_rulebook_sparse_grouped: workspace too small");
  return build_grouped(out_coors, n_out, batch, shape_in_host, C, index_in, perm, tbl, tbl_stride, tile_masks, workspace,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int s2d_conv_out_shape(const int* shape_in_host, const int* ksize_host, const int* stride_host,
                                  const int* pad_host, const int* dilation_host, int* shape_out_host) {
  S2D_REQUIRE(shape_in_host && shape_out_host, "s2d_conv_out_shape: null argument");
  ConvP C;
  int rc = load_conv(C, ksize_host, stride_host, pad_host, dilation_host, "s2d_conv_out_shape");
  if (rc) return rc;
  for (int a = 0; a < 3; ++a)
    shape_out_host[a] = (shape_in_host[a] + 2 * C.p[a] - C.d[a] * (C.k[a] - 1) - 1) / C.s[a] + 1;
  return S2D_OK;
}

extern "C" int s2d_sparse_out_coords(const int* coors_in, int n_in, const int* n_in_dev, int batch,
                                     const int* shape_in_host, const int* ksize_host, const int* stride_host,
                                     const int* pad_host, const int* dilation_host, void* index_out,
                                     size_t index_out_bytes, int* out_coors, int out_capacity, int* n_out,
                                     void* stream) {
  int rc = check_shape(batch, shape_in_host, "s2d_sparse_out_coords");
  if (rc) return rc;
  ConvP C;
  rc = load_conv(C, ksize_host, stride_host, pad_host, dilation_host, "s2d_sparse_out_coords");
  if (rc) return rc;
  int so[3];
  s2d_conv_out_shape(shape_in_host, ksize_host, stride_host, pad_host, dilation_host, so);
  rc = check_shape(batch, so, "s2d_sparse_out_coords(output grid)");
  if (rc) return rc;
  S2D_REQUIRE(n_in >= 0 && out_capacity >= 0 && index_out && n_out && (out_capacity == 0 || out_coors),
              "s2d_sparse_out_coords: null argument");
  const GridIndexLayout L = grid_index_layout(batch, so, out_capacity);
  if (index_out_bytes < L.total) {
    set_error("s2d_sparse_out_coords: index memory %zu B < required %zu B", index_out_bytes, L.total);
    return S2D_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GridIndexPtrs I = grid_index_ptrs(index_out, L);
  const ShapeP So{batch, so[0], so[1], so[2]};
  S2D_CUDA(cudaMemsetAsync(I.words, 0, (size_t)(L.n_words + 1) * 4, st));
  if (n_in > 0)
    mark_outputs_kernel<<<div_up(n_in, 256), 256, 0, st>>>(reinterpret_cast<const int4*>(coors_in), n_in, n_in_dev,
                                                           C, So, I.words);
  rc = build_prefix(I, st);
  if (rc) return rc;
  emit_coords_kernel<<<div_up(L.n_words, 256), 256, 0, st>>>(I.words, I.prefix, L.n_words, So,
                                                             reinterpret_cast<int4*>(out_coors), out_capacity,
                                                             I.perm, n_out);
  S2D_LAUNCH_CHECK();
  count_launches(4 + (n_in > 0 ? 1 : 0));
  return S2D_OK;
}

extern "C" int s2d_rulebook_sparse(const int* out_coors, int n_out, int batch, const int* shape_in_host,
                                   const int* ksize_host, const int* stride_host, const int* pad_host,
                                   const int* dilation_host, const void* index_in, int* tbl, int tbl_stride, unsigned long long* n_pairs, void* stream) {
  int rc = check_shape(batch, shape_in_host, "s2d_rulebook_sparse");
  if (rc) return rc;
  ConvP C;
  rc = load_conv(C, ksize_host, stride_host, pad_host, dilation_host, "s2d_rulebook_sparse");
  if (rc) return rc;
  S2D_REQUIRE(n_out >= 0 && tbl_stride >= n_out, "s2d_rulebook_sparse: tbl_stride %d < n_out %d", tbl_stride, n_out);
  if (n_out == 0) return S2D_OK;
  S2D_REQUIRE(out_coors && index_in && tbl, "s2d_rulebook_sparse: null argument");
  const GridIndexLayout L = grid_index_layout(batch, shape_in_host, 0);
  const GridIndexPtrs I = grid_index_ptrs(index_in, L);
  const ShapeP Si{batch, shape_in_host[0], shape_in_host[1], shape_in_host[2]};
  sparse_table_kernel<<<div_up(n_out, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const int4*>(out_coors), n_out, Si, C, I.view(), tbl, tbl_stride, n_pairs);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}

extern "C" size_t s2d_dense_bev_workspace_bytes(int batch, int D, int H, int W) {
  if (batch < 1 || D < 1 || H < 1 || W < 1) return 0;
  return (size_t)batch * D * H * W * sizeof(int);
}

extern "C" int s2d_dense_bev_tiled(const float* feat, const int* coors, int n_rows, int C, int batch, int D, int H, int W,
                                   float* bev, void* workspace, size_t workspace_bytes, void* stream) {
  S2D_REQUIRE(n_rows >= 0 && C >= 1 && batch >= 1 && D >= 1 && H >= 1 && W >= 1 && bev && workspace,
              "s2d_dense_bev_tiled: bad argument");
  S2D_REQUIRE(workspace_bytes >= s2d_dense_bev_workspace_bytes(batch, D, H, W), "s2d_dense_bev_tiled: workspace too small");
  S2D_REQUIRE(n_rows == 0 || (feat && coors), "s2d_dense_bev_tiled: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* cell_row = static_cast<int*>(workspace);
  S2D_CUDA(cudaMemsetAsync(cell_row, 0xff, s2d_dense_bev_workspace_bytes(batch, D, H, W), st));
  if (n_rows > 0)
    bev_cell_map_kernel<<<div_up(n_rows, 256), 256, 0, st>>>(reinterpret_cast<const int4*>(coors), n_rows, D, H, W, cell_row);
  const long long tasks = (long long)batch * D * H * ((W + 31) / 32);
  dense_bev_tiled_kernel<<<div_up(tasks, 8), 256, 0, st>>>(feat, cell_row, C, batch, D, H, W, bev);
  S2D_LAUNCH_CHECK();
  count_launches(n_rows > 0 ? 2 : 1);
  return S2D_OK;
}

extern "C" int s2d_dense_bev(const float* feat, const int* coors, int n_rows, int C, int batch, int D, int H, int W,
                             float* bev, void* stream) {
  S2D_REQUIRE(n_rows >= 0 && C >= 1 && batch >= 1 && D >= 1 && H >= 1 && W >= 1 && bev, "s2d_dense_bev: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  S2D_CUDA(cudaMemsetAsync(bev, 0, (size_t)batch * C * D * H * W * sizeof(float), st));
  if (n_rows == 0) return S2D_OK;
  S2D_REQUIRE(feat && coors, "s2d_dense_bev: null argument");
  dense_bev_kernel<<<div_up((long long)n_rows * 32, 256), 256, 0, st>>>(feat, reinterpret_cast<const int4*>(coors),
                                                                      n_rows, C, D, H, W, bev);
  S2D_LAUNCH_CHECK();
  count_launches(1);
  return S2D_OK;
}
