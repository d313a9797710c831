// grouping.cuh -- stable counting sort of rulebook rows by their 9-bit neighbour-pattern key (see conv_bf2.cu, "Row grouping"):
// shared by s2d_table_group_rows (key from a finished table) and s2d_rulebook_subm_grouped (key from the occupancy bitmap).
#pragma once
#include "common.cuh"

namespace s2d {

constexpr int kGrpRows = 1024;     // rows per block of the counting sort (one per thread)
constexpr int kGrpBuckets = 512;

// per-warp bucket counts of a block's 1024 rows: s_cnt[w][key] = rows of warp w with that key (match.any: no atomics, so
// the order inside a bucket is the row order = a STABLE sort).  Returns this thread's rank among its warp's equal keys.
__device__ __forceinline__ int group_block_counts(unsigned key, bool valid, unsigned short (*s_cnt)[kGrpBuckets]) {
  uint4* z = reinterpret_cast<uint4*>(&s_cnt[0][0]);
  for (int i = threadIdx.x; i < 32 * kGrpBuckets * 2 / 16; i += kGrpRows) z[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const unsigned m = __match_any_sync(0xffffffffu, valid ? key : 0xffffu);
  const int lane = threadIdx.x & 31;
  if (valid && lane == __ffs(m) - 1) s_cnt[threadIdx.x >> 5][key] = (unsigned short)__popc(m);
  __syncthreads();
  return __popc(m & ((1u << lane) - 1u));
}

// counts[blk][b] -> exclusive prefix over the blocks of each half of the block range; tails[0..511] = start of bucket b in
// the sorted order, tails[512..1023] = rows of bucket b in the first half (added by the scatter to second-half blocks)
static __global__ void __launch_bounds__(1024) group_scan_kernel(int* __restrict__ counts, int nblk, int* __restrict__ tails) {
  __shared__ int s_tot[2][kGrpBuckets];
  __shared__ int s_scan[kGrpBuckets];
  const int b = threadIdx.x & (kGrpBuckets - 1), half = threadIdx.x >> 9;
  const int mid = nblk / 2;
  const int lo = half ? mid : 0, hi = half ? nblk : mid;
  int run = 0;
  int blk = lo;
  for (; blk + 8 <= hi; blk += 8) {
    int c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) c[u] = counts[(size_t)(blk + u) * kGrpBuckets + b];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      counts[(size_t)(blk + u) * kGrpBuckets + b] = run;
      run += c[u];
    }
  }
  for (; blk < hi; ++blk) {
    const int c = counts[(size_t)blk * kGrpBuckets + b];
    counts[(size_t)blk * kGrpBuckets + b] = run;
    run += c;
  }
  s_tot[half][b] = run;
  __syncthreads();
  if (half == 0) s_scan[b] = s_tot[0][b] + s_tot[1][b];
  __syncthreads();
  for (int off = 1; off < kGrpBuckets; off <<= 1) {
    int add = 0;
    if (half == 0 && b >= off) add = s_scan[b - off];
    __syncthreads();
    if (half == 0) s_scan[b] += add;
    __syncthreads();
  }
  if (half == 0) {
    tails[b] = s_scan[b] - (s_tot[0][b] + s_tot[1][b]);
    tails[kGrpBuckets + b] = s_tot[0][b];
  }
}

static __global__ void __launch_bounds__(kGrpRows) group_scatter_kernel(const unsigned short* __restrict__ keys, int n, int nblk,
                                                                 const int* __restrict__ counts, const int* __restrict__ tails,
                                                                 int* __restrict__ perm) {
  __shared__ __align__(16) unsigned short s_cnt[32][kGrpBuckets];
  const int row = blockIdx.x * kGrpRows + threadIdx.x;
  const bool valid = row < n;
  const unsigned key = valid ? keys[row] : 0u;
  const int rank = group_block_counts(key, valid, s_cnt);
  if (threadIdx.x < kGrpBuckets) {                       // exclusive prefix over the block's warps, per bucket
    int run = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) {
      const int c = s_cnt[w][threadIdx.x];
      s_cnt[w][threadIdx.x] = (unsigned short)run;
      run += c;
    }
  }
  __syncthreads();
  if (valid) {
    const int pos = counts[(size_t)blockIdx.x * kGrpBuckets + key] + tails[key] +
                    ((int)blockIdx.x >= nblk / 2 ? tails[kGrpBuckets + key] : 0) + s_cnt[threadIdx.x >> 5][key] + rank;
    perm[pos] = row;
  }
}


inline size_t group_workspace_bytes(int n_rows) {
  const size_t nblk = (size_t)div_up(n_rows > 0 ? n_rows : 1, kGrpRows);
  return (nblk + 2) * kGrpBuckets * sizeof(int) + (((size_t)n_rows * sizeof(unsigned short) + 15) & ~size_t(15)) + 16;
}

}  // namespace s2d
