// grouping.cuh -- stable counting sort of rulebook rows by their 9-bit neighbour-pattern key (see conv_bf2.cu, "Row grouping";
// buckets ordered by descending popcount of the key, i.e. the rows with the most live offsets first):
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

// The prefix over the blocks runs in kGrpSegs independent segments of the block range (a single CTA walking all blocks cost
// 38 us per rulebook of 1.3 M rows):
//   group_scan_kernel   (kGrpSegs / 2 CTAs): counts[blk][b] -> exclusive prefix inside the block's segment; tails[s][b] = rows of
//                       bucket b in segment s
//   group_offsets_kernel (one CTA): tails[s][b] -> start of bucket b in the sorted order + rows of b in the segments before s
// so that the scatter adds counts[blk][key] + tails[segment of blk][key].
constexpr int kGrpSegs = 16;
__host__ __device__ __forceinline__ int group_seg_lo(int s, int nblk) { return (int)((long long)s * nblk / kGrpSegs); }
__device__ __forceinline__ int group_seg_of(int blk, int nblk) {
  int s = (int)((long long)blk * kGrpSegs / nblk);
  while (s + 1 < kGrpSegs && blk >= group_seg_lo(s + 1, nblk)) ++s;
  while (s > 0 && blk < group_seg_lo(s, nblk)) --s;
  return s;
}

static __global__ void __launch_bounds__(1024) group_scan_kernel(int* __restrict__ counts, int nblk, int* __restrict__ tails) {
  const int b = threadIdx.x & (kGrpBuckets - 1), seg = 2 * blockIdx.x + (threadIdx.x >> 9);
  const int lo = group_seg_lo(seg, nblk), hi = group_seg_lo(seg + 1, nblk);
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
  tails[seg * kGrpBuckets + b] = run;
}

static __global__ void __launch_bounds__(kGrpBuckets) group_offsets_kernel(int* __restrict__ tails) {
  __shared__ int s_scan[kGrpBuckets];
  const int b = threadIdx.x;
  int tot[kGrpSegs], total = 0;
#pragma unroll
  for (int s = 0; s < kGrpSegs; ++s) { tot[s] = tails[s * kGrpBuckets + b]; total += tot[s]; }
  // Bucket ORDER: most live offset triples first (ties: smaller key first).  The tile kernel hands its tile groups out
  // dynamically in row order (conv_bf2.cu), so the expensive groups go first and the cheap ones fill the tail.
  int rank = 0;
  const int pc = __popc((unsigned)b);
  for (int o = 0; o < kGrpBuckets; ++o) {
    const int po = __popc((unsigned)o);
    rank += (po > pc || (po == pc && o < b)) ? 1 : 0;
  }
  s_scan[rank] = total;
  __syncthreads();
  for (int off = 1; off < kGrpBuckets; off <<= 1) {
    const int add = b >= off ? s_scan[b - off] : 0;
    __syncthreads();
    s_scan[b] += add;
    __syncthreads();
  }
  int run = s_scan[rank] - total;                        // start of bucket b
#pragma unroll
  for (int s = 0; s < kGrpSegs; ++s) { tails[s * kGrpBuckets + b] = run; run += tot[s]; }
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
    const int pos = counts[(size_t)blockIdx.x * kGrpBuckets + key] + tails[group_seg_of((int)blockIdx.x, nblk) * kGrpBuckets + key] +
                    s_cnt[threadIdx.x >> 5][key] + rank;
    perm[pos] = row;
  }
}

// scan + bucket offsets (two launches)
inline void group_scan(int* counts, int nblk, int* tails, cudaStream_t st) {
  group_scan_kernel<<<kGrpSegs / 2, 1024, 0, st>>>(counts, nblk, tails);
  group_offsets_kernel<<<1, kGrpBuckets, 0, st>>>(tails);
}


inline size_t group_workspace_bytes(int n_rows) {
  const size_t nblk = (size_t)div_up(n_rows > 0 ? n_rows : 1, kGrpRows);
  return (nblk + kGrpSegs) * kGrpBuckets * sizeof(int) + (((size_t)n_rows * sizeof(unsigned short) + 15) & ~size_t(15)) + 16;
}

}  // namespace s2d
