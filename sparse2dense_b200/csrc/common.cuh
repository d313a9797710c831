// common.cuh -- shared helpers of libs2d_b200.so: error reporting, launch sizing, device scan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/s2d_b200.h"

namespace s2d {

// ---- error reporting (thread-local text, integer status; the library never exits) ----------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define S2D_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t _e = (call);                                                    \
    if (_e != cudaSuccess) return s2d::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define S2D_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) return s2d::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

#define S2D_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      s2d::set_error(__VA_ARGS__);  \
      return S2D_ERR_INVALID;       \
    }                               \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// number of kernels this library has launched in this process (s2d_kernel_launches)
void count_launches(int n);

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (256 B aligned slices).
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <class T>
  T* take(size_t n) {
    T* r = reinterpret_cast<T*>(base + off);
    off += align_up(n * sizeof(T), 256);
    return r;
  }
};

// ---- device-wide exclusive scan of int values produced by a functor -------------------------
// out[i] = sum_{j<i} f(j) for i in [0,n); out[n] = total.  Three launches: per-block reduce,
// single-block scan of the block sums, per-block scan + offset.  HBM-bound: reads f twice.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

inline int scan_num_blocks(long long n) { return div_up(n > 0 ? n : 1, kScanTile); }

__device__ __forceinline__ int warp_inclusive_scan(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) >= d) v += t;
  }
  return v;
}

// inclusive block scan of one value per thread; returns the inclusive prefix, *total = block sum
template <int THREADS>
__device__ __forceinline__ int block_inclusive_scan(int v, int* total) {
  __shared__ int warp_sums[THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = warp_inclusive_scan(v);
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < THREADS / 32 ? warp_sums[lane] : 0;
    s = warp_inclusive_scan(s);
    if (lane < THREADS / 32) warp_sums[lane] = s;
  }
  __syncthreads();
  if (w > 0) inc += warp_sums[w - 1];
  *total = warp_sums[THREADS / 32 - 1];
  __syncthreads();
  return inc;
}

template <class F>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(F f, long long n, int* block_sums) {
  const long long base = (long long)blockIdx.x * kScanTile;
  int s = 0;
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) {
    long long i = base + it * kScanThreads + threadIdx.x;  // coalesced
    if (i < n) s += f(i);
  }
  int total;
  block_inclusive_scan<kScanThreads>(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// one block: exclusive scan of block_sums[0..nb) in place, total -> block_sums[nb]
__global__ void scan_block_sums_kernel(int* block_sums, int nb);

template <class F>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(F f, long long n, const int* block_sums,
                                                                  int* out) {
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) {
    long long i = base + it;
    v[it] = i < n ? f(i) : 0;
    s += v[it];
  }
  int total;
  int inc = block_inclusive_scan<kScanThreads>(s, &total);
  int run = block_sums[blockIdx.x] + inc - s;
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) {
    long long i = base + it;
    if (i < n) out[i] = run;
    run += v[it];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) out[n] = block_sums[gridDim.x];
}

// block_sums needs scan_num_blocks(n)+1 ints; out needs n+1 ints.
template <class F>
int exclusive_scan(F f, long long n, int* out, int* block_sums, cudaStream_t st) {
  const int nb = scan_num_blocks(n);
  scan_reduce_kernel<F><<<nb, kScanThreads, 0, st>>>(f, n, block_sums);
  scan_block_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nb);
  scan_apply_kernel<F><<<nb, kScanThreads, 0, st>>>(f, n, block_sums, out);
  S2D_LAUNCH_CHECK();
  return S2D_OK;
}

// ---- occupancy-bitmap coordinate index (see s2d_b200.h) --------------------------------------
// words[w] bit i set <=> flattened voxel 32*w+i is active; prefix[w] = #active voxels before word w;
// perm maps ascending rank -> tensor row (identity for tensors produced by s2d_rulebook_sparse).
struct GridIndexView {
  const uint32_t* words;
  const int* prefix;
  const int* perm;
  long long n_words;

  __device__ __forceinline__ int rank(long long lin) const {
    const uint32_t w = __ldg(words + (lin >> 5));
    const uint32_t bit = 1u << (lin & 31);
    if (!(w & bit)) return -1;
    return __ldg(prefix + (lin >> 5)) + __popc(w & (bit - 1));
  }
  __device__ __forceinline__ int lookup(long long lin) const {
    const int r = rank(lin);
    return r < 0 ? -1 : __ldg(perm + r);
  }
};

struct GridIndexLayout {
  long long n_words;
  size_t words_off, prefix_off, perm_off, sums_off, total;
};
GridIndexLayout grid_index_layout(int batch, const int* shape, int n_rows_capacity);

struct GridIndexPtrs {
  uint32_t* words;
  int* prefix;
  int* perm;
  int* sums;
  long long n_words;
  GridIndexView view() const { return GridIndexView{words, prefix, perm, n_words}; }
};
GridIndexPtrs grid_index_ptrs(const void* index, const GridIndexLayout& L);

}  // namespace s2d
