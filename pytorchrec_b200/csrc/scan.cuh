// Device-wide exclusive scan in three launches (tile sums -> top scan -> apply).
// No inter-CTA spinning: every dependency is a kernel boundary, so nothing can hang the GPU.
#pragma once
#include "common.cuh"

namespace ptrec {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048 consecutive elements per CTA

#ifdef __CUDACC__

__device__ __forceinline__ int warp_inclusive_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Exclusive scan of one int per thread over the CTA (blockDim.x <= 1024, multiple of 32).
// s_warp: >= 33 ints of shared memory.  Returns the exclusive prefix; *total = CTA sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  int inc = warp_inclusive_scan(v);
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarp ? s_warp[lane] : 0;
    int winc = warp_inclusive_scan(w);
    s_warp[lane] = winc - w;  // exclusive prefix of warp sums
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  int res = inc - v + s_warp[warp];
  *total = s_warp[32];
  __syncthreads();  // s_warp may be reused by the caller
  return res;
}

template <class In>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(In in, int64_t n, int* tile_sums) {
  __shared__ int s_warp[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    if (i < n) s += in(i);
  }
  int total;
  block_exclusive_scan(s, s_warp, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single CTA: in-place exclusive scan of tile_sums[0..n_tiles); total -> *total_out (may be null)
static __global__ void __launch_bounds__(1024) scan_top_kernel(int* tile_sums, int n_tiles, int* total_out) {
  __shared__ int s_warp[33];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_tiles ? tile_sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, s_warp, &total);
    const int carry = s_carry;
    if (i < n_tiles) tile_sums[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out != nullptr) *total_out = s_carry;
}

template <class In, class Out>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(In in, Out out, int64_t n, const int* tile_sums) {
  __shared__ int s_warp[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    v[k] = (i < n) ? in(i) : 0;
    s += v[k];
  }
  int total;
  int prefix = block_exclusive_scan(s, s_warp, &total) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    if (i < n) out(i, prefix, v[k]);
    prefix += v[k];
  }
}

#endif  // __CUDACC__

inline int scan_num_tiles(int64_t n) { return (int)((n + kScanTile - 1) / kScanTile); }

}  // namespace ptrec
