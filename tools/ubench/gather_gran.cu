// Microbenchmark: DRAM bytes fetched per random row read on B200, by row size and load flavour.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_gran gather_gran.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void gather_rows(const float* __restrict__ tab, const int* __restrict__ idx, int n, int row_floats, float* out) {
  // one 4-lane group per row when row_floats==16; generic: lanes = row_floats/4
  const int lanes = row_floats / 4;
  const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / lanes;
  const int lane = (blockIdx.x * blockDim.x + threadIdx.x) % lanes;
  if (gid >= n) return;
  const float* p = tab + (size_t)idx[gid] * row_floats + lane * 4;
  float4 r;
  if (MODE == 0) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  } else if (MODE == 1) {
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  } else if (MODE == 2) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  } else if (MODE == 3) {
    asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  } else {
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  }
  reinterpret_cast<float4*>(out)[(size_t)gid * lanes + lane] = r;
}

// TMA 1-D bulk copy of whole rows into shared memory, then coalesced store
__global__ void gather_rows_bulk(const float* __restrict__ tab, const int* __restrict__ idx, int n, int row_floats, float* out) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t bar;
  const int rows_per_cta = blockDim.x;  // one row per thread
  const int base = blockIdx.x * rows_per_cta;
  const int cnt = min(rows_per_cta, n - base);
  if (cnt <= 0) return;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const uint32_t bytes = row_floats * 4;
  if (threadIdx.x < cnt) {
    const float* src = tab + (size_t)idx[base + threadIdx.x] * row_floats;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
        smem_u32(sm + threadIdx.x * row_floats)), "l"(src), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
  }
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes * cnt) : "memory");
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  }
  const int total4 = cnt * row_floats / 4;
  for (int i = threadIdx.x; i < total4; i += blockDim.x)
    reinterpret_cast<float4*>(out + (size_t)base * row_floats)[i] = reinterpret_cast<float4*>(sm)[i];
}

int main() {
  const size_t table_bytes = 2ull << 30;
  const int n = 1 << 21;
  float* tab; cudaMalloc(&tab, table_bytes); cudaMemset(tab, 1, table_bytes);
  int* idx; cudaMalloc(&idx, n * sizeof(int));
  float* out; cudaMalloc(&out, (size_t)n * 256);
  int* h = (int*)malloc(n * sizeof(int));
  for (int rf : {8, 16, 32, 64}) {
    const size_t rows = table_bytes / (rf * 4);
    uint64_t s = 88172645463325252ull;
    for (int i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % rows); }
    cudaMemcpy(idx, h, n * sizeof(int), cudaMemcpyHostToDevice);
    const int lanes = rf / 4;
    const int threads = 256, blocks = (int)(((size_t)n * lanes + threads - 1) / threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
#define RUN(MODE) \
    gather_rows<MODE><<<blocks, threads>>>(tab, idx, n, rf, out); \
    cudaEventRecord(e0); gather_rows<MODE><<<blocks, threads>>>(tab, idx, n, rf, out); cudaEventRecord(e1); cudaEventSynchronize(e1); \
    cudaEventElapsedTime(&ms, e0, e1); printf("row %3d B mode %d: %.1f us  %.0f GB/s (rows only)\n", rf * 4, MODE, ms * 1e3, (double)n * rf * 4 / ms / 1e6);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
    const int rpc = 128;
    gather_rows_bulk<<<(n + rpc - 1) / rpc, rpc, rpc * rf * 4>>>(tab, idx, n, rf, out);
    cudaEventRecord(e0); gather_rows_bulk<<<(n + rpc - 1) / rpc, rpc, rpc * rf * 4>>>(tab, idx, n, rf, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("row %3d B bulk  : %.1f us  %.0f GB/s (rows only)\n", rf * 4, ms * 1e3, (double)n * rf * 4 / ms / 1e6);
  }
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
