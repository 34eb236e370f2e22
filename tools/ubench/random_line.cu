// Microbenchmark: how many RANDOM 128-byte DRAM lines per second can a B200 read, and read-modify-write?
// This is the ceiling of K1 (64-byte rows: one line read per lookup) and K2b (weight | Adagrad-sum row = one line
// read + the same line written back per unique row) at D = 16, where bytes / s is the wrong yardstick because every
// access pays a full line whatever it uses (profiles/r1_ubench_dram_granularity.csv).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o random_line random_line.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

// MODE 0: read 64 B of the line (4 lanes x 16 B), result kept in a register (no output traffic)
// MODE 1: read 64 B, write 64 B to a dense output row (the gather)
// MODE 2: read the whole 128-B line (8 lanes x 16 B), no output
// MODE 3: read-modify-write the whole line (8 lanes)
// MODE 4: RMW the line + read a dense 64-B "gradient" row per line (the fused update)
template <int MODE, int U>
__global__ void __launch_bounds__(256) k(float* __restrict__ tab, const int* __restrict__ idx, int n,
                                         const float* __restrict__ dense_in, float* __restrict__ dense_out, float* sink) {
  constexpr int LPR = (MODE <= 1) ? 4 : 8;
  const int g = (blockIdx.x * 256 + threadIdx.x) / LPR;     // row group
  const int lane = threadIdx.x % LPR;
  const int groups = gridDim.x * 256 / LPR;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int r0 = g; r0 < n; r0 += groups * U) {
    float4 v[U];
    float4 d[U];
    float* p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + u * groups;
      p[u] = nullptr;
      if (r < n) {
        p[u] = tab + (size_t)idx[r] * 32 + lane * 4;
        v[u] = *reinterpret_cast<const float4*>(p[u]);
        if (MODE == 4) d[u] = __ldg(reinterpret_cast<const float4*>(dense_in + (size_t)r * 16 + (lane & 3) * 4));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + u * groups;
      if (p[u] == nullptr) continue;
      if (MODE == 0 || MODE == 2) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      if (MODE == 1) *reinterpret_cast<float4*>(dense_out + (size_t)r * 16 + lane * 4) = v[u];
      if (MODE == 3) { v[u].x += 1.f; v[u].y += 1.f; v[u].z += 1.f; v[u].w += 1.f; *reinterpret_cast<float4*>(p[u]) = v[u]; }
      if (MODE == 4) { v[u].x += d[u].x; v[u].y += d[u].y; v[u].z += d[u].z; v[u].w += d[u].w; *reinterpret_cast<float4*>(p[u]) = v[u]; }
    }
  }
  if (acc.x == 123.456f) *sink = acc.y + acc.z + acc.w;
}

template <int MODE, int U>
void run(const char* name, float* tab, const int* idx, int n, const float* din, float* dout, float* sink) {
  constexpr int LPR = (MODE <= 1) ? 4 : 8;
  const long long groups_needed = ((long long)n + U - 1) / U;
  int blocks = (int)((groups_needed * LPR + 255) / 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 6; ++rep) {
    // a different slice of the index array per repetition: nothing is warm in L2 (table >> L2)
    const int* ix = idx + (size_t)rep * n;
    cudaEventRecord(e0);
    k<MODE, U><<<blocks, 256>>>(tab, ix, n, din, dout, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  printf("%-34s n=%8d U=%d  %8.1f us  %6.2f G lines/s  %7.1f GB/s of 128-B lines%s\n", name, n, U, best * 1e3,
         n / best / 1e6, (double)n * 128 * ((MODE >= 3) ? 2 : 1) / best / 1e6, (MODE >= 3) ? " (read + write)" : "");
}

int main() {
  const size_t lines = (size_t)1 << 25;  // 32 Mi lines x 128 B = 4 GiB
  float* tab; cudaMalloc(&tab, lines * 128); cudaMemset(tab, 0, lines * 128);
  const int nmax = 1 << 22;
  std::vector<int> h((size_t)nmax * 6);
  uint64_t s = 88172645463325252ull;
  for (size_t i = 0; i < h.size(); ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % lines); }
  int* idx; cudaMalloc(&idx, h.size() * 4); cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  float *din, *dout, *sink;
  cudaMalloc(&din, (size_t)nmax * 64); cudaMemset(din, 0, (size_t)nmax * 64);
  cudaMalloc(&dout, (size_t)nmax * 64); cudaMalloc(&sink, 4);
  for (int n : {425984, nmax}) {
    run<0, 1>("read 64 B of a line, no output", tab, idx, n, din, dout, sink);
    run<0, 4>("read 64 B of a line, no output", tab, idx, n, din, dout, sink);
    run<0, 8>("read 64 B of a line, no output", tab, idx, n, din, dout, sink);
    run<1, 1>("gather: read 64 B, write 64 B dense", tab, idx, n, din, dout, sink);
    run<1, 4>("gather: read 64 B, write 64 B dense", tab, idx, n, din, dout, sink);
    run<1, 8>("gather: read 64 B, write 64 B dense", tab, idx, n, din, dout, sink);
    run<2, 1>("read whole line, no output", tab, idx, n, din, dout, sink);
    run<2, 4>("read whole line, no output", tab, idx, n, din, dout, sink);
    run<3, 1>("RMW whole line", tab, idx, n, din, dout, sink);
    run<3, 4>("RMW whole line", tab, idx, n, din, dout, sink);
    run<3, 8>("RMW whole line", tab, idx, n, din, dout, sink);
    run<4, 1>("RMW line + dense 64-B read (update)", tab, idx, n, din, dout, sink);
    run<4, 4>("RMW line + dense 64-B read (update)", tab, idx, n, din, dout, sink);
    run<4, 8>("RMW line + dense 64-B read (update)", tab, idx, n, din, dout, sink);
  }
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
