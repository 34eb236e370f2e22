// K5: DCN-v2 cross layer on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), bf16 in / fp32 acc.
//
//   forward   u = x_l W^T + b ;  out = x0 (.) u + x_l                       (W stored [d_out, d_in] like nn.Linear)
//   dgrad     g_x = g_u W + g_out ,  g_u = g_out (.) x0 ;  also emits g_x (.) x0 (the next layer's g_u)
//   wgrad     gW = g_u^T x_l  (K = batch): both operands transposed to K-major by a tiled transpose kernel,
//             then the same GEMM with an fp32 epilogue.
// Bound: bf16 tensor pipe.  FLOPs per launch = 2*B*d*d; the epilogue streams 3-4 [B, d] bf16 tiles, so the layer
// sits near the ridge (AI ~ 300 FLOP/B) — the elementwise work is fused into the GEMM epilogue to stay on it.
//
// Persistent kernel, one CTA per SM walking 128 x 128 output tiles (n fastest, so concurrently running CTAs share
// the A row-panel in L2).  Warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + tcgen05.mma issuer
// (one elected lane), warps 2-5 = epilogue.  Three pipelines:
//   * operand ring: 4 stages of K-major 128B-swizzled tiles (128 rows x 64 bf16 for A and for B) filled by
//     cp.async.bulk.tensor.2d, described to the MMA by shared-memory matrix descriptors, released by tcgen05.commit;
//   * accumulator double buffer: 2 x 128 TMEM columns, so the MMAs of tile i+1 run while tile i is drained;
//   * epilogue staging: the two [128 x 128] bf16 epilogue operands (x0 / x_l, or g_out / x0) are TMA-loaded into
//     shared memory while the MMAs run; the epilogue warps tcgen05.ld the accumulator, read the operands from the
//     swizzled tiles (conflict-free), write both results back in place and one thread TMA-stores them
//     (cp.async.bulk.tensor.2d.global.shared), so every global access of the epilogue is a full coalesced tile.
// Every mbarrier wait is bounded (trap instead of hang).
#include "tcgen05.cuh"
#include "gemm2sm.cuh"

namespace ptrec {

constexpr int kBM = 128, kBN = 128, kBK = 64, kStages = 4, kUmmaK = 16;
constexpr int kEpiWarps = 8;                              // 2 per TMEM lane quadrant (column halves)
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kStageBytesA = kBM * kBK * 2, kStageBytesB = kBN * kBK * 2;
constexpr uint32_t kTmemCols = 256;                      // two 128-column accumulators
constexpr uint32_t kEpiSubTile = kBM * 64 * 2;           // 128 rows x 64 bf16, one 128B-swizzled box (16 KB)
constexpr uint32_t kEpiOperand = 2 * kEpiSubTile;        // [128 x 128] bf16
constexpr size_t kGemmSmem = (size_t)kStages * (kStageBytesA + kStageBytesB) + 2 * kEpiOperand + 1024 /*align*/ +
                             1024 /*barriers + bias slice*/;

enum EpiMode : int { EPI_F32 = 0, EPI_CROSS_FWD = 1, EPI_CROSS_DGRAD = 2 };

struct EpiArgs {
  int mode;
  int64_t ld;                  // row pitch (elements) of the [M, N] bf16 operands / outputs below
  const __nv_bfloat16* p0;     // CROSS_FWD: x0       CROSS_DGRAD: g_out
  const __nv_bfloat16* p1;     // CROSS_FWD: x_l      CROSS_DGRAD: x0
  const float* bias;           // CROSS_FWD: bias[N] (may be null)
  __nv_bfloat16* o0;           // CROSS_FWD: out      CROSS_DGRAD: g_x
  __nv_bfloat16* o1;           // CROSS_FWD: u (null = skip)   CROSS_DGRAD: g_x (.) x0 (null = skip)
  float* of32;                 // EPI_F32: fp32 output [splits][M, ldf] (split-K partials)
  int64_t ldf;
  int splits;                  // EPI_F32 only: K is cut into `splits` ranges, one output slab each (>= 1)
};

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

__device__ __forceinline__ void load_bf16x8(const __nv_bfloat16* p, float* f) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* p, const float* f) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}

// ---- the GEMM: C[M, N] = A[M, K] * B[N, K]^T, fused epilogue ---------------------------------------------------
struct GemmMaps {
  CUtensorMap a, b;        // mainloop operands
  CUtensorMap p0, p1;      // epilogue operands (loaded)
  CUtensorMap o0, o1;      // epilogue results (stored)
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ GemmMaps maps, int M, int N, int K, EpiArgs ep) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = smem + (size_t)kStages * kStageBytesA;
  unsigned char* sE0 = smem + (size_t)kStages * (kStageBytesA + kStageBytesB);  // operand 0 / result 0
  unsigned char* sE1 = sE0 + kEpiOperand;                                       // operand 1 / result 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(sE1 + kEpiOperand);
  uint64_t* full = bars;                      // [kStages]
  uint64_t* empty = bars + kStages;           // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;    // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* epi_full = acc_empty + 2;         // [1]
  uint64_t* epi_empty = epi_full + 1;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_empty + 1);
  float* s_bias = reinterpret_cast<float*>(bars + 16);       // [kBN] bias slice of the current tile (16-byte aligned)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + kBN - 1) / kBN, tiles_m = (M + kBM - 1) / kBM;
  const int splits = ep.mode == EPI_F32 ? max(ep.splits, 1) : 1;
  const int mn_tiles = tiles_n * tiles_m;
  const int n_tiles = mn_tiles * splits;
  const int total_kb = (K + kBK - 1) / kBK;
  const int kb_per_split = (total_kb + splits - 1) / splits;
  const bool staged_epi = ep.mode != EPI_F32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiWarps);  // one arrival per epilogue warp
    }
    mbar_init(epi_full, 1);
    mbar_init(epi_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b) : "memory");
      int kbg = 0;  // global k-block counter (ring position)
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int mn = tile % mn_tiles, split = tile / mn_tiles;
        const int m0 = (mn / tiles_n) * kBM, n0 = (mn % tiles_n) * kBN;
        const int kb0 = split * kb_per_split, num_kb = max(0, min(total_kb - kb0, kb_per_split));
        const int epi_at = num_kb / 2;  // epilogue operands are requested half-way through the K loop
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % kStages;
          mbar_wait(&empty[s], ((kbg / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], kStageBytesA + kStageBytesB);
          tma_load_2d(sA + (size_t)s * kStageBytesA, &maps.a, (kb0 + kb) * kBK, m0, &full[s]);
          tma_load_2d(sB + (size_t)s * kStageBytesB, &maps.b, (kb0 + kb) * kBK, n0, &full[s]);
          if (staged_epi && kb == epi_at) {
            mbar_wait(epi_empty, (it & 1) ^ 1);  // previous tile's results have left shared memory
            mbar_arrive_expect_tx(epi_full, 2 * kEpiOperand);
            tma_load_2d(sE0, &maps.p0, n0, m0, epi_full);
            tma_load_2d(sE0 + kEpiSubTile, &maps.p0, n0 + 64, m0, epi_full);
            tma_load_2d(sE1, &maps.p1, n0, m0, epi_full);
            tma_load_2d(sE1 + kEpiSubTile, &maps.p1, n0 + 64, m0, epi_full);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int kbg = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kBN);
        const int kb0 = (tile / mn_tiles) * kb_per_split, num_kb = max(0, min(total_kb - kb0, kb_per_split));
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % kStages;
          mbar_wait(&full[s], (kbg / kStages) & 1);
          tcgen05_fence_after();
          const uint64_t adesc = make_sw128_desc(sA + (size_t)s * kStageBytesA);
          const uint64_t bdesc = make_sw128_desc(sB + (size_t)s * kStageBytesB);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // +32 bytes per K=16 slice inside the 128-byte swizzled row (start-address field is in 16-byte units)
            umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[s]);      // frees the stage once the MMAs that read it retire
        }
        umma_commit(&acc_full[buf]);   // accumulator complete
      }
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), +32); two warps share a quadrant and split the columns
    const int e = warp - 2;            // 0 .. kEpiWarps-1
    const int q = warp & 3;
    const int half = e >> 2;           // which 64-column half of the tile this warp drains
    const int r = q * 32 + lane;       // row inside the tile
    const int et = threadIdx.x - 64;   // 0 .. 32*kEpiWarps-1
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int mn = tile % mn_tiles, split = tile / mn_tiles;
      const int m0 = (mn / tiles_n) * kBM, n0 = (mn % tiles_n) * kBN;
      const int buf = it & 1;
      if (ep.mode == EPI_CROSS_FWD) {
        if (et < kBN) s_bias[et] = (ep.bias && n0 + et < N) ? ep.bias[n0 + et] : 0.f;
        asm volatile("bar.sync 2, %0;" ::"n"(32 * kEpiWarps) : "memory");
      }
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tcgen05_fence_after();
      if (staged_epi) mbar_wait(epi_full, it & 1);
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kBN + half * 64);
      const int row = m0 + r;
      uint32_t v[64];
      tmem_ld32(tacc, v);
      tmem_ld32(tacc + 32, v + 32);
      // accumulator columns are in registers: hand the TMEM buffer back to the MMA warp right away
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (!staged_epi) {
        const int col0 = n0 + half * 64;
        if (row < M && col0 < N) {
          float* o = ep.of32 + ((int64_t)split * M + row) * ep.ldf + col0;
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            if (col0 + j < N) {
              *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                               __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
          }
        }
      } else {
        unsigned char* e0 = sE0 + (size_t)half * kEpiSubTile + (size_t)r * 128u;
        unsigned char* e1 = sE1 + (size_t)half * kEpiSubTile + (size_t)r * 128u;
#pragma unroll
        for (int g = 0; g < 8; ++g) {  // 8-column groups of this warp's 64-column half
          const uint32_t off = (uint32_t)((g ^ (r & 7)) << 4);
          float a[8], b[8], r0[8], r1[8];
          load_bf16x8(reinterpret_cast<const __nv_bfloat16*>(e0 + off), a);
          load_bf16x8(reinterpret_cast<const __nv_bfloat16*>(e1 + off), b);
          if (ep.mode == EPI_CROSS_FWD) {
            const float4 b0 = *reinterpret_cast<const float4*>(&s_bias[half * 64 + g * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&s_bias[half * 64 + g * 8 + 4]);
            const float bias8[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float u = __uint_as_float(v[g * 8 + i]) + bias8[i];
              r1[i] = u;                 // u          -> result 1
              r0[i] = a[i] * u + b[i];   // x0*u + x_l -> result 0
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float gx = __uint_as_float(v[g * 8 + i]) + a[i];  // g_u W + g_out -> result 0
              r0[i] = gx;
              r1[i] = gx * b[i];                                      // g_x (.) x0    -> result 1
            }
          }
          store_bf16x8(reinterpret_cast<__nv_bfloat16*>(e0 + off), r0);
          store_bf16x8(reinterpret_cast<__nv_bfloat16*>(e1 + off), r1);
        }
        fence_proxy_async();                                              // generic-proxy writes -> visible to the TMA store
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");  // all epilogue warps
        if (et == 0) {
          tma_store_2d(&maps.o0, sE0, n0, m0);
          if (n0 + 64 < N) tma_store_2d(&maps.o0, sE0 + kEpiSubTile, n0 + 64, m0);
          if (ep.o1) {
            tma_store_2d(&maps.o1, sE1, n0, m0);
            if (n0 + 64 < N) tma_store_2d(&maps.o1, sE1 + kEpiSubTile, n0 + 64, m0);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may be overwritten
          mbar_arrive(epi_empty);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// [R, C] bf16 (pitch ld_in) -> [C, R] bf16 (pitch ld_out), 64x64 tiles through shared memory, 4-byte (bf16 x 2)
// global accesses on both sides: a warp reads 128 contiguous bytes of an input row and writes 128 contiguous bytes of an
// output row (the 32x32 / 2-byte version moved 64 B per warp access and ran at 28 % of the HBM rate: 60 us per
// [32768, 848] operand, profiles/r2_step_kernels_dcn_before.txt).  ld_in, ld_out even; R, C arbitrary.
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld_in, int R, int C,
                      __nv_bfloat16* __restrict__ out, int64_t ld_out) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;  // 8 warps
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
  for (int i = wy; i < 64; i += 8) {
    const int r = r0 + i, c = c0 + 2 * lane;
    __nv_bfloat162 v = __halves2bfloat162(zero, zero);
    if (r < R) {
      if (c + 1 < C) {
        v = *reinterpret_cast<const __nv_bfloat162*>(in + (int64_t)r * ld_in + c);
      } else if (c < C) {
        v = __halves2bfloat162(in[(int64_t)r * ld_in + c], zero);
      }
    }
    *reinterpret_cast<__nv_bfloat162*>(&tile[i][2 * lane]) = v;
  }
  __syncthreads();
  for (int i = wy; i < 64; i += 8) {
    const int c = c0 + i, r = r0 + 2 * lane;  // output row c, output columns r, r + 1
    if (c >= C) continue;
    const __nv_bfloat162 v = __halves2bfloat162(tile[2 * lane][i], tile[2 * lane + 1][i]);
    if (r + 1 < R) {
      *reinterpret_cast<__nv_bfloat162*>(out + (int64_t)c * ld_out + r) = v;
    } else if (r < R) {
      out[(int64_t)c * ld_out + r] = v.x;
    }
  }
}

// out[e] = sum_s partial[s][e]  (fixed order: deterministic split-K)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t n, float* __restrict__ out) {
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= n) return;
  float4 acc = *reinterpret_cast<const float4*>(partial + e);
  for (int s = 1; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(partial + (int64_t)s * n + e);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(out + e) = acc;
}

// ---- host side ---------------------------------------------------------------------------------------------------
// 2-D bf16 row-major [rows, cols] with pitch ld (elements); box = 64 cols x 128 rows, 128B swizzle
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn enc = get_encode();
  PTREC_CHECK_ARG(enc != nullptr, PTREC_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PTREC_CHECK_ARG(r == CUDA_SUCCESS, PTREC_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return PTREC_OK;
}

static int launch_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const EpiArgs& ep,
                       cudaStream_t st) {
  PTREC_CHECK_ARG(aligned16(A) && aligned16(B) && lda % 8 == 0 && ldb % 8 == 0, PTREC_EALIGN,
                  "gemm: operands must be 16-byte aligned with pitches that are multiples of 8 elements");
  PTREC_CHECK_ARG(M > 0 && N > 0 && K > 0 && N % 8 == 0, PTREC_EINVAL, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = make_map(&maps.a, A, M, K, lda);
  if (rc != PTREC_OK) return rc;
  rc = make_map(&maps.b, B, N, K, ldb);
  if (rc != PTREC_OK) return rc;
  if (ep.mode != EPI_F32) {
    // epilogue operands / results are [M, N] bf16 with pitch ep.ld; box = 64 columns x 128 rows, 128B swizzle
    if ((rc = make_map(&maps.p0, ep.p0, M, N, ep.ld)) != PTREC_OK) return rc;
    if ((rc = make_map(&maps.p1, ep.p1, M, N, ep.ld)) != PTREC_OK) return rc;
    if ((rc = make_map(&maps.o0, ep.o0, M, N, ep.ld)) != PTREC_OK) return rc;
    if ((rc = make_map(&maps.o1, ep.o1 ? ep.o1 : ep.o0, M, N, ep.ld)) != PTREC_OK) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    PTREC_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ceil_div(N, kBN) * ceil_div(M, kBM) * (ep.mode == EPI_F32 && ep.splits > 1 ? ep.splits : 1);
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  gemm_bf16_tn_kernel<<<grid, kGemmThreads, kGemmSmem, st>>>(maps, M, N, K, ep);
  PTREC_LAUNCH_CHECK("gemm_bf16_tn_kernel");
  return PTREC_OK;
}

}  // namespace ptrec

using namespace ptrec;

// Round 2: the cross GEMMs run on the CTA-pair kernel of tc_linear.cu (one bf16 operand plane, 256 x 256 pair tiles,
// 16 epilogue warps, TMA stores, MN-major operands for the weight gradient: no transposed copies).  The 128 x 128
// single-CTA kernel of this file stays selectable (identical results up to summation order; the tests run both).
static int g_dcn_2sm = 1;
extern "C" void ptrec_set_dcn_2sm(int32_t enabled) { g_dcn_2sm = enabled ? 1 : 0; }
extern "C" int32_t ptrec_dcn_2sm_enabled(void) { return g_dcn_2sm; }
extern "C" int32_t ptrec_tc_gemm_split3_default_splits(int64_t M, int64_t N, int64_t K);

static int check_cross(const void* a, const void* b, const void* c, int64_t B, int32_t d, int64_t ld) {
  PTREC_CHECK_ARG(a && b && c, PTREC_EINVAL, "dcn_cross: null pointer");
  PTREC_CHECK_ARG(B > 0 && B < (int64_t)0x7fffffff && d >= 8 && d % 8 == 0 && ld >= d && ld % 8 == 0, PTREC_EINVAL,
                  "dcn_cross: need d %% 8 == 0 (pad the concat once), ld %% 8 == 0; got B=%lld d=%d ld=%lld",
                  (long long)B, d, (long long)ld);
  return PTREC_OK;
}

extern "C" int ptrec_dcn_cross_fwd(const void* x_l, const void* x0, const void* weight, const float* bias, int64_t B,
                                   int32_t d, int64_t ld, void* out, void* u_out, void* stream) {
  int rc = check_cross(x_l, x0, weight, B, d, ld);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(out && aligned16(out) && aligned16(x0) && (!u_out || aligned16(u_out)), PTREC_EALIGN, "dcn_cross_fwd: alignment");
  if (g_dcn_2sm) {
    Bf16Gemm g{};
    g.mn_major = false; g.A = x_l; g.lda = ld; g.B = weight; g.ldb = d; g.M = B; g.N = d; g.K = d; g.bias = bias;
    g.dcn_mode = 1; g.p0 = x0; g.p1 = x_l; g.pld = ld; g.o0 = out; g.o1 = u_out; g.splits = 1;
    return gemm_bf16_2sm(g, (cudaStream_t)stream);
  }
  EpiArgs ep{};
  ep.mode = EPI_CROSS_FWD;
  ep.ld = ld;
  ep.p0 = reinterpret_cast<const __nv_bfloat16*>(x0);
  ep.p1 = reinterpret_cast<const __nv_bfloat16*>(x_l);
  ep.bias = bias;
  ep.o0 = reinterpret_cast<__nv_bfloat16*>(out);
  ep.o1 = reinterpret_cast<__nv_bfloat16*>(u_out);
  return launch_gemm(x_l, ld, weight, d, (int)B, d, d, ep, (cudaStream_t)stream);
}

extern "C" int ptrec_dcn_cross_dgrad(const void* g_u, const void* weight_t, const void* g_out, const void* x0, int64_t B,
                                     int32_t d, int64_t ld, void* g_x, void* g_u_prev, void* stream) {
  int rc = check_cross(g_u, weight_t, g_out, B, d, ld);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(x0 && g_x && aligned16(g_x) && aligned16(x0) && aligned16(g_out) && (!g_u_prev || aligned16(g_u_prev)),
                  PTREC_EALIGN, "dcn_cross_dgrad: alignment");
  if (g_dcn_2sm) {
    Bf16Gemm g{};
    g.mn_major = false; g.A = g_u; g.lda = ld; g.B = weight_t; g.ldb = d; g.M = B; g.N = d; g.K = d; g.bias = nullptr;
    g.dcn_mode = 2; g.p0 = g_out; g.p1 = x0; g.pld = ld; g.o0 = g_x; g.o1 = g_u_prev; g.splits = 1;
    return gemm_bf16_2sm(g, (cudaStream_t)stream);
  }
  EpiArgs ep{};
  ep.mode = EPI_CROSS_DGRAD;
  ep.ld = ld;
  ep.p0 = reinterpret_cast<const __nv_bfloat16*>(g_out);
  ep.p1 = reinterpret_cast<const __nv_bfloat16*>(x0);
  ep.o0 = reinterpret_cast<__nv_bfloat16*>(g_x);
  ep.o1 = reinterpret_cast<__nv_bfloat16*>(g_u_prev);
  return launch_gemm(g_u, ld, weight_t, d, (int)B, d, d, ep, (cudaStream_t)stream);
}

static int wgrad_splits(int32_t d) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t mn = ceil_div(d, kBN) * ceil_div(d, kBM);
  int s = (int)(sms / mn);
  return s < 1 ? 1 : (s > 16 ? 16 : s);
}

extern "C" size_t ptrec_dcn_cross_wgrad_workspace_bytes(int64_t B, int32_t d) {
  const size_t ldt = align_up((size_t)B, 8);
  return 2 * align_up((size_t)d * ldt * 2, 256) + align_up((size_t)16 * d * d * 4, 256) + 256;
}

extern "C" int ptrec_dcn_cross_wgrad(const void* g_u, const void* x_l, int64_t B, int32_t d, int64_t ld, float* grad_w,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_cross(g_u, x_l, grad_w, B, d, ld);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(workspace && workspace_bytes >= ptrec_dcn_cross_wgrad_workspace_bytes(B, d), PTREC_EWORKSPACE,
                  "dcn_cross_wgrad: workspace too small");
  PTREC_CHECK_ARG(aligned16(grad_w) && d % 4 == 0, PTREC_EALIGN, "dcn_cross_wgrad: grad_w alignment");
  cudaStream_t st = (cudaStream_t)stream;
  if (g_dcn_2sm) {
    // gW[i, j] = sum_b g_u[b, i] x_l[b, j]: both stored matrices are read MN-major (the batch is their row index)
    Bf16Gemm g{};
    g.mn_major = true; g.A = g_u; g.lda = ld; g.B = x_l; g.ldb = ld; g.M = d; g.N = d; g.K = B;
    g.dcn_mode = 0; g.out = grad_w; g.ldo = d;
    g.splits = std::min<int>(16, std::max<int>(1, ptrec_tc_gemm_split3_default_splits(d, d, B)));
    g.workspace = workspace; g.workspace_bytes = workspace_bytes;
    return gemm_bf16_2sm(g, st);
  }
  const int64_t ldt = (int64_t)align_up((size_t)B, 8);
  __nv_bfloat16* gt = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* xt = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(workspace) +
                                                       align_up((size_t)d * ldt * 2, 256));
  dim3 tb(256), tg((unsigned)ceil_div(d, 64), (unsigned)ceil_div(B, 64));
  transpose_bf16_kernel<<<tg, tb, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g_u), ld, (int)B, d, gt, ldt);
  PTREC_LAUNCH_CHECK("transpose_bf16_kernel");
  transpose_bf16_kernel<<<tg, tb, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x_l), ld, (int)B, d, xt, ldt);
  PTREC_LAUNCH_CHECK("transpose_bf16_kernel");
  float* partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) +
                                            2 * align_up((size_t)d * ldt * 2, 256));
  int splits = wgrad_splits(d);
  {  // every split must own at least one 64-wide K block (an empty split would publish a stale accumulator)
    const int total_kb = (int)ceil_div(B, kBK);
    if (splits > total_kb) splits = total_kb;
    const int per = (int)ceil_div(total_kb, splits);
    splits = (int)ceil_div(total_kb, per);
  }
  EpiArgs ep{};
  ep.mode = EPI_F32;
  ep.of32 = splits > 1 ? partial : grad_w;
  ep.ldf = d;
  ep.splits = splits;
  // gW[i, j] = sum_b g_u[b, i] * x_l[b, j]  ->  A = g_u^T [d, B], B = x_l^T [d, B], both K(=batch)-major;
  // the batch (K) is cut into `splits` ranges so that all SMs have a tile; partials are summed in a fixed order
  int rc2 = launch_gemm(gt, ldt, xt, ldt, d, d, (int)B, ep, st);
  if (rc2 != PTREC_OK || splits == 1) return rc2;
  const int64_t n = (int64_t)d * d;
  splitk_reduce_kernel<<<(unsigned)ceil_div(n / 4, 256), 256, 0, st>>>(partial, splits, n, grad_w);
  PTREC_LAUNCH_CHECK("splitk_reduce_kernel");
  return PTREC_OK;
}
