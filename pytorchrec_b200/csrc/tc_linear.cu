// K6: fp32-faithful Linear layers of the DNN tower on the tcgen05 tensor cores (reference: torchrec/model/layer/
// Dense.py:4-24, MLP.py:8-23 — nn.Linear -> ReLU).  The reference computes these GEMMs in fp32; the north star
// asks for 1e-5 relative agreement, which rules out TF32 / plain bf16.  Every fp32 operand is therefore split
// EXACTLY into three bf16 planes
//        x = x0 + x1 + x2,   x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1)      (3 x 8 = 24 mantissa bits)
// and the product is accumulated in fp32 (TMEM) from the six plane pairs whose weight is >= 2^-16 relative:
//        A B^T ~= A0B2 + A2B0 + A1B1 + A0B1 + A1B0 + A0B0          (dropped pairs are <= 2^-24 relative)
// so the result carries fp32-level error (measured against fp64 in tests) at 6 bf16 MMAs per product instead of the
// fp32 SIMT pipe: ~10x the arithmetic rate of the cuBLAS sgemm the stock nn.Linear falls back to.
//
//   split_kernel<NP>     fp32 [R, C] -> planes [NP, R, Cp] and/or transposed planes [NP, C, Rp]; optional ReLU-backward
//                        mask (g * (y > 0)) and column sums (bias gradient) fused in the same pass.       bound: HBM
//   gemm_split3_kernel   C[M, N] = sum over the plane pairs of A_i[M, K] B_j[N, K]^T, persistent 128 x 256 tiles;
//   gemm_split3_2sm_...  warp 0 TMA producer (3-D maps: k, row, plane), warp 1 TMEM alloc + tcgen05.mma issuer (every
//                        plane tile is reused by 2-3 MMAs), 8 epilogue warps drain the accumulator pair: + bias, ReLU,
//                        fp32 store; or split-K partials (wgrad).  bound: tensor pipe fed through shared memory.
//
// The kernels are templates on the number of operand planes NP:
//   NP = 3  bf16 x 3 (above): exact split, 6 MMAs per product, FLOPs per launch = 12 * M * N * K.
//   NP = 2  fp16 x 2 (default, split3.cuh): x * s = h0 + h1 / 2^11 with a per-tensor power-of-two scale s from the
//           tensor's absolute maximum (absmax_kernel, or a word the producing GEMM's epilogue reduced), 22 mantissa
//           bits; A0 B0 -> main accumulator, A0 B1 + A1 B0 -> correction accumulator (carries 2^11), epilogue
//           (main + 2^-11 corr) / (s_a s_b); 3 MMAs per product, FLOPs per launch = 6 * M * N * K.
//   gemm_split2h_2sm_db_kernel: NP = 2 with 256 x 128 pair tiles and two accumulator pairs in TMEM (MMAs overlap the
//           epilogue); correct, measured slower than the 256-wide tiles, selectable with ptrec_tc_set_bn.
#include "tcgen05.cuh"
#include "split3.cuh"
#include "gemm2sm.cuh"

namespace ptrec {

// ------------------------------------------------------------------------------------------------ split into planes
constexpr int kSpTile = 64;
constexpr int kSpThreads = 256;
constexpr int kSpPitch = kSpTile + 8;  // bf16 elements: rows stay 16-byte aligned

struct Split3Args {
  const float* src;
  int64_t ld;
  int R, C;
  const float* ref;  // ReLU output (or any tensor whose sign gates src); null = no mask
  int64_t ld_ref;
  unsigned short* planes;  // [NP][R][pl_ld] bf16 / fp16 bits, columns [C, Cp) written as zero; null = skip
  int64_t pl_ld, pl_plane;
  int Cp;
  unsigned short* planes_t;  // [NP][C][pt_ld], columns [R, Rp) written as zero; null = skip
  int64_t pt_ld, pt_plane;
  int Rp;
  float* colsum_part;  // [gridDim.y][C] per-row-tile column sums of the masked src; null = skip
  // NP == 2 (fp16 x 2) only: per-CTA maxima of |src| written by absmax_kernel, and where the scale goes
  const float* absmax_part;
  int n_part;
  float* scale_out;
  // NP == 2, "prescaled" mode (scale_in != null): the scale is given (a word the caller carries from step to step,
  // ptrec_tc_scale_roll) instead of being derived from this tensor's maximum, so there is no maximum pass; the
  // kernel raises *max_out to max |masked src| (atomicMax on the bit pattern) for the next roll.
  const float* scale_in;
  uint32_t* max_out;
};

// |src| maxima, one per CTA (no atomics: the split kernel reduces the partials itself).  Warp per row, lanes stride
// over the row's 16-byte chunks: coalesced, no index arithmetic beyond adds.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ src, int64_t ld, int R, int C,
                                                     float* __restrict__ part) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const int c4 = vec ? (C >> 2) : 0;
  float m = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < R; r += (int64_t)gridDim.x * 8) {
    const float* p = src + r * ld;
#pragma unroll 4
    for (int c = lane; c < c4; c += 32) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p) + c);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
    }
    for (int c = 4 * c4 + lane; c < C; c += 32) m = fmaxf(m, fabsf(__ldg(p + c)));
  }
  __shared__ float s_m[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_m[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_m[w]);
    part[blockIdx.x] = m;
  }
}

// NP = 3: exact bf16 x 3 split.  NP = 2: fp16 x 2 split of src * scale (split3.cuh), scale from the |src| maxima.
template <int NP>
__global__ void __launch_bounds__(kSpThreads) split_kernel(const Split3Args a) {
  __shared__ __align__(16) unsigned short s_t[NP][kSpTile][kSpPitch];
  __shared__ float s_cs[16][kSpTile];
  __shared__ float s_red[kSpThreads / 32];
  float scale = 1.f;
  float amax = 0.f;  // prescaled mode: max |masked src| seen by this thread
  if (NP == 2 && a.scale_in != nullptr) {
    scale = *a.scale_in;
  } else if (NP == 2) {
    float m = 0.f;
    for (int i = threadIdx.x; i < a.n_part; i += kSpThreads) m = fmaxf(m, a.absmax_part[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = s_red[0];
#pragma unroll
    for (int w = 1; w < kSpThreads / 32; ++w) m = fmaxf(m, s_red[w]);
    scale = h2_scale(m);  // every thread of every CTA derives the same value
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *a.scale_out = scale;
  }
  const int c0 = blockIdx.x * kSpTile, r0 = blockIdx.y * kSpTile;
  const int cg = threadIdx.x & 15, rr = threadIdx.x >> 4;  // 4 columns x (rows rr, rr+16, rr+32, rr+48)
  const int c = c0 + cg * 4;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  const bool vec_in = (a.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.src) & 15) == 0) && (c + 3 < a.C);
  const bool vec_ref = a.ref != nullptr && (a.ld_ref % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.ref) & 15) == 0) &&
                       (c + 3 < a.C);
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int rl = rr + pass * 16, r = r0 + rl;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < a.R) {
      const float* p = a.src + (int64_t)r * a.ld + c;
      if (vec_in) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (c + i < a.C) v[i] = p[i];
      }
      if (a.ref != nullptr) {
        const float* q = a.ref + (int64_t)r * a.ld_ref + c;
        float m[4] = {1.f, 1.f, 1.f, 1.f};
        if (vec_ref) {
          const float4 t = *reinterpret_cast<const float4*>(q);
          m[0] = t.x; m[1] = t.y; m[2] = t.z; m[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (c + i < a.C) m[i] = q[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (!(m[i] > 0.f)) v[i] = 0.f;
      }
    }
    __align__(8) unsigned short pl[NP][4];
    if (NP == 3) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat16 b0, b1, b2;
        split3(v[i], b0, b1, b2);
        pl[0][i] = __bfloat16_as_ushort(b0);
        pl[1][i] = __bfloat16_as_ushort(b1);
        pl[NP - 1][i] = __bfloat16_as_ushort(b2);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; i += 2) {
        uint32_t q0, q1;
        split2h_pair(v[i], v[i + 1], scale, q0, q1);
        *reinterpret_cast<uint32_t*>(&pl[0][i]) = q0;
        *reinterpret_cast<uint32_t*>(&pl[1][i]) = q1;
        amax = fmaxf(amax, fmaxf(fabsf(v[i]), fabsf(v[i + 1])));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) cs[i] += v[i];
    if (a.planes != nullptr && r < a.R && c < a.Cp) {  // Cp % 8 == 0, c % 4 == 0: the group of 4 is inside the pitch
      unsigned short* o = a.planes + (int64_t)r * a.pl_ld + c;
#pragma unroll
      for (int p = 0; p < NP; ++p) *reinterpret_cast<uint2*>(o + p * a.pl_plane) = *reinterpret_cast<const uint2*>(pl[p]);
    }
    if (a.planes_t != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int p = 0; p < NP; ++p) s_t[p][cg * 4 + i][rl] = pl[p][i];
      }
    }
  }
  if (a.colsum_part != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) s_cs[rr][cg * 4 + i] = cs[i];
  }
  if (NP == 2 && a.max_out != nullptr) {  // one combining atomic per warp (non-negative floats order as uints)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(a.max_out, __float_as_uint(amax));
  }
  __syncthreads();
  if (a.colsum_part != nullptr && threadIdx.x < kSpTile && c0 + (int)threadIdx.x < a.C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += s_cs[k][threadIdx.x];  // fixed order
    a.colsum_part[(int64_t)blockIdx.y * a.C + c0 + threadIdx.x] = s;
  }
  if (a.planes_t != nullptr) {
    const int seg = threadIdx.x & 7;  // 8 rows (16 bytes) of the transposed row
    const int r = r0 + seg * 8;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int cl = (threadIdx.x >> 3) + pass * 32;
      const int cc = c0 + cl;
      if (cc < a.C && r < a.Rp) {  // Rp % 8 == 0: the group of 8 is inside the pitch; rows >= R hold zeros
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          *reinterpret_cast<uint4*>(a.planes_t + (int64_t)p * a.pt_plane + (int64_t)cc * a.pt_ld + r) =
              *reinterpret_cast<const uint4*>(&s_t[p][cl][seg * 8]);
        }
      }
    }
  }
}

// Carried scales of the fused tower.  A slot = {last scale, max |tensor| seen since the last roll (fp32 bits, raised by
// atomicMax from the kernels that wrote the tensor's planes)}.  The roll turns every slot's maximum into the scale of
// the NEXT use — the maximum lands in [2^(13-kH2Headroom), 2^(14-kH2Headroom)), so a tensor may grow 2^kH2Headroom-fold
// from one step to the next before its first plane leaves the fp16 range — copies the scales into a per-call array
// (what the kernels of one forward / backward read: a later roll cannot change the scale of planes already written)
// and raises *err if a maximum did leave the range under the scale it was split with.
constexpr int kH2Headroom = 8;
__global__ void scale_roll_kernel(float* __restrict__ slots, int n, float* __restrict__ call_scales,
                                  int32_t* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* s = slots + 2 * i;
  const float last = s[0], m = s[1];
  float next = last;
  if (m > 0.f) {
    if (last > 0.f && !(m * last < 65504.f)) atomicOr(err, 1);  // inf in a plane written since the last roll
    const float full = h2_scale(m);                              // max -> [2^13, 2^14)
    const int se = max(-126, (int)((__float_as_uint(full) >> 23) & 0xFFu) - 127 - kH2Headroom);
    next = __uint_as_float((uint32_t)(se + 127) << 23);
  }
  if (!(next > 0.f)) next = 1.f;  // never used: a slot is seeded with a measured maximum before its first roll
  s[0] = next;
  s[1] = 0.f;
  call_scales[i] = next;
}

// out[c] = sum over row tiles of part[t][c]: 8 lanes per column stride over the tiles, then a fixed xor tree
// (deterministic bias gradient)
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ part, int tiles, int C,
                                                            float* __restrict__ out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sub = threadIdx.x >> 5;  // 8 warps: warp w sums tiles w, w+8, ... (coalesced across the 32 columns)
  __shared__ float s[8][33];
  float acc = 0.f;
  if (c < C)
    for (int t = sub; t < tiles; t += 8) acc += part[(int64_t)t * C + c];
  s[sub][threadIdx.x & 31] = acc;
  __syncthreads();
  if (sub == 0 && c < C) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += s[k][threadIdx.x];
    out[c] = r;
  }
}

// out[e] = sum_s partial[s][e], fixed order (deterministic split-K)
__global__ void partial_reduce_kernel(const float* __restrict__ partial, int splits, int64_t n, float* __restrict__ out) {
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= n) return;
  float4 acc = *reinterpret_cast<const float4*>(partial + e);
  for (int s = 1; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(partial + (int64_t)s * n + e);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(out + e) = acc;
}

// ------------------------------------------------------------------------------------------------ the GEMM
// Operand delivery from L2 is what bounds this kernel (ncu: ~38 B/clk/SM, the LTS cap), so the CTA tile is as wide as
// TMEM allows: 128 x 256 with K blocks of 32 -> 72 KB per stage for 12 MMAs of N = 256 (0.75x the bytes per FLOP of
// 128 x 128 tiles).  The two fp32 accumulators (main + correction, below) take all 512 TMEM columns, so the
// accumulator is single-buffered: the MMA warp waits for the epilogue to pull a tile into registers (~5 % of a tile).
constexpr int kLBM = 128, kLBN = 256, kLBK = 32, kLUmmaK = 16;
constexpr int kLEpiWarps = 8;
constexpr int kLThreads = 64 + 32 * kLEpiWarps;
constexpr uint32_t kLTileA = kLBM * kLBK * 2;          // one A plane tile:  128 rows x 32 bf16 (8 KB)
constexpr uint32_t kLTileB = kLBN * kLBK * 2;          // one B plane tile:  256 rows x 32 bf16 (16 KB)
// NP planes per operand: 3 (bf16 x 3: A0 A1 A2 B0 B1 B2 = 72 KB per stage, 3 stages) or 2 (fp16 x 2: 48 KB, 4 stages)
template <int NP> struct LinCfg {
  static constexpr int kStages = NP == 3 ? 3 : 4;
  static constexpr uint32_t kStageBytes = NP * (kLTileA + kLTileB);
  // instruction descriptor: D = fp32 (bit 4); A / B format bf16 (1 at bits 7 / 10) or fp16 (0)
  static constexpr uint32_t kIdescFmt = NP == 2 ? (1u << 4) : ((1u << 4) | (1u << 7) | (1u << 10));
};
constexpr uint32_t kLTmemCols = 512;                   // main [0, 256) + correction [256, 512)
constexpr size_t kLSmem = (size_t)3 * 72 * 1024 + 1024 /*align*/ + 2048 /*barriers + bias slice*/;
static_assert((size_t)LinCfg<3>::kStages * LinCfg<3>::kStageBytes + 3072 <= kLSmem, "bf16 x 3 stages exceed the budget");
static_assert((size_t)LinCfg<2>::kStages * LinCfg<2>::kStageBytes + 3072 <= kLSmem, "fp16 x 2 stages exceed the budget");

struct LinEpi {
  const float* bias;  // [N] or null
  int relu;
  float* out;         // [splits][M, ldo] fp32
  int64_t ldo;
  int splits;         // K cut into `splits` ranges (>= 1), one output slab each
  __nv_bfloat16* planes;  // optional (bf16 x 3 only): the result again as planes [3][M][pl_ld] (next layer's operand)
  int64_t pl_ld, pl_plane;
  const float* scale_a;   // fp16 x 2 only: the power-of-two scales the operands were split with (device words)
  const float* scale_b;
  uint32_t* absmax_out;   // optional: max |out| as fp32 bits, combined with atomicMax (the caller zeroes the word); lets the
                          // next split of `out` skip its own pass over the tensor
  // CTA-pair kernel, fp16 x 2 only — the fused tower (ptrec_tc_gemm_split2h_fused): the epilogue hands the result to its
  // consumer in the consumer's operand format, so no split / maximum pass runs between two GEMMs
  unsigned short* h2_planes;  // [2][M][h2_ld] fp16 planes of the (masked) result, split with *h2_scale; out may then be null
  int64_t h2_ld, h2_plane;
  const float* h2_scale;      // device word: power-of-two scale carried from the previous step (ptrec_tc_scale_roll)
  const uint32_t* mask_in;    // [M][mask_ld] bit c%32 of word c/32 set = keep element (row, c): ReLU backward of the layer below
  uint32_t* mask_out;         // same layout: bit set where the result (after bias / ReLU) is > 0
  int64_t mask_ld;            // words per row, a multiple of 4 >= ceil(N / 32)
  float* colsum_part;         // [ceil(M / 32)][N] column sums of the masked result per 32-row block (bias gradient)
  // CTA-pair kernel with ONE operand plane (NP = 1: a plain bf16 GEMM) — the DCN-v2 cross layers (K5, dcn_cross.cu):
  //   dcn_mode 1 (forward):        u = acc + bias;  o0 = dp0 * u + dp1 (dp0 = x0, dp1 = x_l);  o1 = u
  //   dcn_mode 2 (input gradient): g_x = acc + dp0 (dp0 = g_out);  o0 = g_x;  o1 = g_x * dp1 (dp1 = x0)
  // dp0 / dp1 are bf16 [M][dld]; o0 / o1 leave through the tensor maps p / q (bf16 [M][dld], o1 optional)
  int dcn_mode;
  const __nv_bfloat16* dp0;
  const __nv_bfloat16* dp1;
  int64_t dld;
  int dcn_o1;
};

// acc = main + correction.  bf16 x 3: plain sum.  fp16 x 2: the correction accumulator holds 2^11 (A0 B1 + A1 B0) of
// the SCALED operands; inv_a / inv_b (exact powers of two) undo the operand scales.
template <int NP>
__device__ __forceinline__ float lin_combine(float main_acc, float corr_acc, float inv_a, float inv_b) {
  if (NP == 3) return main_acc + corr_acc;
  return (fmaf(corr_acc, kH2SecondInv, main_acc) * inv_a) * inv_b;
}

struct LinMaps {
  CUtensorMap a, b;  // 3-D: K-major (k, row, plane) or MN-major (mn, k row, plane)
  // CTA-pair kernel: where the epilogue's TMA stores go.  c: fp32 result (col, row, split-K slab), boxes of 32 x 32,
  // 128-byte swizzle; p: fp16 planes of the result (col, row, plane), boxes of 32 x 32 x 1, 64-byte swizzle
  CUtensorMap c, p;
  CUtensorMap q;     // NP = 1 (DCN epilogue): second bf16 output
  CUtensorMap r0, r1;  // NP = 1 (DCN epilogue): the two bf16 epilogue operands dp0 / dp1, boxes of 32 x 32, 64-byte swizzle
};

// K-major, 64B swizzle (rows of 32 bf16): 8-row atoms of 512 B; LBO unused (1), SBO = 512 B
__device__ __forceinline__ uint64_t make_sw64_desc(const void* smem_tile) {
  const uint32_t addr = smem_u32(smem_tile);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// MN-major, 128B swizzle, boxes of 64 mn-elements x 32 k-rows (4 KB): SBO = 1024 B between groups of 8 k-rows,
// LBO = 4096 B between blocks of 64 mn-elements
__device__ __forceinline__ uint64_t make_sw128_mn32_desc(const void* smem_tile) {
  const uint32_t addr = smem_u32(smem_tile);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(4096 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// MN_MAJOR = false: operands are K-major planes [3][M or N][K]            (y = x W^T, dx = g W^T^T)
// MN_MAJOR = true : operands are MN-major planes [3][K][M or N] — the reduction dimension is the ROW of the stored
//                   matrices, which is how activations and gradients already lie for dW = g^T x (K = batch): the
//                   tensor core reads them transposed straight from shared memory, no transpose pass.
template <bool MN_MAJOR, int NP>
__global__ void __launch_bounds__(kLThreads, 1)
gemm_split3_kernel(const __grid_constant__ LinMaps maps, int M, int N, int K, LinEpi ep) {
  constexpr int kLStages = LinCfg<NP>::kStages;
  constexpr uint32_t kLStageBytes = LinCfg<NP>::kStageBytes;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kLStages * kLStageBytes);
  uint64_t* full = bars;                    // [kLStages]
  uint64_t* empty = bars + kLStages;        // [kLStages]
  uint64_t* acc_full = bars + 2 * kLStages; // [1]
  uint64_t* acc_empty = acc_full + 1;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  float* s_bias = reinterpret_cast<float*>(bars + 16);  // [kLBN], 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + kLBN - 1) / kLBN, tiles_m = (M + kLBM - 1) / kLBM;
  const int splits = max(ep.splits, 1);
  const int mn_tiles = tiles_n * tiles_m;
  const int n_tiles = mn_tiles * splits;
  const int total_kb = (K + kLBK - 1) / kLBK;
  const int kb_per_split = (total_kb + splits - 1) / splits;
  // tile order: n slowest, so that the widest column tiles come first and the narrow remainder tiles fill the tail
  // of the round-robin (longest-processing-time-first for the two tile widths a Linear layer produces)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kLStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kLEpiWarps);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kLTmemCols)
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
      int kbg = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mn = tile % mn_tiles, split = tile / mn_tiles;
        const int m0 = (mn % tiles_m) * kLBM, n0 = (mn / tiles_m) * kLBN;
        const int kb0 = split * kb_per_split, num_kb = max(0, min(total_kb - kb0, kb_per_split));
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % kLStages;
          mbar_wait(&empty[s], ((kbg / kLStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], kLStageBytes);
          unsigned char* sa = smem + (size_t)s * kLStageBytes;
          unsigned char* sb = sa + NP * kLTileA;
          const int k0 = (kb0 + kb) * kLBK;
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            if (MN_MAJOR) {  // boxes of 64 (mn) x 32 (k rows), 4 KB each: 2 per A tile, 4 per B tile
#pragma unroll
              for (int j = 0; j < kLBM / 64; ++j)
                tma_load_3d(sa + (size_t)p * kLTileA + (size_t)j * 4096, &maps.a, m0 + 64 * j, k0, p, &full[s]);
#pragma unroll
              for (int j = 0; j < kLBN / 64; ++j)
                tma_load_3d(sb + (size_t)p * kLTileB + (size_t)j * 4096, &maps.b, n0 + 64 * j, k0, p, &full[s]);
            } else {
              tma_load_3d(sa + (size_t)p * kLTileA, &maps.a, k0, m0, p, &full[s]);
              tma_load_3d(sb + (size_t)p * kLTileB, &maps.b, k0, n0, p, &full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int kbg = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(acc_empty, (it & 1) ^ 1);  // the epilogue holds the previous tile in registers
        tcgen05_fence_after();
        // Two accumulators per tile.  The tensor core truncates (does not round) when it aligns an MMA result with
        // the running sum, a bias that grows with the number of accumulations; keeping the five small pairs in their
        // own accumulator makes their truncation relative to a 2^-8 smaller magnitude and leaves the main one with
        // the K/16 accumulations of a plain bf16 GEMM.  The epilogue adds the two in fp32.
        const uint32_t tmem_d = tmem_base;
        const uint32_t tmem_c = tmem_base + (uint32_t)kLBN;
        const int mn = tile % mn_tiles;
        const int n0 = (mn / tiles_m) * kLBN;
        const int n_eff = min(kLBN, (N - n0 + 15) & ~15);  // the last column tile issues narrower MMAs
        const uint32_t idesc = LinCfg<NP>::kIdescFmt | ((uint32_t)(n_eff >> 3) << 17) |
                               ((uint32_t)(kLBM >> 4) << 24) | (MN_MAJOR ? ((1u << 15) | (1u << 16)) : 0u);
        const int kb0 = (tile / mn_tiles) * kb_per_split, num_kb = max(0, min(total_kb - kb0, kb_per_split));
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % kLStages;
          mbar_wait(&full[s], (kbg / kLStages) & 1);
          tcgen05_fence_after();
          unsigned char* sa = smem + (size_t)s * kLStageBytes;
          unsigned char* sb = sa + NP * kLTileA;
          uint64_t ad[NP], bd[NP];
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            ad[p] = MN_MAJOR ? make_sw128_mn32_desc(sa + (size_t)p * kLTileA) : make_sw64_desc(sa + (size_t)p * kLTileA);
            bd[p] = MN_MAJOR ? make_sw128_mn32_desc(sb + (size_t)p * kLTileB) : make_sw64_desc(sb + (size_t)p * kLTileB);
          }
#pragma unroll
          for (int k = 0; k < kLBK / kLUmmaK; ++k) {
            // K-major: +32 bytes per K=16 slice inside the swizzled row; MN-major: 16 k-rows of 128 bytes = +2048 bytes
            const uint64_t o = MN_MAJOR ? (uint64_t)(k * 128) : (uint64_t)(k * 2);
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            if (NP == 3) {
              umma_bf16(tmem_c, ad[0] + o, bd[NP - 1] + o, idesc, acc);  // smallest pairs first
              umma_bf16(tmem_c, ad[NP - 1] + o, bd[0] + o, idesc, 1u);
              umma_bf16(tmem_c, ad[1] + o, bd[1] + o, idesc, 1u);
              umma_bf16(tmem_c, ad[0] + o, bd[1] + o, idesc, 1u);
              umma_bf16(tmem_c, ad[1] + o, bd[0] + o, idesc, 1u);
            } else {  // fp16 x 2: the two cross terms (second planes carry 2^11) go to the correction accumulator
              umma_bf16(tmem_c, ad[0] + o, bd[1] + o, idesc, acc);
              umma_bf16(tmem_c, ad[1] + o, bd[0] + o, idesc, 1u);
            }
            umma_bf16(tmem_d, ad[0] + o, bd[0] + o, idesc, acc);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(acc_full);
      }
    }
  } else {
    const int e = warp - 2;
    const int q = warp & 3;   // TMEM lane quadrant this warp may access
    const int half = e >> 2;  // 128-column half of the tile
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const float inv_a = (NP == 2) ? 1.f / ep.scale_a[0] : 1.f;  // exact: the scales are powers of two
    const float inv_b = (NP == 2) ? 1.f / ep.scale_b[0] : 1.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int mn = tile % mn_tiles, split = tile / mn_tiles;
      const int m0 = (mn % tiles_m) * kLBM, n0 = (mn / tiles_m) * kLBN;
      if (ep.bias != nullptr) {
        asm volatile("bar.sync 2, %0;" ::"n"(32 * kLEpiWarps) : "memory");  // previous tile's readers are done
        if (et < kLBN) s_bias[et] = (n0 + et < N) ? ep.bias[n0 + et] : 0.f;
        asm volatile("bar.sync 2, %0;" ::"n"(32 * kLEpiWarps) : "memory");
      }
      mbar_wait(acc_full, it & 1);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128);
      const int row = m0 + r, col0 = n0 + half * 128;
      const bool live = col0 < N;  // warp-uniform: a narrow remainder tile leaves the second half empty
      uint32_t v[128];
      if (live) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          uint32_t w[32];
          tmem_ld32(tacc + h * 32, v + h * 32);
          tmem_ld32(tacc + kLBN + h * 32, w);  // + correction accumulator
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[h * 32 + j] = __float_as_uint(lin_combine<NP>(__uint_as_float(v[h * 32 + j]), __uint_as_float(w[j]),
                                                            inv_a, inv_b));
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      float amax = 0.f;
      if (live && row < M) {
        float* o = ep.out + ((int64_t)split * M + row) * ep.ldo + col0;
        const bool has_bias = ep.bias != nullptr;
#pragma unroll
        for (int j = 0; j < 128; j += 4) {
          if (col0 + j < N) {  // ldo % 4 == 0 and ldo >= round_up(N, 4): the whole group is inside the pitch
            float4 t = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]));
            if (has_bias) {
              const float4 b = *reinterpret_cast<const float4*>(&s_bias[half * 128 + j]);
              t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
            }
            if (ep.relu) {
              t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f);
            }
            *reinterpret_cast<float4*>(o + j) = t;
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
            if (NP == 3 && ep.planes != nullptr) {  // pl_ld >= N rounded up to 8: the group of 4 is inside the pitch
              const float tv[4] = {t.x, t.y, t.z, t.w};
              split3_store4(tv, ep.planes + (int64_t)row * ep.pl_ld + col0 + j, ep.pl_plane);
            }
          }
        }
      }
      if (ep.absmax_out != nullptr) {  // warp-uniform branch; one combining atomic per warp and tile
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if (lane == 0 && amax > 0.f) atomicMax(ep.absmax_out, __float_as_uint(amax));  // non-negative floats order as uints
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kLTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ 2-SM variant
// The same GEMM with CTA pairs (cta_group::2): a pair owns a 256 x 256 tile — each CTA holds 128 rows of A, its own
// 128 accumulator rows in TMEM, and HALF of the B tile; the pair's tensor cores share the B halves.  Per CTA and K
// block that is 48 KB from L2 (72 KB in the 1-SM kernel) and two thirds of the shared-memory operand reads, the two
// limits the 1-SM kernel runs into.  Both CTAs run a TMA producer (completing on the leader's barrier), only the
// leader issues MMAs; tcgen05.commit multicasts the stage release / accumulator-ready arrivals to both CTAs, and both
// CTAs' epilogue warps arrive remotely on the leader's accumulator-free barrier.
#ifndef K2_STAGES
#define K2_STAGES 4
#endif
// 16 epilogue warps (4 per scheduler): the epilogue is a chain of dependent fixed-latency instructions per element, and
// with 2 warps per scheduler it ran at ~0.1 instructions per cycle and warp (8.2k cycles per tile, tools/diag_k6_timeline.py)
constexpr int k2EpiWarps = 16;
constexpr int k2Threads = 64 + 32 * k2EpiWarps;
constexpr uint32_t k2Staging = k2EpiWarps * 4096;  // epilogue: one 32 x 32 fp32 (or 2 x 32 x 32 fp16) box per warp
constexpr size_t k2Smem = 232448;                  // everything a CTA may use: stages + staging + barriers
constexpr size_t k2StageBudget = k2Smem - k2Staging - 1024 /*align*/ - 2048 /*barriers + bias slice*/;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;         // clears the CTA-pair bit of a shared::cluster address -> leader

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // remote arrive on the leader CTA's barrier
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

#ifdef PTREC_K6_TIMELINE
// Diagnostic build (make EXTRA=-DPTREC_K6_TIMELINE; tools/diag_k6_timeline.py): the first epilogue thread of every CTA
// stamps %globaltimer at the kernel's milestones into g_k6_timeline[cta][32].
__device__ unsigned long long* g_k6_timeline = nullptr;
__device__ __forceinline__ void k6_stamp(int slot) {
  if (g_k6_timeline != nullptr && threadIdx.x == 64 && slot < 32) {  // SM cycle counter: ~0.52 ns at 1.9 GHz
    g_k6_timeline[(size_t)blockIdx.x * 32 + slot] = (unsigned long long)clock64();
  }
}
#define K6_STAMP(slot) k6_stamp(slot)
#else
#define K6_STAMP(slot)
#endif

// BK = K elements per stage: 32 (64-byte swizzle, 4 stages of 48 KB) or 64 (128-byte swizzle, 2 stages of 96 KB)
template <bool MN_MAJOR, int BK, int NP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
gemm_split3_2sm_kernel(const __grid_constant__ LinMaps maps, int M, int N, int K, LinEpi ep) {
  // a stage = NP planes of A (128 rows) and of this CTA's B half: fp16 x 2, BK 32: 32 KB -> 5 stages; bf16 x 3: 48 KB -> 3
  constexpr uint32_t k2Tile = 128 * BK * 2;     // any plane tile: 128 rows (or 2 boxes of 64 mn) x BK 16-bit elements
  constexpr uint32_t k2StageBytes = 2 * NP * k2Tile; // A planes, then B planes (B = this CTA's half)
  // NP = 1 (K5): one plane leaves TMEM columns [256, 512) free, so the accumulator is double-buffered (the MMAs of tile
  // i + 1 run under the epilogue of tile i).  Every epilogue warp's 4 KB staging holds the two 32 x 32 bf16 boxes of its
  // epilogue operands for ONE 32-column chunk, fetched by TMA (chunk 0 while the tile's MMAs run, chunk 1 once chunk 0's
  // stores have read the boxes) and overwritten in place by the two results, which leave by TMA stores: no per-thread
  // global access in the epilogue.  (Boxes for both chunks — 8 KB per warp — left 6 operand stages = 96 KB in flight per
  // SM, and the mainloop ran at the rate that many bytes cover the L2 / HBM latency: 18k cycles per tile for 6k cycles
  // of MMAs; the exposed fetch of chunk 1 hides under the next tile's mainloop.)
  constexpr int kAcc = NP == 1 ? 2 : 1;
  constexpr uint32_t kWarpStg = 4096u;
  constexpr uint32_t kStaging = k2EpiWarps * kWarpStg;
  static_assert(kStaging >= k2Staging, "staging");
  constexpr int k2Stages = (int)((k2StageBudget + k2Staging - kStaging) / k2StageBytes);
  static_assert(k2Stages >= 1 && 2 * k2Stages + 2 * kAcc + 1 <= 32, "a stage exceeds the shared-memory budget / barrier block");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* staging = smem + (size_t)k2Stages * k2StageBytes;  // 1024-byte aligned: the stages are multiples of 1 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStaging);
  uint64_t* full = bars;                    // [k2Stages]  (the leader's are the ones in use)
  uint64_t* empty = bars + k2Stages;        // [k2Stages]
  uint64_t* acc_full = bars + 2 * k2Stages; // [kAcc]
  uint64_t* acc_empty = acc_full + kAcc;    // [kAcc]  (leader's)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAcc);
  float* s_bias = reinterpret_cast<float*>(bars + 32);  // [256], 256 bytes into the barrier block (up to 15 stages of barriers before it)
  uint64_t* in_bar = bars + 32 + 128;  // [k2EpiWarps] (NP = 1): a warp's epilogue operands have landed; ends 1408 B into the block

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int tiles_n = (N + 255) / 256, tiles_m = (M + 255) / 256;
  const int splits = max(ep.splits, 1);
  const int mn_tiles = tiles_n * tiles_m;
  const int n_tiles = mn_tiles * splits;
  const int total_kb = (K + BK - 1) / BK;
  const int kb_per_split = (total_kb + splits - 1) / splits;
  const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  // NP = 1 (K5): column tiles of equal width (a multiple of 32: the epilogue's chunk) instead of 256, 256, ... remainder,
  // and the column tile is the FAST index of the tile order: the tiles that run at the same time share row blocks of A,
  // which is then read from HBM once (with rows fastest, the d = 848 cross layer re-read its 55 MB input from HBM for
  // every one of its 4 column tiles: the kernel ran at the HBM rate, 20k cycles per tile for 7k cycles of MMAs)
  int bn = 256;
  if (NP == 1) {
    bn = (((N + tiles_n - 1) / tiles_n) + 31) & ~31;
    if ((tiles_n - 1) * bn >= N) bn = 256;
  }
#define K2_TILE_MN(mn, m_idx, n_idx)                       \
  int m_idx, n_idx;                                        \
  if (NP == 1) { n_idx = (mn) % tiles_n; m_idx = (mn) / tiles_n; } \
  else { m_idx = (mn) % tiles_m; n_idx = (mn) / tiles_m; }
  K6_STAMP(0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < k2Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < kAcc; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 2 * k2EpiWarps);  // the epilogue warps of both CTAs
    }
    if (NP == 1)
      for (int w = 0; w < k2EpiWarps; ++w) mbar_init(&in_bar[w], 1);
    fence_mbar_init();
  }
  cluster_sync_all();  // the peer's barriers exist before anything arrives on them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kLTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b) : "memory");
      int kbg = 0;
      for (int tile = cid; tile < n_tiles; tile += n_clusters) {
        const int mn = tile % mn_tiles, split = tile / mn_tiles;
        K2_TILE_MN(mn, m_idx, n_idx)
        const int m0 = m_idx * 256 + rank * 128, n0 = n_idx * bn;
        const int n_eff = min(bn, (N - n0 + 15) & ~15);
        const int nb0 = n0 + rank * (n_eff >> 1);  // this CTA's half of the B tile
        const int kb0 = split * kb_per_split, num_kb = max(0, min(total_kb - kb0, kb_per_split));
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % k2Stages;
          mbar_wait(&empty[s], ((kbg / k2Stages) & 1) ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * k2StageBytes);  // both CTAs' bytes land on the leader
          unsigned char* sa = smem + (size_t)s * k2StageBytes;
          unsigned char* sb = sa + NP * k2Tile;
          const int k0 = (kb0 + kb) * BK;
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            if (MN_MAJOR) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                tma_load_3d_2sm(sa + (size_t)p * k2Tile + (size_t)j * (k2Tile / 2), &maps.a, m0 + 64 * j, k0, p, &full[s]);
                tma_load_3d_2sm(sb + (size_t)p * k2Tile + (size_t)j * (k2Tile / 2), &maps.b, nb0 + 64 * j, k0, p, &full[s]);
              }
            } else {
              tma_load_3d_2sm(sa + (size_t)p * k2Tile, &maps.a, k0, m0, p, &full[s]);
              tma_load_3d_2sm(sb + (size_t)p * k2Tile, &maps.b, k0, nb0, p, &full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      int kbg = 0, it = 0;
      for (int tile = cid; tile < n_tiles; tile += n_clusters, ++it) {
        const int ab = it % kAcc;
        mbar_wait(&acc_empty[ab], ((it / kAcc) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)ab * 256u;
        const uint32_t tmem_c = tmem_base + 256u;  // NP > 1 only
        const int mn = tile % mn_tiles;
        K2_TILE_MN(mn, m_idx, n_idx)
        (void)m_idx;
        const int n0 = n_idx * bn;
        const int n_eff = min(bn, (N - n0 + 15) & ~15);
        const uint32_t idesc = LinCfg<NP>::kIdescFmt | ((uint32_t)(n_eff >> 3) << 17) |
                               ((uint32_t)(256 >> 4) << 24) | (MN_MAJOR ? ((1u << 15) | (1u << 16)) : 0u);
        const int kb0 = (tile / mn_tiles) * kb_per_split, num_kb = max(0, min(total_kb - kb0, kb_per_split));
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % k2Stages;
          mbar_wait(&full[s], (kbg / k2Stages) & 1);
          tcgen05_fence_after();
          unsigned char* sa = smem + (size_t)s * k2StageBytes;
          unsigned char* sb = sa + NP * k2Tile;
          uint64_t ad[NP], bd[NP];
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const void* pa = sa + (size_t)p * k2Tile;
            const void* pb = sb + (size_t)p * k2Tile;
            if (BK == 64) {
              ad[p] = MN_MAJOR ? make_sw128_mn_desc(pa) : make_sw128_desc(pa);
              bd[p] = MN_MAJOR ? make_sw128_mn_desc(pb) : make_sw128_desc(pb);
            } else {
              ad[p] = MN_MAJOR ? make_sw128_mn32_desc(pa) : make_sw64_desc(pa);
              bd[p] = MN_MAJOR ? make_sw128_mn32_desc(pb) : make_sw64_desc(pb);
            }
          }
#pragma unroll
          for (int k = 0; k < BK / kLUmmaK; ++k) {
            const uint64_t o = MN_MAJOR ? (uint64_t)(k * 128) : (uint64_t)(k * 2);
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            if constexpr (NP == 3) {
              umma_bf16_2sm(tmem_c, ad[0] + o, bd[NP - 1] + o, idesc, acc);
              umma_bf16_2sm(tmem_c, ad[NP - 1] + o, bd[0] + o, idesc, 1u);
              umma_bf16_2sm(tmem_c, ad[1] + o, bd[1] + o, idesc, 1u);
              umma_bf16_2sm(tmem_c, ad[0] + o, bd[1] + o, idesc, 1u);
              umma_bf16_2sm(tmem_c, ad[1] + o, bd[0] + o, idesc, 1u);
            } else if constexpr (NP == 2) {
              umma_bf16_2sm(tmem_c, ad[0] + o, bd[1] + o, idesc, acc);
              umma_bf16_2sm(tmem_c, ad[1] + o, bd[0] + o, idesc, 1u);
            }  // NP == 1: a plain bf16 product, main accumulator only
            umma_bf16_2sm(tmem_d, ad[0] + o, bd[0] + o, idesc, acc);
          }
          umma_commit_2sm(&empty[s]);
        }
        umma_commit_2sm(&acc_full[ab]);
      }
    }
  } else {
    // Epilogue, thread = accumulator row (TMEM lane), a warp = 32 rows (its TMEM lane quadrant) x 64 columns, walked in
    // two chunks of 32: bias, ReLU, the ReLU-backward bit mask of the layer below (one word per row and chunk), the > 0
    // bits of the result, |max|, column sums (bias gradient) by a fixed butterfly over the warp's rows.  The result
    // leaves as fp32 and / or as the consumer's two fp16 planes through a swizzled 4 KB staging box per warp and one
    // TMA store per box (whole 128-byte lines, clipped at M / N by the tensor map): per-thread-row stores (32 lines per
    // instruction, 16 or 32 bytes each) back up the LSU and measured 1.5x - 2.5x this epilogue's time
    // (tools/diag_k6_timeline.py).  TMEM goes back to the tensor core as soon as the warp's second chunk is in registers.
    const int e = warp - 2;
    const int q = warp & 3;   // TMEM lane quadrant this warp may access
    const int grp = e >> 2;   // 64-column group of the tile
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const float inv_a = (NP == 2) ? 1.f / ep.scale_a[0] : 1.f;
    const float inv_b = (NP == 2) ? 1.f / ep.scale_b[0] : 1.f;
    const bool want_h2 = NP == 2 && ep.h2_planes != nullptr;
    const float oscale = want_h2 ? ep.h2_scale[0] : 1.f;
    unsigned char* stg = staging + e * kWarpStg;  // this warp's staging box(es)
    const bool dcn = NP == 1 && ep.dcn_mode != 0;
    const bool masked = NP == 2 && (ep.mask_in != nullptr || ep.mask_out != nullptr || ep.colsum_part != nullptr ||
                                    ep.absmax_out != nullptr || want_h2);
    const bool has_bias = ep.bias != nullptr;
    float amax = 0.f;
    int it = 0, nin = 0;  // nin: operand fetches of this warp so far (the parity of its barrier)
    for (int tile = cid; tile < n_tiles; tile += n_clusters, ++it) {
      const int mn = tile % mn_tiles, split = tile / mn_tiles;
      K2_TILE_MN(mn, m_idx, n_idx)
      const int m0 = m_idx * 256 + rank * 128, n0 = n_idx * bn;
      const int n_end = min(N, n0 + bn);  // columns of this tile
      const int ab = it % kAcc;
      if (has_bias) {  // the two barriers keep every epilogue warp inside the same tile: one bias slice is enough
        asm volatile("bar.sync 2, %0;" ::"n"(32 * k2EpiWarps) : "memory");
        if (et < 256) s_bias[et] = (n0 + et < N) ? ep.bias[n0 + et] : 0.f;
        asm volatile("bar.sync 2, %0;" ::"n"(32 * k2EpiWarps) : "memory");
      }
      const int row = m0 + r, col0 = n0 + grp * 64;
      const bool live = col0 < n_end;
      // a chunk's epilogue operands: boxes of 32 rows x 32 columns (2 KB), dp0 at stg, dp1 at stg + 2048, requested once
      // the previous stores have read the boxes (loads of clipped boxes still deliver every byte)
      auto fetch_operands = [&](int cl) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive_expect_tx(&in_bar[e], 4096u);
          tma_load_3d(stg, &maps.r0, col0 + cl * 32, m0 + q * 32, 0, &in_bar[e]);
          tma_load_3d(stg + 2048, &maps.r1, col0 + cl * 32, m0 + q * 32, 0, &in_bar[e]);
        }
        __syncwarp();
      };
      if (dcn && live) fetch_operands(0);  // before the accumulator is waited for
      uint2 mw = make_uint2(0xffffffffu, 0xffffffffu);
      if (NP == 2 && ep.mask_in != nullptr && live && row < M)  // requested before the accumulator is waited for
        mw = *reinterpret_cast<const uint2*>(ep.mask_in + (int64_t)row * ep.mask_ld + (col0 >> 5));
      K6_STAMP(1 + 4 * it);
      mbar_wait(&acc_full[ab], (it / kAcc) & 1);
      tcgen05_fence_after();
      K6_STAMP(2 + 4 * it);
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 256 + grp * 64);
      uint32_t mo[2] = {0u, 0u};
#pragma unroll
      for (int cl = 0; cl < 2; ++cl) {
        const int cc0 = col0 + cl * 32;
        const bool on = cc0 < n_end;  // warp-uniform; a narrow tile leaves chunks (or the whole warp) empty
        float x[32];
        if (on) {
          if constexpr (NP == 1) {
            tmem_ld32_async(tacc + cl * 32, reinterpret_cast<uint32_t*>(x));
            tmem_wait_ld();
          } else {
            uint32_t w[32];
            tmem_ld32_async(tacc + cl * 32, reinterpret_cast<uint32_t*>(x));
            tmem_ld32_async(tacc + 256 + cl * 32, w);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = lin_combine<NP>(x[j], __uint_as_float(w[j]), inv_a, inv_b);
          }
        }
        if (cl == 1) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&acc_empty[ab]);
          K6_STAMP(3 + 4 * it);
        }
        if (on) {
          if (has_bias) {  // 8 independent 16-byte broadcast loads
            float4 bq[8];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) bq[j4] = *reinterpret_cast<const float4*>(&s_bias[grp * 64 + cl * 32 + 4 * j4]);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              x[4 * j4] += bq[j4].x; x[4 * j4 + 1] += bq[j4].y; x[4 * j4 + 2] += bq[j4].z; x[4 * j4 + 3] += bq[j4].w;
            }
          }
          if (ep.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          if (masked) {  // warp-uniform: the plain fp32 product (wgrad partials, per-layer path) skips all of this
            const int nvalid = min(32, N - cc0);
            uint32_t keep = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);  // TMEM columns >= N hold nothing
            keep = row < M ? (keep & (cl == 0 ? mw.x : mw.y)) : 0u;
            uint32_t bits = 0u;
            if (__all_sync(0xffffffffu, keep == 0xffffffffu)) {  // interior chunk, nothing masked: no selects
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                bits |= (x[j] > 0.f ? 1u : 0u) << j;
                amax = fmaxf(amax, fabsf(x[j]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float t = ((keep >> j) & 1u) ? x[j] : 0.f;
                bits |= (t > 0.f ? 1u : 0u) << j;
                amax = fmaxf(amax, fabsf(t));
                x[j] = t;
              }
            }
            mo[cl] = bits;
          }
          if (dcn) {
            // DCN-v2 cross layer epilogue: this thread's row of the two bf16 operand boxes (32 columns = 64 bytes each,
            // SWIZZLE_64B: 16-byte chunk g of row l at g ^ ((l >> 1) & 3)) is read 8 columns at a time and overwritten
            // by the two bf16 results; rows / columns outside the matrices were zero-filled by the loads and are
            // clipped by the stores
            unsigned char* box = stg;
            mbar_wait(&in_bar[e], (nin++) & 1);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int o = lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4);
              const uint4 qa = *reinterpret_cast<const uint4*>(box + o);
              const uint4 qb = *reinterpret_cast<const uint4*>(box + 2048 + o);
              const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&qa);
              const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&qb);
              uint32_t w0[4], w1[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 a = __bfloat1622float2(ha[i]), b = __bfloat1622float2(hb[i]);
                const float v0 = x[8 * g + 2 * i], v1 = x[8 * g + 2 * i + 1];
                float2 o0, o1;
                if (ep.dcn_mode == 1) {   // u = acc + bias (added above); x0 * u + x_l; u
                  o0 = make_float2(a.x * v0 + b.x, a.y * v1 + b.y);
                  o1 = make_float2(v0, v1);
                } else {                  // g_x = acc + g_out; g_x; g_x * x0
                  o0 = make_float2(v0 + a.x, v1 + a.y);
                  o1 = make_float2(o0.x * b.x, o0.y * b.y);
                }
                const __nv_bfloat162 p0 = __floats2bfloat162_rn(o0.x, o0.y), p1 = __floats2bfloat162_rn(o1.x, o1.y);
                w0[i] = *reinterpret_cast<const uint32_t*>(&p0);
                w1[i] = *reinterpret_cast<const uint32_t*>(&p1);
              }
              *reinterpret_cast<uint4*>(box + o) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
              *reinterpret_cast<uint4*>(box + 2048 + o) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&maps.p, box, cc0, m0 + q * 32, 0);
              if (ep.dcn_o1) tma_store_3d(&maps.q, box + 2048, cc0, m0 + q * 32, 0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (cl == 0 && cc0 + 32 < n_end) fetch_operands(1);
          } else if (ep.out != nullptr) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the box's last store has read it
            __syncwarp();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)  // row = lane, 128 B per row, 16-byte chunk j4 at j4 ^ (row & 7): SWIZZLE_128B
              *reinterpret_cast<float4*>(stg + lane * 128 + ((j4 ^ (lane & 7)) << 4)) =
                  make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&maps.c, stg, cc0, m0 + q * 32, split);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
          if (want_h2) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 32; g += 16) {  // 16 columns at a time (register pressure), two fp16 per word
              uint32_t p0[8], p1[8];
#pragma unroll
              for (int j = 0; j < 16; j += 2) split2h_pair(x[g + j], x[g + j + 1], oscale, p0[j >> 1], p1[j >> 1]);
#pragma unroll
              for (int h = 0; h < 2; ++h) {  // 64 B per row and plane, 16-byte chunk j4 at j4 ^ ((row >> 1) & 3): SWIZZLE_64B
                const int j4 = (g >> 3) + h;
                const int o = lane * 64 + ((j4 ^ ((lane >> 1) & 3)) << 4);
                *reinterpret_cast<uint4*>(stg + o) = make_uint4(p0[4 * h], p0[4 * h + 1], p0[4 * h + 2], p0[4 * h + 3]);
                *reinterpret_cast<uint4*>(stg + 2048 + o) = make_uint4(p1[4 * h], p1[4 * h + 1], p1[4 * h + 2], p1[4 * h + 3]);
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&maps.p, stg, cc0, m0 + q * 32, 0);
              tma_store_3d(&maps.p, stg + 2048, cc0, m0 + q * 32, 1);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
          if (NP == 3 && ep.planes != nullptr && row < M) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (cc0 + j < N) split3_store4(x + j, ep.planes + (int64_t)row * ep.pl_ld + cc0 + j, ep.pl_plane);
          }
          if (NP == 2 && ep.colsum_part != nullptr) {
            // column sums over the warp's 32 rows: a butterfly that halves the columns a lane carries at every step
            // (31 shuffles for 32 columns); lane l ends with column l.  Fixed order: deterministic.
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const bool up = (lane & o) != 0;
#pragma unroll
              for (int k = 0; k < o; ++k) {
                const float send = up ? x[k] : x[k + o];
                const float recv = __shfl_xor_sync(0xffffffffu, send, o);
                x[k] = (up ? x[k + o] : x[k]) + recv;
              }
            }
            if (cc0 + lane < N) ep.colsum_part[(int64_t)((m0 + q * 32) >> 5) * N + cc0 + lane] = x[0];
          }
        }
      }
      if (NP == 2 && ep.mask_out != nullptr && live && row < M)
        *reinterpret_cast<uint2*>(ep.mask_out + (int64_t)row * ep.mask_ld + (col0 >> 5)) = make_uint2(mo[0], mo[1]);
      K6_STAMP(4 + 4 * it);
    }
    K6_STAMP(30);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // this warp's TMA stores have landed
    K6_STAMP(31);
    if (ep.absmax_out != nullptr) {  // one combining atomic per warp (non-negative floats order as uints)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (lane == 0 && amax > 0.f) atomicMax(ep.absmax_out, __float_as_uint(amax));
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();  // both CTAs are done with TMEM and with each other's barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kLTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------- fp16 x 2, double-buffered accumulators
// With two operand planes the accumulator pair (main + correction) of a 128-column tile takes 256 TMEM columns, so TWO
// tiles fit: the MMAs of tile i+1 run while the epilogue warps drain tile i (the 256-column kernels above hold one tile
// and stall the tensor core for every epilogue — ~1/3 of the time at the DNN-tower shapes, K ~ 400).  CTA pairs own
// 256 x 128 tiles: per CTA and K block 128 rows of A (2 planes x 8 KB) and its 64-row half of B (2 x 4 KB) = 24 KB,
// 8 stages.  Shared-memory operand reads: 6 KB per MMA of 64 tensor-core cycles = 96 B/clk of the 128 available.
constexpr int kDbBN = 128, kDbBK = 32, kDbStages = 8;
constexpr uint32_t kDbTileA = 128 * kDbBK * 2;                  // 8 KB
constexpr uint32_t kDbTileB = 64 * kDbBK * 2;                   // 4 KB
constexpr uint32_t kDbStageBytes = 2 * (kDbTileA + kDbTileB);   // A0 A1 B0 B1 = 24 KB
static_assert((size_t)kDbStages * kDbStageBytes + 3072 <= k2Smem, "stages exceed the shared-memory budget");

template <bool MN_MAJOR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kLThreads, 1)
gemm_split2h_2sm_db_kernel(const __grid_constant__ LinMaps maps, int M, int N, int K, LinEpi ep) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kDbStages * kDbStageBytes);
  uint64_t* full = bars;                       // [kDbStages]  (the leader's are the ones in use)
  uint64_t* empty = bars + kDbStages;          // [kDbStages]
  uint64_t* acc_full = bars + 2 * kDbStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]  (leader's)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_bias = reinterpret_cast<float*>(bars + 32);  // [kDbBN], 256 bytes into the barrier block

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int tiles_n = (N + kDbBN - 1) / kDbBN, tiles_m = (M + 255) / 256;
  const int splits = max(ep.splits, 1);
  const int mn_tiles = tiles_n * tiles_m;
  const int n_tiles = mn_tiles * splits;
  const int total_kb = (K + kDbBK - 1) / kDbBK;
  const int kb_per_split = (total_kb + splits - 1) / splits;
  const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDbStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 2 * kLEpiWarps);  // the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  cluster_sync_all();  // the peer's barriers exist before anything arrives on them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kLTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b) : "memory");
      int kbg = 0;
      for (int tile = cid; tile < n_tiles; tile += n_clusters) {
        const int mn = tile % mn_tiles, split = tile / mn_tiles;
        const int m0 = (mn % tiles_m) * 256 + rank * 128, n0 = (mn / tiles_m) * kDbBN;
        const int n_eff = min(kDbBN, (N - n0 + 15) & ~15);
        const int nb0 = n0 + rank * (n_eff >> 1);  // this CTA's half of the B tile
        const int kb0 = split * kb_per_split, num_kb = max(0, min(total_kb - kb0, kb_per_split));
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % kDbStages;
          mbar_wait(&empty[s], ((kbg / kDbStages) & 1) ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * kDbStageBytes);  // both CTAs' bytes land on the leader
          unsigned char* sa = smem + (size_t)s * kDbStageBytes;
          unsigned char* sb = sa + 2 * kDbTileA;
          const int k0 = (kb0 + kb) * kDbBK;
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            if (MN_MAJOR) {  // boxes of 64 (mn) x 32 (k rows), 4 KB: two for the 128 rows of A, one for the B half
              tma_load_3d_2sm(sa + (size_t)p * kDbTileA, &maps.a, m0, k0, p, &full[s]);
              tma_load_3d_2sm(sa + (size_t)p * kDbTileA + 4096, &maps.a, m0 + 64, k0, p, &full[s]);
              tma_load_3d_2sm(sb + (size_t)p * kDbTileB, &maps.b, nb0, k0, p, &full[s]);
            } else {
              tma_load_3d_2sm(sa + (size_t)p * kDbTileA, &maps.a, k0, m0, p, &full[s]);
              tma_load_3d_2sm(sb + (size_t)p * kDbTileB, &maps.b, k0, nb0, p, &full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      int kbg = 0, it = 0;
      for (int tile = cid; tile < n_tiles; tile += n_clusters, ++it) {
        const int buf = it & 1, use = it >> 1;
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);  // the epilogue has drained this buffer's previous tile
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 2 * kDbBN);   // main accumulator
        const uint32_t tmem_c = tmem_d + (uint32_t)kDbBN;                  // correction accumulator
        const int mn = tile % mn_tiles;
        const int n0 = (mn / tiles_m) * kDbBN;
        const int n_eff = min(kDbBN, (N - n0 + 15) & ~15);
        const uint32_t idesc = LinCfg<2>::kIdescFmt | ((uint32_t)(n_eff >> 3) << 17) | ((uint32_t)(256 >> 4) << 24) |
                               (MN_MAJOR ? ((1u << 15) | (1u << 16)) : 0u);
        const int kb0 = (tile / mn_tiles) * kb_per_split, num_kb = max(0, min(total_kb - kb0, kb_per_split));
        for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
          const int s = kbg % kDbStages;
          mbar_wait(&full[s], (kbg / kDbStages) & 1);
          tcgen05_fence_after();
          unsigned char* sa = smem + (size_t)s * kDbStageBytes;
          unsigned char* sb = sa + 2 * kDbTileA;
          uint64_t ad[2], bd[2];
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const void* pa = sa + (size_t)p * kDbTileA;
            const void* pb = sb + (size_t)p * kDbTileB;
            ad[p] = MN_MAJOR ? make_sw128_mn32_desc(pa) : make_sw64_desc(pa);
            bd[p] = MN_MAJOR ? make_sw128_mn32_desc(pb) : make_sw64_desc(pb);
          }
#pragma unroll
          for (int k = 0; k < kDbBK / kLUmmaK; ++k) {
            const uint64_t o = MN_MAJOR ? (uint64_t)(k * 128) : (uint64_t)(k * 2);
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            umma_bf16_2sm(tmem_c, ad[0] + o, bd[1] + o, idesc, acc);
            umma_bf16_2sm(tmem_c, ad[1] + o, bd[0] + o, idesc, 1u);
            umma_bf16_2sm(tmem_d, ad[0] + o, bd[0] + o, idesc, acc);
          }
          umma_commit_2sm(&empty[s]);
        }
        umma_commit_2sm(&acc_full[buf]);
      }
    }
  } else {
    const int e = warp - 2;
    const int q = warp & 3;   // TMEM lane quadrant this warp may access
    const int half = e >> 2;  // 64-column half of the tile
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const float inv_a = 1.f / ep.scale_a[0];  // exact: the scales are powers of two
    const float inv_b = 1.f / ep.scale_b[0];
    int it = 0;
    for (int tile = cid; tile < n_tiles; tile += n_clusters, ++it) {
      const int buf = it & 1, use = it >> 1;
      const int mn = tile % mn_tiles, split = tile / mn_tiles;
      const int m0 = (mn % tiles_m) * 256 + rank * 128, n0 = (mn / tiles_m) * kDbBN;
      if (ep.bias != nullptr) {
        asm volatile("bar.sync 2, %0;" ::"n"(32 * kLEpiWarps) : "memory");  // previous tile's readers are done
        if (et < kDbBN) s_bias[et] = (n0 + et < N) ? ep.bias[n0 + et] : 0.f;
        asm volatile("bar.sync 2, %0;" ::"n"(32 * kLEpiWarps) : "memory");
      }
      mbar_wait(&acc_full[buf], use & 1);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * 2 * kDbBN) + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
      const int row = m0 + r, col0 = n0 + half * 64;
      const bool live = col0 < N;  // warp-uniform: a narrow remainder tile leaves the second half empty
      uint32_t v[64];
      if (live) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t w[32];
          tmem_ld32(tacc + h * 32, v + h * 32);
          tmem_ld32(tacc + kDbBN + h * 32, w);  // correction accumulator
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[h * 32 + j] = __float_as_uint(lin_combine<2>(__uint_as_float(v[h * 32 + j]), __uint_as_float(w[j]),
                                                           inv_a, inv_b));
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&acc_empty[buf]);  // the tensor core may reuse this buffer for tile it + 2
      float amax = 0.f;
      if (live && row < M) {
        float* o = ep.out + ((int64_t)split * M + row) * ep.ldo + col0;
        const bool has_bias = ep.bias != nullptr;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          if (col0 + j < N) {  // ldo % 4 == 0 and ldo >= round_up(N, 4): the whole group is inside the pitch
            float4 t = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]));
            if (has_bias) {
              const float4 b = *reinterpret_cast<const float4*>(&s_bias[half * 64 + j]);
              t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
            }
            if (ep.relu) {
              t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f);
            }
            *reinterpret_cast<float4*>(o + j) = t;
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
          }
        }
      }
      if (ep.absmax_out != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if (lane == 0 && amax > 0.f) atomicMax(ep.absmax_out, __float_as_uint(amax));
      }
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();  // both CTAs are done with TMEM and with each other's barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kLTmemCols) : "memory");
  }
}

// K-major operand: 3-D bf16 map over planes [3][rows][ld]: dims (cols, rows, 3), box 32 x box_rows x 1, 64B swizzle
static int make_map3(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int64_t plane,
                     int box_rows, int bk, int np) {
  EncodeTiledFn enc = get_encode();
  PTREC_CHECK_ARG(enc != nullptr, PTREC_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)np};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane * 2};
  cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, np == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PTREC_CHECK_ARG(r == CUDA_SUCCESS, PTREC_ECUDA, "cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
  return PTREC_OK;
}

// MN-major operand: planes [3][k_rows][ld] with the M/N dimension contiguous; box 64 (mn) x 32 (k rows) x 1
static int make_map3_mn(CUtensorMap* map, const void* base, int64_t k_rows, int64_t mn, int64_t ld, int64_t plane,
                        int bk, int np) {
  EncodeTiledFn enc = get_encode();
  PTREC_CHECK_ARG(enc != nullptr, PTREC_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)mn, (cuuint64_t)k_rows, (cuuint64_t)np};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)bk, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, np == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PTREC_CHECK_ARG(r == CUDA_SUCCESS, PTREC_ECUDA, "cuTensorMapEncodeTiled(3d, mn-major) failed (%d)", (int)r);
  return PTREC_OK;
}

// Output of the CTA-pair kernel's TMA stores: 3-D (cols, rows, slab) over [slabs][rows][ld], boxes of 32 x 32 x 1.
// fp32 result: slab = split-K partial, 128-byte swizzle; fp16 planes: slab = plane, 64-byte swizzle.
static int make_map_out(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int64_t slab,
                        int64_t slabs, bool f16, bool bf16 = false) {
  EncodeTiledFn enc = get_encode();
  PTREC_CHECK_ARG(enc != nullptr, PTREC_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  if (bf16) f16 = true;  // 16-bit elements, 64-byte swizzle; only the element type differs
  const cuuint64_t esz = f16 ? 2 : 4;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slabs};
  cuuint64_t strides[2] = {(cuuint64_t)ld * esz, (cuuint64_t)slab * esz};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : (f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32), 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PTREC_CHECK_ARG(r == CUDA_SUCCESS, PTREC_ECUDA, "cuTensorMapEncodeTiled(output) failed (%d)", (int)r);
  return PTREC_OK;
}

}  // namespace ptrec

using namespace ptrec;

constexpr int kAbsmaxMaxParts = 1024;  // CTAs of absmax_kernel (one partial maximum each)

extern "C" size_t ptrec_tc_split3_workspace_bytes(int64_t R, int64_t C) {
  return align_up((size_t)(ceil_div(R, kSpTile) * C) * sizeof(float) + 16, 256);
}
extern "C" size_t ptrec_tc_split2h_workspace_bytes(int64_t R, int64_t C) {
  return ptrec_tc_split3_workspace_bytes(R, C) + kAbsmaxMaxParts * sizeof(float);
}

// np = 3: bf16 x 3 planes; np = 2: fp16 x 2 planes of src * scale (scale_out receives the power-of-two scale)
static int split_impl(int np, const float* src, int64_t ld, int64_t R, int64_t C, const float* relu_ref, int64_t ld_ref,
                      void* planes, int64_t pl_ld, void* planes_t, int64_t pt_ld, float* colsum, float* scale_out,
                      const float* absmax_in, void* workspace, size_t workspace_bytes, void* stream,
                      const float* scale_in = nullptr, float* max_out = nullptr) {
  PTREC_CHECK_ARG(src != nullptr && (planes || planes_t || colsum), PTREC_EINVAL, "tc_split: null pointer");
  PTREC_CHECK_ARG(R >= 1 && C >= 1 && R < (1ll << 31) && C < (1ll << 31) && ld >= C, PTREC_EINVAL,
                  "tc_split: bad shape R=%lld C=%lld ld=%lld", (long long)R, (long long)C, (long long)ld);
  PTREC_CHECK_ARG(!planes || (pl_ld % 8 == 0 && pl_ld >= C && aligned16(planes)), PTREC_EALIGN,
                  "tc_split: plane pitch must be a multiple of 8 elements >= C, base 16-byte aligned");
  PTREC_CHECK_ARG(!planes_t || (pt_ld % 8 == 0 && pt_ld >= R && aligned16(planes_t)), PTREC_EALIGN,
                  "tc_split: transposed plane pitch must be a multiple of 8 elements >= R, base 16-byte aligned");
  PTREC_CHECK_ARG(!relu_ref || ld_ref >= C, PTREC_EINVAL, "tc_split: bad ld_ref");
  const size_t need = np == 2 ? ptrec_tc_split2h_workspace_bytes(R, C) : ptrec_tc_split3_workspace_bytes(R, C);
  PTREC_CHECK_ARG(!(colsum || (np == 2 && !scale_in)) || (workspace && workspace_bytes >= need), PTREC_EWORKSPACE,
                  "tc_split: workspace too small (%zu < %zu)", workspace_bytes, need);
  PTREC_CHECK_ARG(np == 3 || scale_out != nullptr || scale_in != nullptr, PTREC_EINVAL, "tc_split2h: scale_out is null");
  cudaStream_t st = (cudaStream_t)stream;
  Split3Args a;
  a.src = src; a.ld = ld; a.R = (int)R; a.C = (int)C; a.ref = relu_ref; a.ld_ref = ld_ref;
  a.planes = reinterpret_cast<unsigned short*>(planes); a.pl_ld = pl_ld; a.pl_plane = R * pl_ld;
  a.Cp = (int)std::min<int64_t>(pl_ld, (C + 7) / 8 * 8);
  a.planes_t = reinterpret_cast<unsigned short*>(planes_t); a.pt_ld = pt_ld; a.pt_plane = C * pt_ld;
  a.Rp = (int)std::min<int64_t>(pt_ld, (R + 7) / 8 * 8);
  a.colsum_part = colsum ? reinterpret_cast<float*>(workspace) : nullptr;
  a.absmax_part = nullptr; a.n_part = 0; a.scale_out = scale_out;
  a.scale_in = scale_in; a.max_out = reinterpret_cast<uint32_t*>(max_out);
  dim3 grid((unsigned)ceil_div(C, kSpTile), (unsigned)ceil_div(R, kSpTile));
  if (np == 2 && scale_in != nullptr) {
    split_kernel<2><<<grid, kSpThreads, 0, st>>>(a);
  } else if (np == 2) {
    if (absmax_in != nullptr) {  // the producer of src already reduced max |src| into one word (GEMM epilogue)
      a.absmax_part = absmax_in; a.n_part = 1;
    } else {
      float* part = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) +
                                             ptrec_tc_split3_workspace_bytes(R, C));
      const int parts = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(R, (int64_t)8), kAbsmaxMaxParts));
      absmax_kernel<<<parts, 256, 0, st>>>(src, ld, (int)R, (int)C, part);  // a warp per row, <= 8192 warps
      PTREC_LAUNCH_CHECK("absmax_kernel");
      a.absmax_part = part; a.n_part = parts;
    }
    split_kernel<2><<<grid, kSpThreads, 0, st>>>(a);
  } else {
    split_kernel<3><<<grid, kSpThreads, 0, st>>>(a);
  }
  PTREC_LAUNCH_CHECK("split_kernel");
  if (colsum) {
    colsum_reduce_kernel<<<(unsigned)ceil_div(C, 32), 256, 0, st>>>(a.colsum_part, (int)grid.y, (int)C, colsum);
    PTREC_LAUNCH_CHECK("colsum_reduce_kernel");
  }
  return PTREC_OK;
}

extern "C" int ptrec_tc_split3(const float* src, int64_t ld, int64_t R, int64_t C, const float* relu_ref,
                               int64_t ld_ref, void* planes, int64_t pl_ld, void* planes_t, int64_t pt_ld,
                               float* colsum, void* workspace, size_t workspace_bytes, void* stream) {
  return split_impl(3, src, ld, R, C, relu_ref, ld_ref, planes, pl_ld, planes_t, pt_ld, colsum, nullptr, nullptr,
                    workspace, workspace_bytes, stream);
}

extern "C" int ptrec_tc_split2h(const float* src, int64_t ld, int64_t R, int64_t C, const float* relu_ref,
                                int64_t ld_ref, void* planes, int64_t pl_ld, void* planes_t, int64_t pt_ld,
                                float* colsum, float* scale_out, const float* absmax_in, void* workspace,
                                size_t workspace_bytes, void* stream) {
  return split_impl(2, src, ld, R, C, relu_ref, ld_ref, planes, pl_ld, planes_t, pt_ld, colsum, scale_out, absmax_in,
                    workspace, workspace_bytes, stream);
}

static int g_tc_bk = 32;
extern "C" void ptrec_tc_set_bk(int32_t bk) { g_tc_bk = bk == 64 ? 64 : 32; }
static int g_tc_2sm = 1;
extern "C" void ptrec_tc_set_2sm(int32_t enabled) { g_tc_2sm = enabled ? 1 : 0; }
// fp16 x 2, CTA pairs: tile width 256 (one accumulator pair in TMEM) or 128 (two: MMAs overlap the epilogue)
static int g_tc_bn = 256;
extern "C" void ptrec_tc_set_bn(int32_t bn) { g_tc_bn = bn == 128 ? 128 : 256; }
extern "C" int32_t ptrec_tc_get_bn(void) { return g_tc_bn; }
extern "C" int32_t ptrec_tc_2sm_enabled(void) { return g_tc_2sm; }

extern "C" size_t ptrec_tc_gemm_split3_workspace_bytes(int64_t M, int64_t ldo, int32_t splits) {
  return splits > 1 ? align_up((size_t)splits * M * ldo * sizeof(float), 256) : 0;
}

// default split count for a K-heavy product (weight gradients): fill the SMs without starving a split of k-blocks
extern "C" int32_t ptrec_tc_gemm_split3_default_splits(int64_t M, int64_t N, int64_t K) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = g_tc_2sm ? ceil_div(M, 256) * ceil_div(N, 256) : ceil_div(M, kLBM) * ceil_div(N, kLBN);
  const int64_t kbs = ceil_div(K, kLBK);
  if (g_tc_2sm) sms /= 2;  // work items are CTA pairs
  int64_t s = tiles >= sms ? 1 : sms / tiles;
  s = std::max<int64_t>(1, std::min<int64_t>(s, kbs / 4 > 0 ? kbs / 4 : 1));
  return (int32_t)std::min<int64_t>(s, 64);
}

template <int NP>
static int gemm_launch(bool mn_major, bool two_sm, bool db, int bk, int sms, const LinMaps& maps, int64_t M, int64_t N,
                       int64_t K, const LinEpi& ep, float* out, cudaStream_t st) {
  static bool attr_set = false;  // one flag per plane count
  if (!attr_set) {
    PTREC_CUDA(cudaFuncSetAttribute(gemm_split3_kernel<false, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLSmem));
    PTREC_CUDA(cudaFuncSetAttribute(gemm_split3_kernel<true, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLSmem));
    PTREC_CUDA(cudaFuncSetAttribute(gemm_split3_2sm_kernel<false, 32, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
    PTREC_CUDA(cudaFuncSetAttribute(gemm_split3_2sm_kernel<true, 32, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
    PTREC_CUDA(cudaFuncSetAttribute(gemm_split3_2sm_kernel<false, 64, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
    PTREC_CUDA(cudaFuncSetAttribute(gemm_split3_2sm_kernel<true, 64, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
    if (NP == 2) {
      PTREC_CUDA(cudaFuncSetAttribute(gemm_split2h_2sm_db_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
      PTREC_CUDA(cudaFuncSetAttribute(gemm_split2h_2sm_db_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
    }
    attr_set = true;
  }
  const int splits = ep.splits;
  if (NP == 2 && db) {
    const int64_t pair_tiles = ceil_div(N, kDbBN) * ceil_div(M, 256) * splits;
    const unsigned grid2 = (unsigned)(2 * std::min<int64_t>(pair_tiles, sms / 2));
    if (mn_major)
      gemm_split2h_2sm_db_kernel<true><<<grid2, kLThreads, k2Smem, st>>>(maps, (int)M, (int)N, (int)K, ep);
    else
      gemm_split2h_2sm_db_kernel<false><<<grid2, kLThreads, k2Smem, st>>>(maps, (int)M, (int)N, (int)K, ep);
    PTREC_LAUNCH_CHECK("gemm_split2h_2sm_db_kernel");
  } else if (two_sm) {
    const int64_t pair_tiles = ceil_div(N, 256) * ceil_div(M, 256) * splits;
    const int64_t clusters = std::min<int64_t>(pair_tiles, sms / 2);
    const unsigned grid2 = (unsigned)(2 * clusters);
    if (mn_major && bk == 64)
      gemm_split3_2sm_kernel<true, 64, NP><<<grid2, k2Threads, k2Smem, st>>>(maps, (int)M, (int)N, (int)K, ep);
    else if (mn_major)
      gemm_split3_2sm_kernel<true, 32, NP><<<grid2, k2Threads, k2Smem, st>>>(maps, (int)M, (int)N, (int)K, ep);
    else if (bk == 64)
      gemm_split3_2sm_kernel<false, 64, NP><<<grid2, k2Threads, k2Smem, st>>>(maps, (int)M, (int)N, (int)K, ep);
    else
      gemm_split3_2sm_kernel<false, 32, NP><<<grid2, k2Threads, k2Smem, st>>>(maps, (int)M, (int)N, (int)K, ep);
    PTREC_LAUNCH_CHECK("gemm_split3_2sm_kernel");
  } else {
    const int64_t tiles = ceil_div(N, kLBN) * ceil_div(M, kLBM) * splits;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    if (mn_major)
      gemm_split3_kernel<true, NP><<<grid, kLThreads, kLSmem, st>>>(maps, (int)M, (int)N, (int)K, ep);
    else
      gemm_split3_kernel<false, NP><<<grid, kLThreads, kLSmem, st>>>(maps, (int)M, (int)N, (int)K, ep);
    PTREC_LAUNCH_CHECK("gemm_split3_kernel");
  }
  if (splits > 1) {
    const int64_t n = M * ep.ldo;
    partial_reduce_kernel<<<(unsigned)ceil_div(n / 4, 256), 256, 0, st>>>(ep.out, splits, n, out);
    PTREC_LAUNCH_CHECK("partial_reduce_kernel");
  }
  return PTREC_OK;
}

struct LinFused {  // ptrec_tc_gemm_split2h_fused: what the epilogue writes besides / instead of the fp32 result
  void* h2_planes = nullptr;
  int64_t h2_ld = 0;
  const float* h2_scale = nullptr;
  const uint32_t* mask_in = nullptr;
  uint32_t* mask_out = nullptr;
  int64_t mask_ld = 0;
  float* colsum = nullptr;
};

static int gemm_split_impl(int np, bool mn_major, const void* a_planes, const float* scale_a, int64_t M, int64_t lda,
                           const void* b_planes, const float* scale_b, int64_t N, int64_t ldb, int64_t K,
                           const float* bias, int32_t relu, float* out, int64_t ldo, void* out_planes,
                           int64_t out_planes_ld, float* absmax_out, int32_t splits, void* workspace,
                           size_t workspace_bytes, void* stream, const LinFused* fu = nullptr) {
  PTREC_CHECK_ARG(a_planes && b_planes && (out || (fu && fu->h2_planes)), PTREC_EINVAL, "tc_gemm_split: null pointer");
  if (out == nullptr) ldo = (N + 3) / 4 * 4;
  PTREC_CHECK_ARG(np == 3 || (scale_a && scale_b), PTREC_EINVAL, "tc_gemm_split2h: null operand scale");
  PTREC_CHECK_ARG(M >= 1 && N >= 1 && K >= 1 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), PTREC_EINVAL,
                  "tc_gemm_split: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  PTREC_CHECK_ARG(aligned16(a_planes) && aligned16(b_planes) && aligned16(out) && lda % 8 == 0 && ldb % 8 == 0 &&
                      lda >= (mn_major ? M : K) && ldb >= (mn_major ? N : K) && ldo % 4 == 0 && ldo >= (N + 3) / 4 * 4,
                  PTREC_EALIGN, "tc_gemm_split: pitches must be multiples of 8 (planes) / 4 (out) elements");
  if (splits < 1) splits = 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool two_sm = g_tc_2sm != 0 && sms >= 2;
  const bool db = np == 2 && two_sm && g_tc_bn == 128;  // double-buffered 256 x 128 pair tiles (K blocks of 32)
  const int bk = db ? kDbBK : (two_sm ? g_tc_bk : kLBK);
  const int total_kb = (int)ceil_div(K, (int64_t)bk);
  if (splits > total_kb) splits = total_kb;
  while (splits > 1 && (int64_t)(splits - 1) * ceil_div(total_kb, splits) >= total_kb) --splits;  // no empty split
  PTREC_CHECK_ARG(splits == 1 || (!bias && !relu && !out_planes && !absmax_out), PTREC_EINVAL,
                  "tc_gemm_split: split-K has no bias / ReLU / planes / absmax epilogue");
  PTREC_CHECK_ARG(!out_planes || (np == 3 && aligned16(out_planes) && out_planes_ld % 8 == 0 &&
                                  out_planes_ld >= (N + 3) / 4 * 4),
                  PTREC_EALIGN, "tc_gemm_split: output planes need bf16 x 3 and a pitch that is a multiple of 8 >= N");
  PTREC_CHECK_ARG(splits == 1 || (workspace && workspace_bytes >= ptrec_tc_gemm_split3_workspace_bytes(M, ldo, splits)),
                  PTREC_EWORKSPACE, "tc_gemm_split: workspace too small for %d split-K partials", splits);
  cudaStream_t st = (cudaStream_t)stream;
  LinMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = mn_major ? make_map3_mn(&maps.a, a_planes, K, M, lda, K * lda, bk, np)
                    : make_map3(&maps.a, a_planes, M, K, lda, M * lda, kLBM, bk, np);
  if (rc != PTREC_OK) return rc;
  rc = mn_major ? make_map3_mn(&maps.b, b_planes, K, N, ldb, K * ldb, bk, np)
                : make_map3(&maps.b, b_planes, N, K, ldb, N * ldb, db ? 64 : (two_sm ? 128 : kLBN), bk, np);
  if (rc != PTREC_OK) return rc;
  LinEpi ep;
  ep.bias = bias; ep.relu = relu; ep.ldo = ldo; ep.splits = splits;
  ep.planes = reinterpret_cast<__nv_bfloat16*>(out_planes); ep.pl_ld = out_planes_ld; ep.pl_plane = M * out_planes_ld;
  ep.out = splits > 1 ? reinterpret_cast<float*>(workspace) : out;
  ep.scale_a = scale_a; ep.scale_b = scale_b;
  ep.absmax_out = reinterpret_cast<uint32_t*>(absmax_out);
  ep.h2_planes = nullptr; ep.h2_ld = 0; ep.h2_plane = 0; ep.h2_scale = nullptr;
  ep.mask_in = nullptr; ep.mask_out = nullptr; ep.mask_ld = 0; ep.colsum_part = nullptr;
  ep.dcn_mode = 0; ep.dp0 = nullptr; ep.dp1 = nullptr; ep.dld = 0; ep.dcn_o1 = 0;
  if (two_sm && !db && ep.out != nullptr) {  // the CTA-pair kernel stores through TMA
    rc = make_map_out(&maps.c, ep.out, M, N, ldo, M * ldo, splits, false);
    if (rc != PTREC_OK) return rc;
  }
  if (fu != nullptr) {
    PTREC_CHECK_ARG(np == 2 && two_sm && !db && !mn_major && splits == 1, PTREC_EUNSUPPORTED,
                    "tc_gemm_split2h_fused needs the CTA-pair fp16 x 2 kernel (256-wide tiles), K-major operands, no split-K");
    PTREC_CHECK_ARG(!fu->h2_planes || (fu->h2_scale && aligned16(fu->h2_planes) && fu->h2_ld % 8 == 0 && fu->h2_ld >= N),
                    PTREC_EALIGN, "tc_gemm_split2h_fused: output planes need a scale word and a pitch that is a multiple of 8 >= N");
    PTREC_CHECK_ARG(!(fu->mask_in || fu->mask_out) || (fu->mask_ld % 4 == 0 && fu->mask_ld * 32 >= N &&
                                                        (!fu->mask_in || aligned16(fu->mask_in)) &&
                                                        (!fu->mask_out || aligned16(fu->mask_out))),
                    PTREC_EALIGN, "tc_gemm_split2h_fused: mask pitch must be a multiple of 4 words covering N, base 16-byte aligned");
    const size_t need = align_up((size_t)ceil_div(M, (int64_t)32) * N * sizeof(float), 256);
    PTREC_CHECK_ARG(!fu->colsum || (workspace && workspace_bytes >= need), PTREC_EWORKSPACE,
                    "tc_gemm_split2h_fused: workspace too small for the column-sum partials (%zu < %zu)", workspace_bytes, need);
    ep.h2_planes = reinterpret_cast<unsigned short*>(fu->h2_planes); ep.h2_ld = fu->h2_ld; ep.h2_plane = M * fu->h2_ld;
    ep.h2_scale = fu->h2_scale; ep.mask_in = fu->mask_in; ep.mask_out = fu->mask_out; ep.mask_ld = fu->mask_ld;
    ep.colsum_part = fu->colsum ? reinterpret_cast<float*>(workspace) : nullptr;
    if (ep.h2_planes != nullptr) {
      rc = make_map_out(&maps.p, ep.h2_planes, M, N, ep.h2_ld, ep.h2_plane, 2, true);
      if (rc != PTREC_OK) return rc;
    }
    const int rc2 = gemm_launch<2>(false, true, false, bk, sms, maps, M, N, K, ep, out, st);
    if (rc2 != PTREC_OK) return rc2;
    if (fu->colsum) {
      cudaStream_t rs = reduce_stream_after(st);
      colsum_reduce_kernel<<<(unsigned)ceil_div(N, (int64_t)32), 256, 0, rs>>>(ep.colsum_part, (int)ceil_div(M, (int64_t)32),
                                                                                (int)N, fu->colsum);
      PTREC_LAUNCH_CHECK("colsum_reduce_kernel");
    }
    return PTREC_OK;
  }
  return np == 3 ? gemm_launch<3>(mn_major, two_sm, false, bk, sms, maps, M, N, K, ep, out, st)
                 : gemm_launch<2>(mn_major, two_sm, db, bk, sms, maps, M, N, K, ep, out, st);
}

extern "C" int ptrec_tc_gemm_split3(const void* a_planes, int64_t M, int64_t lda, const void* b_planes, int64_t N,
                                    int64_t ldb, int64_t K, const float* bias, int32_t relu, float* out, int64_t ldo,
                                    void* out_planes, int64_t out_planes_ld, int32_t splits, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  return gemm_split_impl(3, false, a_planes, nullptr, M, lda, b_planes, nullptr, N, ldb, K, bias, relu, out, ldo,
                         out_planes, out_planes_ld, nullptr, splits, workspace, workspace_bytes, stream);
}

extern "C" int ptrec_tc_gemm_split3_tn(const void* a_planes, int64_t M, int64_t lda, const void* b_planes, int64_t N,
                                       int64_t ldb, int64_t K, float* out, int64_t ldo, int32_t splits, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  return gemm_split_impl(3, true, a_planes, nullptr, M, lda, b_planes, nullptr, N, ldb, K, nullptr, 0, out, ldo, nullptr,
                         0, nullptr, splits, workspace, workspace_bytes, stream);
}

extern "C" int ptrec_tc_gemm_split2h(const void* a_planes, const float* scale_a, int64_t M, int64_t lda,
                                     const void* b_planes, const float* scale_b, int64_t N, int64_t ldb, int64_t K,
                                     const float* bias, int32_t relu, float* out, int64_t ldo, float* absmax_out,
                                     int32_t splits, void* workspace, size_t workspace_bytes, void* stream) {
  return gemm_split_impl(2, false, a_planes, scale_a, M, lda, b_planes, scale_b, N, ldb, K, bias, relu, out, ldo,
                         nullptr, 0, absmax_out, splits, workspace, workspace_bytes, stream);
}

extern "C" int ptrec_tc_gemm_split2h_tn(const void* a_planes, const float* scale_a, int64_t M, int64_t lda,
                                        const void* b_planes, const float* scale_b, int64_t N, int64_t ldb, int64_t K,
                                        float* out, int64_t ldo, int32_t splits, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  return gemm_split_impl(2, true, a_planes, scale_a, M, lda, b_planes, scale_b, N, ldb, K, nullptr, 0, out, ldo,
                         nullptr, 0, nullptr, splits, workspace, workspace_bytes, stream);
}

// ---- NP = 1: the CTA-pair kernel as a plain bf16 GEMM (K5, the DCN-v2 cross layers; declared in gemm2sm.cuh) ----
namespace ptrec {

int gemm_bf16_2sm(const Bf16Gemm& g, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  PTREC_CHECK_ARG(sms >= 2, PTREC_EUNSUPPORTED, "gemm_bf16_2sm needs CTA pairs");
  PTREC_CHECK_ARG(g.A && g.B && g.M >= 1 && g.N >= 1 && g.K >= 1 && aligned16(g.A) && aligned16(g.B) && g.lda % 8 == 0 &&
                      g.ldb % 8 == 0, PTREC_EALIGN, "gemm_bf16_2sm: operands");
  static bool attr_set = false;
  if (!attr_set) {
    PTREC_CUDA(cudaFuncSetAttribute(gemm_split3_2sm_kernel<false, 32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
    PTREC_CUDA(cudaFuncSetAttribute(gemm_split3_2sm_kernel<true, 32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
    attr_set = true;
  }
  const int bk = 32;
  LinMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = g.mn_major ? make_map3_mn(&maps.a, g.A, g.K, g.M, g.lda, g.K * g.lda, bk, 1)
                      : make_map3(&maps.a, g.A, g.M, g.K, g.lda, g.M * g.lda, kLBM, bk, 1);
  if (rc != PTREC_OK) return rc;
  rc = g.mn_major ? make_map3_mn(&maps.b, g.B, g.K, g.N, g.ldb, g.K * g.ldb, bk, 1)
                  : make_map3(&maps.b, g.B, g.N, g.K, g.ldb, g.N * g.ldb, 128, bk, 1);
  if (rc != PTREC_OK) return rc;
  int splits = g.splits < 1 ? 1 : g.splits;
  const int total_kb = (int)ceil_div(g.K, (int64_t)bk);
  if (splits > total_kb) splits = total_kb;
  while (splits > 1 && (int64_t)(splits - 1) * ceil_div(total_kb, splits) >= total_kb) --splits;
  LinEpi ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = g.bias; ep.relu = 0; ep.splits = splits; ep.ldo = g.ldo;
  ep.dcn_mode = g.dcn_mode; ep.dp0 = reinterpret_cast<const __nv_bfloat16*>(g.p0);
  ep.dp1 = reinterpret_cast<const __nv_bfloat16*>(g.p1); ep.dld = g.pld; ep.dcn_o1 = g.o1 != nullptr;
  if (g.dcn_mode != 0) {
    PTREC_CHECK_ARG(!g.mn_major && splits == 1 && g.p0 && g.p1 && g.o0 && g.pld % 8 == 0 && g.N % 8 == 0 && aligned16(g.p0) &&
                        aligned16(g.p1) && aligned16(g.o0) && (!g.o1 || aligned16(g.o1)), PTREC_EALIGN,
                    "gemm_bf16_2sm: cross-layer epilogue operands");
    if ((rc = make_map_out(&maps.p, g.o0, g.M, g.N, g.pld, g.M * g.pld, 1, false, true)) != PTREC_OK) return rc;
    if (g.o1 && (rc = make_map_out(&maps.q, g.o1, g.M, g.N, g.pld, g.M * g.pld, 1, false, true)) != PTREC_OK) return rc;
    if ((rc = make_map_out(&maps.r0, g.p0, g.M, g.N, g.pld, g.M * g.pld, 1, false, true)) != PTREC_OK) return rc;
    if ((rc = make_map_out(&maps.r1, g.p1, g.M, g.N, g.pld, g.M * g.pld, 1, false, true)) != PTREC_OK) return rc;
  } else {
    PTREC_CHECK_ARG(g.out && aligned16(g.out) && g.ldo % 4 == 0 && g.ldo >= (g.N + 3) / 4 * 4, PTREC_EALIGN,
                    "gemm_bf16_2sm: fp32 output");
    PTREC_CHECK_ARG(splits == 1 || (g.workspace && g.workspace_bytes >= (size_t)splits * g.M * g.ldo * sizeof(float)),
                    PTREC_EWORKSPACE, "gemm_bf16_2sm: workspace too small for %d split-K partials", splits);
    ep.out = splits > 1 ? reinterpret_cast<float*>(g.workspace) : g.out;
    if ((rc = make_map_out(&maps.c, ep.out, g.M, g.N, g.ldo, g.M * g.ldo, splits, false)) != PTREC_OK) return rc;
  }
  const int64_t pair_tiles = ceil_div(g.N, (int64_t)256) * ceil_div(g.M, (int64_t)256) * splits;
  const unsigned grid2 = (unsigned)(2 * std::min<int64_t>(pair_tiles, sms / 2));
  if (g.mn_major)
    gemm_split3_2sm_kernel<true, 32, 1><<<grid2, k2Threads, k2Smem, st>>>(maps, (int)g.M, (int)g.N, (int)g.K, ep);
  else
    gemm_split3_2sm_kernel<false, 32, 1><<<grid2, k2Threads, k2Smem, st>>>(maps, (int)g.M, (int)g.N, (int)g.K, ep);
  PTREC_LAUNCH_CHECK("gemm_split3_2sm_kernel<.,32,1>");
  if (g.dcn_mode == 0 && splits > 1) {
    const int64_t n = g.M * g.ldo;
    partial_reduce_kernel<<<(unsigned)ceil_div(n / 4, (int64_t)256), 256, 0, st>>>(ep.out, splits, n, g.out);
    PTREC_LAUNCH_CHECK("partial_reduce_kernel");
  }
  return PTREC_OK;
}

}  // namespace ptrec

// ---- the fused tower: carried scales, prescaled split, GEMM epilogue that writes its consumer's operand planes ----
extern "C" int ptrec_tc_scale_roll(float* slots, int32_t n_slots, float* call_scales, int32_t* err, void* stream) {
  PTREC_CHECK_ARG(slots && call_scales && err && n_slots >= 1, PTREC_EINVAL, "tc_scale_roll: null pointer");
  scale_roll_kernel<<<(unsigned)ceil_div(n_slots, 128), 128, 0, (cudaStream_t)stream>>>(slots, n_slots, call_scales, err);
  PTREC_LAUNCH_CHECK("scale_roll_kernel");
  return PTREC_OK;
}

extern "C" int ptrec_tc_split2h_prescaled(const float* src, int64_t ld, int64_t R, int64_t C, const float* relu_ref,
                                          int64_t ld_ref, void* planes, int64_t pl_ld, void* planes_t, int64_t pt_ld,
                                          float* colsum, const float* scale_in, float* max_out, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  PTREC_CHECK_ARG(scale_in != nullptr, PTREC_EINVAL, "tc_split2h_prescaled: scale_in is null");
  return split_impl(2, src, ld, R, C, relu_ref, ld_ref, planes, pl_ld, planes_t, pt_ld, colsum, nullptr, nullptr,
                    workspace, workspace_bytes, stream, scale_in, max_out);
}

extern "C" size_t ptrec_tc_gemm_fused_workspace_bytes(int64_t M, int64_t N) {
  return align_up((size_t)ceil_div(M, (int64_t)32) * N * sizeof(float), 256);
}

extern "C" int ptrec_tc_gemm_split2h_fused(const void* a_planes, const float* scale_a, int64_t M, int64_t lda,
                                           const void* b_planes, const float* scale_b, int64_t N, int64_t ldb, int64_t K,
                                           const float* bias, int32_t relu, float* out, int64_t ldo, void* out_planes,
                                           int64_t out_planes_ld, const float* out_scale, const uint32_t* mask_in,
                                           uint32_t* mask_out, int64_t mask_ld, float* colsum, float* absmax_out,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  LinFused fu;
  fu.h2_planes = out_planes; fu.h2_ld = out_planes_ld; fu.h2_scale = out_scale;
  fu.mask_in = mask_in; fu.mask_out = mask_out; fu.mask_ld = mask_ld; fu.colsum = colsum;
  return gemm_split_impl(2, false, a_planes, scale_a, M, lda, b_planes, scale_b, N, ldb, K, bias, relu, out, ldo, nullptr,
                         0, absmax_out, 1, workspace, workspace_bytes, stream, &fu);
}

#ifdef PTREC_K6_TIMELINE
extern "C" int ptrec_debug_k6_timeline(unsigned long long* buf) {
  return cudaMemcpyToSymbol(ptrec::g_k6_timeline, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif
