// K4 on the tensor cores: the DIN activation unit's two hidden layers as tcgen05 MMAs over tiles of 128 history
// positions (~27 kFLOP per position against 128 bytes of keys — on CUDA cores the unit is compute-bound
// at 12 TFLOP/s, 2 % of the HBM roofline; din_attn.cu keeps that SIMT kernel for the other shapes and as the
// cross-check).  No reference counterpart (the reference has no DIN); conventions of the history inputs follow
// torchrec/model/SASRec.py:95-110 and utils.py:5-10 (right-padded [B, L], a length column).
//
//   per sample b:   M_b = (W1k - W1d) + W1p diag(q)  [H1, DQ],   c_b = (W1q + W1d) q + b1          (algebraic fusion)
//   per tile of 128 positions l:
//     G1   H1pre [128, H1]  = K_tile [128, DQ] . M_b^T             tcgen05.mma, M = 128, N = H1, K = DQ
//          h1 = relu(H1pre + c_b)                                   TMEM -> registers -> fp16 planes in shared memory
//     G2   H2pre [128, H2p] = h1 [128, H1] . W2^T                   tcgen05.mma, M = 128, N = 48 (H2 = 40 padded), K = H1
//          a_l = W3 relu(H2pre + b2) + b3 ;  pooled += sum_l a_l k_l (fp32 keys)
// fp32-faithful operands (the north star asks 1e-5): every operand tile is multiplied by a power of two that puts its
// largest magnitude in [2^13, 2^14) and split into two fp16 planes x s = h0 + h1 / 2^11 (22 mantissa bits; the K6 scheme,
// tc_linear.cu); a product is three MMAs — A0 B0 into a "main" accumulator, A0 B1 + A1 B0 into a "correction" one — and
// the epilogue computes (main + corr / 2^11) / (s_a s_b) in fp32.
// Operand tiles are written by the threads themselves (keys come from a gather, M_b and h1 are computed here), in the
// un-swizzled core-matrix layout: element (r, k) of a [R, K] K-major tile at
//     (r / 8) * (K / 8) * 128 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2   bytes
// i.e. 8 x 8 core matrices of 128 contiguous bytes, LBO = 128 (next core matrix along K), SBO = K / 8 * 128 (next 8 rows).
#include <cuda_fp16.h>

#include "tcgen05.cuh"

namespace ptrec {

constexpr int kTcPos = 128;  // positions per tile = MMA M = threads per CTA (thread t owns position t: TMEM lane t)

__device__ __forceinline__ uint64_t make_nosw_desc(const void* smem_tile, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t addr = smem_u32(smem_tile);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  return d;                // layout type 0: no swizzle
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 16 columns of two accumulators (main, correction) with ONE wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float* va, float* vb) {
  uint32_t a[16], b[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]),
        "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
      : "r"(ta));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]),
        "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
      : "r"(tb));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    va[i] = __uint_as_float(a[i]);
    vb[i] = __uint_as_float(b[i]);
  }
}
// power of two s with s * amax in [2^13, 2^14)  (1 for an all-zero tile)
__device__ __forceinline__ float h2_scale(float amax) {
  if (!(amax > 0.f)) return 1.f;
  const int e = (int)((__float_as_uint(amax) >> 23) & 0xffu) - 127;  // floor(log2 amax) for normal numbers
  return __uint_as_float((uint32_t)(127 + 13 - max(e, -100)) << 23);
}
// 8 consecutive fp32 values (already multiplied by the tile scale) -> one 16-byte chunk of each plane
__device__ __forceinline__ void h2_split8(const float* x, uint4* p0, uint4* p1) {
  __half2 a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half h00 = __float2half_rn(x[2 * i]), h01 = __float2half_rn(x[2 * i + 1]);
    const __half h10 = __float2half_rn((x[2 * i] - __half2float(h00)) * 2048.f);
    const __half h11 = __float2half_rn((x[2 * i + 1] - __half2float(h01)) * 2048.f);
    a[i] = __halves2half2(h00, h01);
    b[i] = __halves2half2(h10, h11);
  }
  *p0 = *reinterpret_cast<uint4*>(a);
  *p1 = *reinterpret_cast<uint4*>(b);
}
__device__ __forceinline__ float block_max_128(float v, float* s_red) {  // 4 warps; s_red: 4 floats; all threads call
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();  // s_red may still be read from the previous call
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  return fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
}

template <int DQ, int H1, int H2>
struct DinTcSmem {
  static constexpr int H2P = (H2 + 15) / 16 * 16;  // MMA N of the second layer
  // fp16 operand planes, core-matrix layout (see the file comment); 1024-byte aligned as a block
  alignas(1024) unsigned char k[2][kTcPos * DQ * 2];
  alignas(128) unsigned char m[2][H1 * DQ * 2];
  alignas(128) unsigned char h1[2][kTcPos * H1 * 2];
  alignas(128) unsigned char w2[2][H2P * H1 * 2];
  // fp32 side (W1 itself stays in global memory: 40 KB shared by every CTA, L2-resident — keeping the three derived
  // [H1, DQ] matrices here cost 30 KB and the second resident CTA per SM)
  alignas(16) float b1[H1];
  alignas(16) float b2[H2P];
  alignas(16) float W3[H2P];
  alignas(16) float c[H1];
  alignas(16) float q[DQ];
  alignas(16) float kf[kTcPos][DQ + 1];
  float a[kTcPos];
  float red[kTcPos / DQ][DQ];
  float mx[4];
  float b3;
  float s_w2;                      // scale of the W2 planes
  alignas(8) uint64_t bar[2];
  uint32_t tmem_slot;
};

template <int DQ, int H1, int H2>
__global__ void __launch_bounds__(kTcPos, 2)
din_fwd_tc_kernel(const float* __restrict__ q, int64_t q_stride, const float* __restrict__ keys, int64_t ksb, int64_t ksl,
                  const int32_t* __restrict__ lens, int64_t B, int L, const float* __restrict__ W1,
                  const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                  const float* __restrict__ W3, const float* __restrict__ b3, float* __restrict__ out,
                  float* __restrict__ scores) {
  using S = DinTcSmem<DQ, H1, H2>;
  constexpr int H2P = S::H2P;
  static_assert(DQ % 16 == 0 && H1 % 16 == 0 && H1 <= 128 && DQ <= 64, "MMA shape constraints");
  constexpr uint32_t kColsD1 = H1, kColsD2 = H2P;
  constexpr uint32_t kTmemCols = 256;  // D1 main | D1 corr | D2 main | D2 corr
  static_assert(2 * kColsD1 + 2 * kColsD2 <= kTmemCols, "accumulators exceed the TMEM allocation");
  extern __shared__ unsigned char smem_raw[];
  S* s = reinterpret_cast<S*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int t = threadIdx.x, warp = t >> 5;

  // ---- once per CTA: weights (fp32 side), W2 planes, barriers, TMEM -------------------------------------------------
  for (int e = t; e < H1; e += kTcPos) s->b1[e] = b1[e];
  for (int e = t; e < H2P; e += kTcPos) {
    s->b2[e] = e < H2 ? b2[e] : 0.f;
    s->W3[e] = e < H2 ? W3[e] : 0.f;
  }
  float wmax = 0.f;
  for (int e = t; e < H2 * H1; e += kTcPos) wmax = fmaxf(wmax, fabsf(W2[e]));
  wmax = block_max_128(wmax, s->mx);
  const float sw2 = h2_scale(wmax);
  if (t == 0) {
    s->b3 = b3[0];
    s->s_w2 = sw2;
    mbar_init(&s->bar[0], 1);
    mbar_init(&s->bar[1], 1);
    fence_mbar_init();
  }
  for (int ch = t; ch < H2P * (H1 / 8); ch += kTcPos) {  // 16-byte chunks of the W2 planes: row m, columns 8*cj..
    const int mrow = ch / (H1 / 8), cj = ch - mrow * (H1 / 8);
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = mrow < H2 ? W2[mrow * H1 + cj * 8 + i] * sw2 : 0.f;
    const uint32_t off = (uint32_t)(mrow >> 3) * (H1 / 8) * 128 + cj * 128 + (mrow & 7) * 16;
    h2_split8(x, reinterpret_cast<uint4*>(s->w2[0] + off), reinterpret_cast<uint4*>(s->w2[1] + off));
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s->tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s->tmem_slot;
  const uint32_t tD1 = tmem, tC1 = tmem + kColsD1, tD2 = tmem + 2 * kColsD1, tC2 = tD2 + kColsD2;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;  // this warp's TMEM lanes
  // instruction descriptors: D fp32 (bit 4), A / B fp16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  constexpr uint32_t kIdesc1 = (1u << 4) | ((uint32_t)(H1 >> 3) << 17) | ((uint32_t)(kTcPos >> 4) << 24);
  constexpr uint32_t kIdesc2 = (1u << 4) | ((uint32_t)(H2P >> 3) << 17) | ((uint32_t)(kTcPos >> 4) << 24);
  uint32_t phase = 0;

  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const int len = lens ? min(max(lens[b], 0), L) : L;
    __syncthreads();  // the previous sample's readers of q / c / red are done
    if (t < DQ) s->q[t] = q[b * q_stride + t];
    __syncthreads();
    // ---- per-sample operand B of G1: M_b planes, and c_b ---------------------------------------------------------
    float mv[(H1 * DQ / 8 + kTcPos - 1) / kTcPos][8];
    float mmax = 0.f;
#pragma unroll
    for (int r = 0; r < (H1 * DQ / 8 + kTcPos - 1) / kTcPos; ++r) {
      const int ch = t + r * kTcPos;  // chunk: row j, columns 8*ci..
      if (ch < H1 * DQ / 8) {
        const int j = ch / (DQ / 8), ci = ch - j * (DQ / 8);
        const float* row = W1 + (int64_t)j * 4 * DQ + ci * 8;  // [W1q | W1k | W1d | W1p], DQ columns each
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 wk = __ldg(reinterpret_cast<const float4*>(row + DQ + 4 * h));
          const float4 wd = __ldg(reinterpret_cast<const float4*>(row + 2 * DQ + 4 * h));
          const float4 wp = __ldg(reinterpret_cast<const float4*>(row + 3 * DQ + 4 * h));
          const float* qq = &s->q[ci * 8 + 4 * h];
          mv[r][4 * h + 0] = (wk.x - wd.x) + wp.x * qq[0];
          mv[r][4 * h + 1] = (wk.y - wd.y) + wp.y * qq[1];
          mv[r][4 * h + 2] = (wk.z - wd.z) + wp.z * qq[2];
          mv[r][4 * h + 3] = (wk.w - wd.w) + wp.w * qq[3];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) mmax = fmaxf(mmax, fabsf(mv[r][i]));
      }
    }
    if (t < H1) {
      float acc = s->b1[t];
      const float* row = W1 + (int64_t)t * 4 * DQ;
#pragma unroll
      for (int i = 0; i < DQ; i += 4) {
        const float4 wq = __ldg(reinterpret_cast<const float4*>(row + i));
        const float4 wd = __ldg(reinterpret_cast<const float4*>(row + 2 * DQ + i));
        acc += (wq.x + wd.x) * s->q[i] + (wq.y + wd.y) * s->q[i + 1] + (wq.z + wd.z) * s->q[i + 2] +
               (wq.w + wd.w) * s->q[i + 3];
      }
      s->c[t] = acc;
    }
    mmax = block_max_128(mmax, s->mx);
    const float sm = h2_scale(mmax);
#pragma unroll
    for (int r = 0; r < (H1 * DQ / 8 + kTcPos - 1) / kTcPos; ++r) {
      const int ch = t + r * kTcPos;
      if (ch < H1 * DQ / 8) {
        const int j = ch / (DQ / 8), ci = ch - j * (DQ / 8);
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = mv[r][i] * sm;
        const uint32_t off = (uint32_t)(j >> 3) * (DQ / 8) * 128 + ci * 128 + (j & 7) * 16;
        h2_split8(x, reinterpret_cast<uint4*>(s->m[0] + off), reinterpret_cast<uint4*>(s->m[1] + off));
      }
    }
    float acc_out = 0.f;  // this thread's share of pooled[b][t % DQ]
    const int oi = t % DQ, opart = t / DQ;
    for (int l0 = 0; l0 < max(len, 1); l0 += kTcPos) {
      const int n = max(0, min(kTcPos, len - l0));
      // ---- operand A of G1: this thread's key row (zeros past the end of the history) ----------------------------
      float kk[DQ];
      float kmax = 0.f;
      if (t < n) {
        const float* kp = keys + b * ksb + (int64_t)(l0 + t) * ksl;
#pragma unroll
        for (int x = 0; x < DQ; x += 4) {
          const float4 v = ldg_stream_f4(kp + x);
          kk[x] = v.x; kk[x + 1] = v.y; kk[x + 2] = v.z; kk[x + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int x = 0; x < DQ; ++x) kk[x] = 0.f;
      }
#pragma unroll
      for (int x = 0; x < DQ; ++x) {
        kmax = fmaxf(kmax, fabsf(kk[x]));
        s->kf[t][x] = kk[x];
      }
      kmax = block_max_128(kmax, s->mx);
      const float sk = h2_scale(kmax);
#pragma unroll
      for (int ci = 0; ci < DQ / 8; ++ci) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = kk[ci * 8 + i] * sk;
        const uint32_t off = (uint32_t)(t >> 3) * (DQ / 8) * 128 + ci * 128 + (t & 7) * 16;
        h2_split8(x, reinterpret_cast<uint4*>(s->k[0] + off), reinterpret_cast<uint4*>(s->k[1] + off));
      }
      // the planes were written through the generic proxy: make them visible to the tensor core (async proxy)
      fence_proxy_async();
      tcgen05_fence_before();
      __syncthreads();
      if (t == 0) {
        tcgen05_fence_after();
#pragma unroll
        for (int ks = 0; ks < DQ / 16; ++ks) {  // one MMA consumes K = 16: two core matrices = 256 bytes
          const uint64_t a0 = make_nosw_desc(s->k[0] + ks * 256, 128, (DQ / 8) * 128);
          const uint64_t a1 = make_nosw_desc(s->k[1] + ks * 256, 128, (DQ / 8) * 128);
          const uint64_t b0 = make_nosw_desc(s->m[0] + ks * 256, 128, (DQ / 8) * 128);
          const uint64_t b1d = make_nosw_desc(s->m[1] + ks * 256, 128, (DQ / 8) * 128);
          umma_bf16(tC1, a0, b1d, kIdesc1, ks > 0 ? 1u : 0u);
          umma_bf16(tC1, a1, b0, kIdesc1, 1u);
          umma_bf16(tD1, a0, b0, kIdesc1, ks > 0 ? 1u : 0u);
        }
        umma_commit(&s->bar[0]);
      }
      mbar_wait(&s->bar[0], phase);
      tcgen05_fence_after();
      // ---- epilogue 1: h1 = relu(H1pre + c) for this thread's position; planes of h1 ------------------------------
      // (two passes over TMEM — the maximum of the tile first, the planes second — instead of H1 live registers)
      const float inv1 = 1.f / (sk * sm);
      float hmax = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < H1; c0 += 16) {  // tcgen05.ld is warp-collective: every lane executes it, rows >= n are masked
        float dm[16], dc[16];
        tmem_ld16x2(tD1 + lane_base + c0, tC1 + lane_base + c0, dm, dc);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float h = (dm[i] + dc[i] * (1.f / 2048.f)) * inv1 + s->c[c0 + i];
          hmax = fmaxf(hmax, t < n ? h : 0.f);
        }
      }
      hmax = block_max_128(hmax, s->mx);  // relu: only positive values matter (the initial 0 covers all-negative tiles)
      const float sh = h2_scale(hmax);
#pragma unroll
      for (int c0 = 0; c0 < H1; c0 += 16) {
        float dm[16], dc[16], x[16];
        tmem_ld16x2(tD1 + lane_base + c0, tC1 + lane_base + c0, dm, dc);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float h = fmaxf((dm[i] + dc[i] * (1.f / 2048.f)) * inv1 + s->c[c0 + i], 0.f);
          x[i] = t < n ? h * sh : 0.f;
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cj = c0 / 8 + hh;
          const uint32_t off = (uint32_t)(t >> 3) * (H1 / 8) * 128 + cj * 128 + (t & 7) * 16;
          h2_split8(x + 8 * hh, reinterpret_cast<uint4*>(s->h1[0] + off), reinterpret_cast<uint4*>(s->h1[1] + off));
        }
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncthreads();
      if (t == 0) {
        tcgen05_fence_after();
#pragma unroll
        for (int ks = 0; ks < H1 / 16; ++ks) {
          const uint64_t a0 = make_nosw_desc(s->h1[0] + ks * 256, 128, (H1 / 8) * 128);
          const uint64_t a1 = make_nosw_desc(s->h1[1] + ks * 256, 128, (H1 / 8) * 128);
          const uint64_t b0 = make_nosw_desc(s->w2[0] + ks * 256, 128, (H1 / 8) * 128);
          const uint64_t b1d = make_nosw_desc(s->w2[1] + ks * 256, 128, (H1 / 8) * 128);
          umma_bf16(tC2, a0, b1d, kIdesc2, ks > 0 ? 1u : 0u);
          umma_bf16(tC2, a1, b0, kIdesc2, 1u);
          umma_bf16(tD2, a0, b0, kIdesc2, ks > 0 ? 1u : 0u);
        }
        umma_commit(&s->bar[1]);
      }
      mbar_wait(&s->bar[1], phase);
      tcgen05_fence_after();
      phase ^= 1u;
      // ---- epilogue 2: a_l, then pooled += sum_l a_l k_l ---------------------------------------------------------
      const float inv2 = 1.f / (sh * s->s_w2);
      float a = s->b3;
#pragma unroll
      for (int c0 = 0; c0 < H2P; c0 += 16) {
        float dm[16], dc[16];
        tmem_ld16x2(tD2 + lane_base + c0, tC2 + lane_base + c0, dm, dc);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          a += s->W3[c0 + i] * fmaxf((dm[i] + dc[i] * (1.f / 2048.f)) * inv2 + s->b2[c0 + i], 0.f);
      }
      if (t >= n) a = 0.f;
      if (scores && t < n) scores[b * L + l0 + t] = a;
      s->a[t] = a;
      tcgen05_fence_before();
      __syncthreads();
      for (int ll = opart; ll < n; ll += kTcPos / DQ) acc_out += s->a[ll] * s->kf[ll][oi];
      // (the next tile's writes of kf / a come after its first block_max_128 barrier... which is BEFORE them: sync here)
      __syncthreads();
    }
    if (scores) {
      for (int l = len + t; l < L; l += kTcPos) scores[b * L + l] = 0.f;
    }
    s->red[opart][oi] = acc_out;
    __syncthreads();
    if (t < DQ) {
      float r = 0.f;
#pragma unroll
      for (int p = 0; p < kTcPos / DQ; ++p) r += s->red[p][t];
      out[b * DQ + t] = r;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

template <int DQ, int H1, int H2>
static int din_fwd_tc_launch(const float* q, int64_t qs, const float* keys, int64_t ksb, int64_t ksl, const int32_t* lens,
                             int64_t B, int L, const float* W1, const float* b1, const float* W2, const float* b2,
                             const float* W3, const float* b3, float* out, float* scores, cudaStream_t st) {
  const size_t smem = sizeof(DinTcSmem<DQ, H1, H2>) + 1024;
  PTREC_CUDA(cudaFuncSetAttribute(din_fwd_tc_kernel<DQ, H1, H2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(B < (int64_t)sms * 2 ? B : (int64_t)sms * 2);  // two CTAs per SM (~100 KB of shared memory each)
  din_fwd_tc_kernel<DQ, H1, H2><<<grid, kTcPos, smem, st>>>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3,
                                                             out, scores);
  PTREC_LAUNCH_CHECK("din_fwd_tc_kernel");
  return PTREC_OK;
}

// entry used by din_attn.cu's dispatcher: PTREC_OK, or PTREC_EUNSUPPORTED when this shape has no tensor-core build
int din_fwd_tc(const float* q, int64_t qs, const float* keys, int64_t ksb, int64_t ksl, const int32_t* lens, int64_t B,
               int L, int DQ, int H1, int H2, const float* W1, const float* b1, const float* W2, const float* b2,
               const float* W3, const float* b3, float* out, float* scores, cudaStream_t st) {
  if (DQ == 32 && H1 == 80 && H2 == 40)
    return din_fwd_tc_launch<32, 80, 40>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3, out, scores, st);
  if (DQ == 32 && H1 == 64 && H2 == 32)
    return din_fwd_tc_launch<32, 64, 32>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3, out, scores, st);
  if (DQ == 16 && H1 == 80 && H2 == 40)
    return din_fwd_tc_launch<16, 80, 40>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3, out, scores, st);
  if (DQ == 16 && H1 == 64 && H2 == 32)
    return din_fwd_tc_launch<16, 64, 32>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3, out, scores, st);
  return PTREC_EUNSUPPORTED;
}

}  // namespace ptrec
