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

#include "din_keys.cuh"
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
                  float* __restrict__ scores, const DinKeyIds kid) {
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
      if (t < n && kid.ids[0] != nullptr) {  // gather fused in: the two halves of the key are table rows
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int64_t id = kid.ids[p][b * kid.ids_sb + kid.ids_off + l0 + t];
          const bool ok = (uint64_t)id < (uint64_t)kid.rows[p];
          if (!ok && kid.err != nullptr) *kid.err = 1;
          const float* kp = kid.base[p] + (ok ? id : 0) * kid.stride[p];
#pragma unroll
          for (int x = 0; x < DQ / 2; x += 4) {
            const float4 v = ok ? __ldg(reinterpret_cast<const float4*>(kp + x)) : make_float4(0.f, 0.f, 0.f, 0.f);
            kk[p * (DQ / 2) + x] = v.x; kk[p * (DQ / 2) + x + 1] = v.y; kk[p * (DQ / 2) + x + 2] = v.z; kk[p * (DQ / 2) + x + 3] = v.w;
          }
        }
      } else if (t < n) {
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
                             const float* W3, const float* b3, float* out, float* scores, cudaStream_t st,
                             const DinKeyIds& kid) {
  const size_t smem = sizeof(DinTcSmem<DQ, H1, H2>) + 1024;
  PTREC_CUDA(cudaFuncSetAttribute(din_fwd_tc_kernel<DQ, H1, H2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(B < (int64_t)sms * 2 ? B : (int64_t)sms * 2);  // two CTAs per SM (~100 KB of shared memory each)
  din_fwd_tc_kernel<DQ, H1, H2><<<grid, kTcPos, smem, st>>>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3,
                                                             out, scores, kid);
  PTREC_LAUNCH_CHECK("din_fwd_tc_kernel");
  return PTREC_OK;
}

// ------------------------------------------------------------------------------------------------ backward
// Recomputes the unit (G1, G2 as in the forward) and runs its four gradient contractions on the tensor core as well.
// Every operand tile is stored ONCE, in the un-swizzled core-matrix layout, and read in both orientations: a tile
// stored [l][x] (x contiguous) is a K-major operand when the MMA reduces over x, and the SAME bytes are an MN-major
// operand (major bit of the instruction descriptor; LBO = 128 along x, SBO = the 8-row group stride along l) when the
// MMA reduces over the positions l — no transposed copies.
//   G3  dH1pre [128, H1] = dH2 [128, H2p] . W2              A = dH2 (K-major), B = W2 planes read MN-major
//   G4  [H1e, H2p]       = H1e^T . dH2    (over l)          A = H1e planes read MN-major (M = 128 covers the H1e rows),
//                                                           B = dH2 read MN-major;  H1e = [h1 | 1]: row H1 = grad b2
//   G5  [H1, DQe]        = dH1^T . Ke     (over l)          A = dH1 read MN-major, B = Ke = [k | 1] read MN-major:
//                                                           columns < DQ = grad M_b, column DQ = grad c_b
//   G6  dk [128, DQ]     = dH1 . M_b                        A = dH1 (K-major), B = M_b planes read MN-major
// The "ones" columns are stored unscaled (1.0 in the high plane), so their results carry only the other operand's
// scale.  Results leave TMEM per tile (the tiles have different scales) into fp32 accumulators in shared memory, one
// row per thread; a CTA owns kTcBwdGroup consecutive samples and writes one partial row (the layout of
// din_attn.cu's backward, reduced by its two-level kernel): short fixed-order chains, bit-reproducible.
constexpr int kTcBwdGroup = 8;

template <int DQ, int H1, int H2>
struct DinTcBwdSmem {
  static constexpr int H2P = (H2 + 15) / 16 * 16;
  static constexpr int DQE = DQ + 16;   // keys + the ones column (+ zero padding to the MMA granule)
  static constexpr int H1E = H1 + 16;   // h1 + the ones column
  alignas(1024) unsigned char k[2][kTcPos * DQE * 2];
  alignas(128) unsigned char m[2][H1 * DQ * 2];
  alignas(128) unsigned char h1[2][kTcPos * H1E * 2];   // h1 planes, later overwritten by the dH1 planes
  alignas(128) unsigned char w2[2][H2P * H1E * 2];
  alignas(128) unsigned char dh2[2][kTcPos * H2P * 2];
  alignas(16) float kf[kTcPos][DQ + 1];
  alignas(16) float accWkd[H1][DQ + 1];
  alignas(16) float accW1p[H1][DQ + 1];
  alignas(16) float accWq[H1][DQ + 1];
  alignas(16) float accW2T[H1][H2P + 1];
  alignas(16) float tmp[H1][DQ + 1];     // per-row contributions to grad q, summed over the rows in order
  alignas(16) float b1[H1];
  alignas(16) float b2[H2P];
  alignas(16) float W3[H2P];
  alignas(16) float c[H1];
  alignas(16) float q[DQ];
  alignas(16) float gp[DQ];
  float accb1[H1];
  float accb2[H2P];
  float accW3[H2P];
  float wpart[4][H2P + 1];               // per-warp partial column sums (grad W3)
  float bpart[4];
  float accb3;
  float mx[8];
  float b3;
  float s_w2;
  alignas(8) uint64_t bar[4];
  uint32_t tmem_slot;
};

// The same tile read MN-major (reduction over its ROWS): core matrices are the same 8 x 16-byte blocks; for the
// un-swizzled layout both majors name the strides alike — SBO between core matrices along M / N (here: the 128-byte
// step between column groups), LBO between core matrices along K (here: the 8-row group stride).  (Verified on B200:
// the swapped assignment produces garbage, this one matches the fp32 kernel to 3e-7; tools/diag_din_bwd_tc.py.)
__device__ __forceinline__ uint64_t make_nosw_mn_desc(const void* smem_tile, uint32_t group_stride) {
  return make_nosw_desc(smem_tile, group_stride, 128);
}

// three MMAs of one K = 16 step of a split product
__device__ __forceinline__ void umma3(uint32_t tD, uint32_t tC, uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1,
                                      uint32_t idesc, uint32_t acc) {
  umma_bf16(tC, a0, b1, idesc, acc);
  umma_bf16(tC, a1, b0, idesc, 1u);
  umma_bf16(tD, a0, b0, idesc, acc);
}

// Thread layout of the backward: 256 threads = two TEAMS of 128.  Thread (team, r) works on position r (TMEM lane r:
// warps w and w + 4 share a lane quadrant) and on the 16-column chunks of every epilogue whose index has the team's
// parity, so each epilogue's work is halved per thread and every scheduler has two warps to interleave (one 4-warp
// team left the tensor core and the LSU idle behind a single dependent instruction stream: 26 us per tile, IPC 0.13,
// profiles/r2_ncu_full_din_tc_raw.csv).  Row-level scalars (the tile maxima, a_l, <g_pooled, k_l>) are combined
// through shared memory in a fixed order.
constexpr int kTcBwdThreads = 2 * kTcPos;

__device__ __forceinline__ float block_max_256(float v, float* s_red) {  // 8 warps; s_red: 8 floats; all threads call
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float m = s_red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, s_red[i]);
  return m;
}

template <int DQ, int H1, int H2>
__global__ void __launch_bounds__(kTcBwdThreads, 1)
din_bwd_tc_kernel(const float* __restrict__ q, int64_t q_stride, const float* __restrict__ keys, int64_t ksb, int64_t ksl,
                  const int32_t* __restrict__ lens, int64_t B, int L, const float* __restrict__ W1,
                  const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                  const float* __restrict__ W3, const float* __restrict__ b3, const float* __restrict__ g_pooled,
                  float* __restrict__ g_q, float* __restrict__ g_keys, int64_t gksb, int64_t gksl,
                  float* __restrict__ partials, const DinKeyIds kid) {
  using S = DinTcBwdSmem<DQ, H1, H2>;
  constexpr int H2P = S::H2P, DQE = S::DQE, H1E = S::H1E;
  static_assert(DQ == 32 && H1 % 16 == 0 && H1E <= 128 && H1 + 1 <= kTcPos, "MMA shape / team split constraints");
  constexpr uint32_t kTmemCols = 256;
  static_assert(2 * H1 + 2 * H2P <= kTmemCols && 2 * DQE <= 2 * H1 && 2 * DQ <= 2 * H2P, "accumulators exceed TMEM");
  extern __shared__ unsigned char smem_raw[];
  S* s = reinterpret_cast<S*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int team = t >> 7, r = t & (kTcPos - 1);
  constexpr uint32_t kSboK = (DQE / 8) * 128, kSboM = (DQ / 8) * 128, kSboH = (H1E / 8) * 128, kSboW = (H1E / 8) * 128,
                     kSboG = (H2P / 8) * 128;

  // ---- once per CTA --------------------------------------------------------------------------------------------------
  for (int e = t; e < H1; e += kTcBwdThreads) {
    s->b1[e] = b1[e];
    s->accb1[e] = 0.f;
  }
  for (int e = t; e < H2P; e += kTcBwdThreads) {
    s->b2[e] = e < H2 ? b2[e] : 0.f;
    s->W3[e] = e < H2 ? W3[e] : 0.f;
    s->accb2[e] = 0.f;
    s->accW3[e] = 0.f;
  }
  for (int e = t; e < H1 * (DQ + 1); e += kTcBwdThreads) {
    (&s->accWkd[0][0])[e] = 0.f;
    (&s->accW1p[0][0])[e] = 0.f;
    (&s->accWq[0][0])[e] = 0.f;
  }
  for (int e = t; e < H1 * (H2P + 1); e += kTcBwdThreads) (&s->accW2T[0][0])[e] = 0.f;
  float wmax = 0.f;
  for (int e = t; e < H2 * H1; e += kTcBwdThreads) wmax = fmaxf(wmax, fabsf(W2[e]));
  wmax = block_max_256(wmax, s->mx);
  const float sw2 = h2_scale(wmax);
  if (t == 0) {
    s->b3 = b3[0];
    s->s_w2 = sw2;
    s->accb3 = 0.f;
    for (int i = 0; i < 4; ++i) mbar_init(&s->bar[i], 1);
    fence_mbar_init();
  }
  for (int ch = t; ch < H2P * (H1E / 8); ch += kTcBwdThreads) {  // W2 planes [H2P rows m][H1E columns j], zero padded
    const int mrow = ch / (H1E / 8), cj = ch - mrow * (H1E / 8);
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (mrow < H2 && cj * 8 + i < H1) ? W2[mrow * H1 + cj * 8 + i] * sw2 : 0.f;
    const uint32_t off = (uint32_t)(mrow >> 3) * kSboW + cj * 128 + (mrow & 7) * 16;
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
  const uint32_t tX = tmem, tY = tmem + 2 * H1;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;  // both teams' warps of a quadrant read the same lanes
  constexpr uint32_t kM = (uint32_t)(kTcPos >> 4) << 24, kF = 1u << 4, kAmn = 1u << 15, kBmn = 1u << 16;
  constexpr uint32_t kI1 = kF | kM | ((uint32_t)(H1 >> 3) << 17);
  constexpr uint32_t kI2 = kF | kM | ((uint32_t)(H2P >> 3) << 17);
  constexpr uint32_t kI3 = kF | kM | ((uint32_t)(H1 >> 3) << 17) | kBmn;
  constexpr uint32_t kI4 = kF | kM | ((uint32_t)(H2P >> 3) << 17) | kAmn | kBmn;
  constexpr uint32_t kI5 = kF | kM | ((uint32_t)(DQE >> 3) << 17) | kAmn | kBmn;
  constexpr uint32_t kI6 = kF | kM | ((uint32_t)(DQ >> 3) << 17) | kBmn;
  uint32_t phase = 0;
  const __half one = __float2half_rn(1.f);
  float* s_pair = &s->kf[0][0];  // [2][kTcPos] per-team row partials (kf is not otherwise used by the backward)

  const int64_t b_end = min(B, ((int64_t)blockIdx.x + 1) * kTcBwdGroup);
  for (int64_t b = (int64_t)blockIdx.x * kTcBwdGroup; b < b_end; ++b) {
    const int len = lens ? min(max(lens[b], 0), L) : L;
    __syncthreads();
    if (t < DQ) {
      s->q[t] = q[b * q_stride + t];
      s->gp[t] = g_pooled[b * DQ + t];
    }
    __syncthreads();
    // ---- M_b planes and c_b ------------------------------------------------------------------------------------------
    constexpr int kMCh = (H1 * DQ / 8 + kTcBwdThreads - 1) / kTcBwdThreads;
    float mv[kMCh][8];
    float mmax = 0.f;
#pragma unroll
    for (int rr = 0; rr < kMCh; ++rr) {
      const int ch = t + rr * kTcBwdThreads;
      if (ch < H1 * DQ / 8) {
        const int j = ch / (DQ / 8), ci = ch - j * (DQ / 8);
        const float* row = W1 + (int64_t)j * 4 * DQ + ci * 8;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 wk = __ldg(reinterpret_cast<const float4*>(row + DQ + 4 * h));
          const float4 wd = __ldg(reinterpret_cast<const float4*>(row + 2 * DQ + 4 * h));
          const float4 wp = __ldg(reinterpret_cast<const float4*>(row + 3 * DQ + 4 * h));
          const float* qq = &s->q[ci * 8 + 4 * h];
          mv[rr][4 * h + 0] = (wk.x - wd.x) + wp.x * qq[0];
          mv[rr][4 * h + 1] = (wk.y - wd.y) + wp.y * qq[1];
          mv[rr][4 * h + 2] = (wk.z - wd.z) + wp.z * qq[2];
          mv[rr][4 * h + 3] = (wk.w - wd.w) + wp.w * qq[3];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) mmax = fmaxf(mmax, fabsf(mv[rr][i]));
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
    mmax = block_max_256(mmax, s->mx);
    const float sm = h2_scale(mmax);
#pragma unroll
    for (int rr = 0; rr < kMCh; ++rr) {
      const int ch = t + rr * kTcBwdThreads;
      if (ch < H1 * DQ / 8) {
        const int j = ch / (DQ / 8), ci = ch - j * (DQ / 8);
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = mv[rr][i] * sm;
        const uint32_t off = (uint32_t)(j >> 3) * kSboM + ci * 128 + (j & 7) * 16;
        h2_split8(x, reinterpret_cast<uint4*>(s->m[0] + off), reinterpret_cast<uint4*>(s->m[1] + off));
      }
    }
    float gq_acc = 0.f;  // thread i < DQ: grad q[i], summed over the tiles
    for (int l0 = 0; l0 < max(len, 1); l0 += kTcPos) {
      const int n = max(0, min(kTcPos, len - l0));
      const bool live = r < n;
      // ---- Ke planes: team 0 formats columns [0, 16), team 1 columns [16, 32) and the ones / zero chunks ---------------
      constexpr int kHalf = DQ / 2;
      float kk[kHalf];
      float kmax = 0.f, gpart = 0.f;
      if (live && kid.ids[0] != nullptr) {  // gather fused in: team p reads its half from table p
        const int64_t id = kid.ids[team][b * kid.ids_sb + kid.ids_off + l0 + r];
        const bool ok = (uint64_t)id < (uint64_t)kid.rows[team];
        if (!ok && kid.err != nullptr) *kid.err = 1;
        const float* kp = kid.base[team] + (ok ? id : 0) * kid.stride[team];
#pragma unroll
        for (int x = 0; x < kHalf; x += 4) {
          const float4 v = ok ? __ldg(reinterpret_cast<const float4*>(kp + x)) : make_float4(0.f, 0.f, 0.f, 0.f);
          kk[x] = v.x; kk[x + 1] = v.y; kk[x + 2] = v.z; kk[x + 3] = v.w;
        }
      } else if (live) {
        const float* kp = keys + b * ksb + (int64_t)(l0 + r) * ksl + team * kHalf;
#pragma unroll
        for (int x = 0; x < kHalf; x += 4) {
          const float4 v = ldg_stream_f4(kp + x);
          kk[x] = v.x; kk[x + 1] = v.y; kk[x + 2] = v.z; kk[x + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int x = 0; x < kHalf; ++x) kk[x] = 0.f;
      }
#pragma unroll
      for (int x = 0; x < kHalf; ++x) {
        kmax = fmaxf(kmax, fabsf(kk[x]));
        gpart += s->gp[team * kHalf + x] * kk[x];
      }
      s_pair[team * kTcPos + r] = gpart;
      kmax = block_max_256(kmax, s->mx);  // (its barriers also publish s_pair)
      const float ga_raw = s_pair[r] + s_pair[kTcPos + r];  // d loss / d a_l = <g_pooled, k_l>, same order in both teams
      const float sk = h2_scale(kmax);
      {
        const uint32_t rowoff = (uint32_t)(r >> 3) * kSboK + (r & 7) * 16;
#pragma unroll
        for (int ci = 0; ci < kHalf / 8; ++ci) {
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = kk[ci * 8 + i] * sk;
          const uint32_t o = rowoff + (team * (kHalf / 8) + ci) * 128;
          h2_split8(x, reinterpret_cast<uint4*>(s->k[0] + o), reinterpret_cast<uint4*>(s->k[1] + o));
        }
        if (team == 1) {  // columns DQ .. DQE: the ones column (unscaled, live positions only), then zeros
          __half ext[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) ext[i] = __float2half_rn(0.f);
          const uint4 zero = *reinterpret_cast<const uint4*>(ext);
          ext[0] = live ? one : ext[0];
          *reinterpret_cast<uint4*>(s->k[0] + rowoff + (DQ / 8) * 128) = *reinterpret_cast<const uint4*>(ext);
          *reinterpret_cast<uint4*>(s->k[1] + rowoff + (DQ / 8) * 128) = zero;
          *reinterpret_cast<uint4*>(s->k[0] + rowoff + (DQ / 8 + 1) * 128) = zero;
          *reinterpret_cast<uint4*>(s->k[1] + rowoff + (DQ / 8 + 1) * 128) = zero;
        }
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncthreads();
      if (t == 0) {  // G1
        tcgen05_fence_after();
#pragma unroll
        for (int ks = 0; ks < DQ / 16; ++ks)
          umma3(tX, tX + H1, make_nosw_desc(s->k[0] + ks * 256, 128, kSboK), make_nosw_desc(s->k[1] + ks * 256, 128, kSboK),
                make_nosw_desc(s->m[0] + ks * 256, 128, kSboM), make_nosw_desc(s->m[1] + ks * 256, 128, kSboM), kI1,
                ks > 0 ? 1u : 0u);
        umma_commit(&s->bar[0]);
      }
      mbar_wait(&s->bar[0], phase);
      tcgen05_fence_after();
      // ---- epilogue 1: h1 -> H1e planes (this team's chunks); ReLU mask bits of those chunks ---------------------------
      const float inv1 = 1.f / (sk * sm);
      float hmax = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < H1; c0 += 16) {
        if (((c0 >> 4) & 1) != team) continue;
        float dm[16], dc[16];
        tmem_ld16x2(tX + lane_base + c0, tX + H1 + lane_base + c0, dm, dc);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          hmax = fmaxf(hmax, live ? (dm[i] + dc[i] * (1.f / 2048.f)) * inv1 + s->c[c0 + i] : 0.f);
      }
      hmax = block_max_256(hmax, s->mx);
      const float sh = h2_scale(hmax);
      uint32_t mask1[(H1 + 31) / 32];
#pragma unroll
      for (int w = 0; w < (H1 + 31) / 32; ++w) mask1[w] = 0u;
      {
        const uint32_t rowoff = (uint32_t)(r >> 3) * kSboH + (r & 7) * 16;
#pragma unroll
        for (int c0 = 0; c0 < H1; c0 += 16) {
          if (((c0 >> 4) & 1) != team) continue;
          float dm[16], dc[16], x[16];
          tmem_ld16x2(tX + lane_base + c0, tX + H1 + lane_base + c0, dm, dc);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float h = (dm[i] + dc[i] * (1.f / 2048.f)) * inv1 + s->c[c0 + i];
            const bool on = live && h > 0.f;
            if (on) mask1[(c0 + i) >> 5] |= 1u << ((c0 + i) & 31);
            x[i] = on ? h * sh : 0.f;
          }
          h2_split8(x, reinterpret_cast<uint4*>(s->h1[0] + rowoff + (c0 / 8) * 128), reinterpret_cast<uint4*>(s->h1[1] + rowoff + (c0 / 8) * 128));
          h2_split8(x + 8, reinterpret_cast<uint4*>(s->h1[0] + rowoff + (c0 / 8 + 1) * 128),
                    reinterpret_cast<uint4*>(s->h1[1] + rowoff + (c0 / 8 + 1) * 128));
        }
        if (team == 1) {
          __half ext[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) ext[i] = __float2half_rn(0.f);
          const uint4 zero = *reinterpret_cast<const uint4*>(ext);
          ext[0] = live ? one : ext[0];
          *reinterpret_cast<uint4*>(s->h1[0] + rowoff + (H1 / 8) * 128) = *reinterpret_cast<const uint4*>(ext);
          *reinterpret_cast<uint4*>(s->h1[1] + rowoff + (H1 / 8) * 128) = zero;
          *reinterpret_cast<uint4*>(s->h1[0] + rowoff + (H1 / 8 + 1) * 128) = zero;
          *reinterpret_cast<uint4*>(s->h1[1] + rowoff + (H1 / 8 + 1) * 128) = zero;
        }
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncthreads();
      if (t == 0) {  // G2
        tcgen05_fence_after();
#pragma unroll
        for (int ks = 0; ks < H1 / 16; ++ks)
          umma3(tY, tY + H2P, make_nosw_desc(s->h1[0] + ks * 256, 128, kSboH), make_nosw_desc(s->h1[1] + ks * 256, 128, kSboH),
                make_nosw_desc(s->w2[0] + ks * 256, 128, kSboW), make_nosw_desc(s->w2[1] + ks * 256, 128, kSboW), kI2,
                ks > 0 ? 1u : 0u);
        umma_commit(&s->bar[1]);
      }
      mbar_wait(&s->bar[1], phase);
      tcgen05_fence_after();
      // ---- epilogue 2: a_l (two partial sums, team 0 first), dH2 planes, grad W3 / b3 ----------------------------------
      const float inv2 = 1.f / (sh * s->s_w2);
      const float ga = live ? ga_raw : 0.f;
      float apart = 0.f;
      float gmax = 0.f;
      constexpr int kC2 = (H2P / 16 + 1) / 2;  // chunks of H2P a team can own
      float dh2[kC2][16];
      float gw3[kC2][16];
#pragma unroll
      for (int cc = 0; cc < kC2; ++cc) {
        const int c0 = (2 * cc + team) * 16;
        if (c0 < H2P) {
          float dm[16], dc[16];
          tmem_ld16x2(tY + lane_base + c0, tY + H2P + lane_base + c0, dm, dc);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float pre = (dm[i] + dc[i] * (1.f / 2048.f)) * inv2 + s->b2[c0 + i];
            const float rl = fmaxf(pre, 0.f);
            apart += s->W3[c0 + i] * rl;
            dh2[cc][i] = (live && pre > 0.f) ? ga * s->W3[c0 + i] : 0.f;
            gw3[cc][i] = ga * rl;
            gmax = fmaxf(gmax, fabsf(dh2[cc][i]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) dh2[cc][i] = gw3[cc][i] = 0.f;
        }
      }
      s_pair[team * kTcPos + r] = apart;
      gmax = block_max_256(gmax, s->mx);  // (its barriers also publish s_pair)
      const float a = live ? s->b3 + s_pair[r] + s_pair[kTcPos + r] : 0.f;
      const float sg2 = h2_scale(gmax);
      {
        const uint32_t rowoff = (uint32_t)(r >> 3) * kSboG + (r & 7) * 16;
#pragma unroll
        for (int cc = 0; cc < kC2; ++cc) {
          const int c0 = (2 * cc + team) * 16;
          if (c0 < H2P) {
            float x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = dh2[cc][i] * sg2;
            h2_split8(x, reinterpret_cast<uint4*>(s->dh2[0] + rowoff + (c0 / 8) * 128), reinterpret_cast<uint4*>(s->dh2[1] + rowoff + (c0 / 8) * 128));
            h2_split8(x + 8, reinterpret_cast<uint4*>(s->dh2[0] + rowoff + (c0 / 8 + 1) * 128),
                      reinterpret_cast<uint4*>(s->dh2[1] + rowoff + (c0 / 8 + 1) * 128));
          }
        }
      }
      // column sums over the positions: warp tree, then this team's four warps in order (by the column's owner below)
#pragma unroll
      for (int cc = 0; cc < kC2; ++cc) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) gw3[cc][i] += __shfl_xor_sync(0xffffffffu, gw3[cc][i], o);
        }
      }
      float gb3 = team == 0 ? ga : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) gb3 += __shfl_xor_sync(0xffffffffu, gb3, o);
      if (lane == 0) {
#pragma unroll
        for (int cc = 0; cc < kC2; ++cc) {
          const int c0 = (2 * cc + team) * 16;
          if (c0 < H2P) {
#pragma unroll
            for (int i = 0; i < 16; ++i) s->wpart[warp & 3][c0 + i] = gw3[cc][i];
          }
        }
        if (team == 0) s->bpart[warp] = gb3;
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncthreads();
      if (t == 0) {  // G3 into region X, G4 into region Y
        tcgen05_fence_after();
#pragma unroll
        for (int ks = 0; ks < H2P / 16; ++ks)
          umma3(tX, tX + H1, make_nosw_desc(s->dh2[0] + ks * 256, 128, kSboG), make_nosw_desc(s->dh2[1] + ks * 256, 128, kSboG),
                make_nosw_mn_desc(s->w2[0] + ks * 2 * kSboW, kSboW), make_nosw_mn_desc(s->w2[1] + ks * 2 * kSboW, kSboW),
                kI3, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < kTcPos / 16; ++ks)
          umma3(tY, tY + H2P, make_nosw_mn_desc(s->h1[0] + ks * 2 * kSboH, kSboH),
                make_nosw_mn_desc(s->h1[1] + ks * 2 * kSboH, kSboH), make_nosw_mn_desc(s->dh2[0] + ks * 2 * kSboG, kSboG),
                make_nosw_mn_desc(s->dh2[1] + ks * 2 * kSboG, kSboG), kI4, ks > 0 ? 1u : 0u);
        umma_commit(&s->bar[2]);
      }
      if (t < H2) {  // grad W3 / b3 of this tile (each column's four warp partials in order); overlaps the MMAs
        s->accW3[t] += ((s->wpart[0][t] + s->wpart[1][t]) + s->wpart[2][t]) + s->wpart[3][t];
      } else if (t == H2P) {
        s->accb3 += ((s->bpart[0] + s->bpart[1]) + s->bpart[2]) + s->bpart[3];
      }
      mbar_wait(&s->bar[2], phase);
      tcgen05_fence_after();
      // ---- epilogue 3: dH1 planes (into the H1e buffer: G4 is done with it); grad W2 / b2 rows ---------------------------
      const float inv3 = 1.f / (sg2 * s->s_w2);
      float g1max = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < H1; c0 += 16) {
        if (((c0 >> 4) & 1) != team) continue;
        float dm[16], dc[16];
        tmem_ld16x2(tX + lane_base + c0, tX + H1 + lane_base + c0, dm, dc);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const bool on = (mask1[(c0 + i) >> 5] >> ((c0 + i) & 31)) & 1u;
          g1max = fmaxf(g1max, on ? fabsf((dm[i] + dc[i] * (1.f / 2048.f)) * inv3) : 0.f);
        }
      }
      {  // row r of G4: r < H1 -> grad W2[:, r] of this tile, r == H1 -> grad b2; this team's chunks of the columns
        const float inv4 = r < H1 ? 1.f / (sh * sg2) : 1.f / sg2;
#pragma unroll
        for (int c0 = 0; c0 < H2P; c0 += 16) {
          if (((c0 >> 4) & 1) != team) continue;
          float dm[16], dc[16];
          tmem_ld16x2(tY + lane_base + c0, tY + H2P + lane_base + c0, dm, dc);
          if (r < H1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) s->accW2T[r][c0 + i] += (dm[i] + dc[i] * (1.f / 2048.f)) * inv4;
          } else if (r == H1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) s->accb2[c0 + i] += (dm[i] + dc[i] * (1.f / 2048.f)) * inv4;
          }
        }
      }
      g1max = block_max_256(g1max, s->mx);
      const float sg1 = h2_scale(g1max);
      {
        const uint32_t rowoff = (uint32_t)(r >> 3) * kSboH + (r & 7) * 16;
#pragma unroll
        for (int c0 = 0; c0 < H1; c0 += 16) {
          if (((c0 >> 4) & 1) != team) continue;
          float dm[16], dc[16], x[16];
          tmem_ld16x2(tX + lane_base + c0, tX + H1 + lane_base + c0, dm, dc);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const bool on = (mask1[(c0 + i) >> 5] >> ((c0 + i) & 31)) & 1u;
            x[i] = on ? (dm[i] + dc[i] * (1.f / 2048.f)) * inv3 * sg1 : 0.f;
          }
          h2_split8(x, reinterpret_cast<uint4*>(s->h1[0] + rowoff + (c0 / 8) * 128), reinterpret_cast<uint4*>(s->h1[1] + rowoff + (c0 / 8) * 128));
          h2_split8(x + 8, reinterpret_cast<uint4*>(s->h1[0] + rowoff + (c0 / 8 + 1) * 128),
                    reinterpret_cast<uint4*>(s->h1[1] + rowoff + (c0 / 8 + 1) * 128));
        }
        if (team == 1) {  // the ones column of H1e becomes a zero column of the dH1 planes
          __half ext[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) ext[i] = __float2half_rn(0.f);
          *reinterpret_cast<uint4*>(s->h1[0] + rowoff + (H1 / 8) * 128) = *reinterpret_cast<const uint4*>(ext);
        }
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncthreads();
      if (t == 0) {  // G5 into region X, G6 into region Y
        tcgen05_fence_after();
#pragma unroll
        for (int ks = 0; ks < kTcPos / 16; ++ks)
          umma3(tX, tX + DQE, make_nosw_mn_desc(s->h1[0] + ks * 2 * kSboH, kSboH),
                make_nosw_mn_desc(s->h1[1] + ks * 2 * kSboH, kSboH), make_nosw_mn_desc(s->k[0] + ks * 2 * kSboK, kSboK),
                make_nosw_mn_desc(s->k[1] + ks * 2 * kSboK, kSboK), kI5, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < H1 / 16; ++ks)
          umma3(tY, tY + DQ, make_nosw_desc(s->h1[0] + ks * 256, 128, kSboH), make_nosw_desc(s->h1[1] + ks * 256, 128, kSboH),
                make_nosw_mn_desc(s->m[0] + ks * 2 * kSboM, kSboM), make_nosw_mn_desc(s->m[1] + ks * 2 * kSboM, kSboM),
                kI6, ks > 0 ? 1u : 0u);
        umma_commit(&s->bar[3]);
      }
      mbar_wait(&s->bar[3], phase);
      tcgen05_fence_after();
      phase ^= 1u;
      // ---- epilogue 4: grad keys (this team's 16 columns); row r of G5 -> grads of M_b / c_b -> W1 accumulators, grad q -
      {
        const float inv6 = 1.f / (sg1 * sm);
        const int c0 = team * 16;
        float dm[16], dc[16], gk[16];
        tmem_ld16x2(tY + lane_base + c0, tY + DQ + lane_base + c0, dm, dc);
#pragma unroll
        for (int i = 0; i < 16; ++i) gk[i] = (dm[i] + dc[i] * (1.f / 2048.f)) * inv6 + a * s->gp[c0 + i];
        if (live) {
          float* gkp = g_keys + b * gksb + (int64_t)(l0 + r) * gksl + c0;
#pragma unroll
          for (int x = 0; x < 16; x += 4) st_f4(gkp + x, make_float4(gk[x], gk[x + 1], gk[x + 2], gk[x + 3]));
        }
      }
      {
        const float inv5 = 1.f / (sg1 * sk);
        const int c0 = team * 16;
        float dm[16], dc[16], em[16], ec[16];
        tmem_ld16x2(tX + lane_base + c0, tX + DQE + lane_base + c0, dm, dc);
        tmem_ld16x2(tX + lane_base + DQ, tX + DQE + lane_base + DQ, em, ec);  // the ones column: only dH1's scale
        const float dcj = (em[0] + ec[0] * (1.f / 2048.f)) / sg1;
        if (r < H1) {
          const float* row = W1 + (int64_t)r * 4 * DQ + c0;
          if (team == 0) s->accb1[r] += dcj;
#pragma unroll
          for (int i4 = 0; i4 < 16; i4 += 4) {
            const float4 wq = __ldg(reinterpret_cast<const float4*>(row + i4));
            const float4 wd = __ldg(reinterpret_cast<const float4*>(row + 2 * DQ + i4));
            const float4 wp = __ldg(reinterpret_cast<const float4*>(row + 3 * DQ + i4));
            const float wqs[4] = {wq.x + wd.x, wq.y + wd.y, wq.z + wd.z, wq.w + wd.w};
            const float wps[4] = {wp.x, wp.y, wp.z, wp.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i = c0 + i4 + u;
              const float dMv = (dm[i4 + u] + dc[i4 + u] * (1.f / 2048.f)) * inv5;
              const float qv = s->q[i];
              s->accWkd[r][i] += dMv;
              s->accW1p[r][i] += dMv * qv;
              s->accWq[r][i] += dcj * qv;
              s->tmp[r][i] = wqs[u] * dcj + wps[u] * dMv;  // d c / d q = W1q + W1d ; d M / d q = W1p (column-wise)
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncthreads();
      if (t < DQ) {
        float rs = 0.f;
        for (int j = 0; j < H1; ++j) rs += s->tmp[j][t];
        gq_acc += rs;
      }
    }
    // zero gradient for the padded tail of the history; grad q
    for (int64_t e = len * (int64_t)DQ + t; e < (int64_t)L * DQ; e += kTcBwdThreads) {
      const int64_t l = e / DQ;
      g_keys[b * gksb + l * gksl + (e - l * DQ)] = 0.f;
    }
    if (t < DQ) g_q[b * DQ + t] = gq_acc;
  }
  __syncthreads();
  // ---- per-CTA partial weight gradients: [W1 (H1 x 4DQ) | b1 | W2 (H2 x H1) | b2 | W3 | b3] (din_attn.cu's layout) ----------
  float* P = partials + (int64_t)blockIdx.x * (H1 * 4 * DQ + H1 + H2 * H1 + H2 + H2 + 1);
  for (int e = t; e < H1 * DQ; e += kTcBwdThreads) {
    const int j = e / DQ, i = e - j * DQ;
    float* row = P + (int64_t)j * 4 * DQ;
    const float a3 = s->accWq[j][i], a1 = s->accWkd[j][i];
    row[i] = a3;                 // W1q
    row[DQ + i] = a1;            // W1k
    row[2 * DQ + i] = a3 - a1;   // W1d  (Wq = W1q + W1d, Wkd = W1k - W1d)
    row[3 * DQ + i] = s->accW1p[j][i];
  }
  for (int e = t; e < H1; e += kTcBwdThreads) P[H1 * 4 * DQ + e] = s->accb1[e];
  float* PW2 = P + H1 * 4 * DQ + H1;
  for (int e = t; e < H2 * H1; e += kTcBwdThreads) {
    const int m = e / H1, j = e - m * H1;
    PW2[e] = s->accW2T[j][m];
  }
  for (int e = t; e < H2; e += kTcBwdThreads) {
    PW2[H2 * H1 + e] = s->accb2[e];
    PW2[H2 * H1 + H2 + e] = s->accW3[e];
  }
  if (t == 0) PW2[H2 * H1 + 2 * H2] = s->accb3;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

template <int DQ, int H1, int H2>
static int din_bwd_tc_launch(const float* q, int64_t qs, const float* keys, int64_t ksb, int64_t ksl, const int32_t* lens,
                             int64_t B, int L, const float* W1, const float* b1, const float* W2, const float* b2,
                             const float* W3, const float* b3, const float* g_pooled, float* g_q, float* g_keys,
                             int64_t gksb, int64_t gksl, float* partials, int* n_rows, cudaStream_t st,
                             const DinKeyIds& kid) {
  const size_t smem = sizeof(DinTcBwdSmem<DQ, H1, H2>) + 1024;
  PTREC_CUDA(cudaFuncSetAttribute(din_bwd_tc_kernel<DQ, H1, H2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)((B + kTcBwdGroup - 1) / kTcBwdGroup);
  din_bwd_tc_kernel<DQ, H1, H2><<<grid, kTcBwdThreads, smem, st>>>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3,
                                                             g_pooled, g_q, g_keys, gksb, gksl, partials, kid);
  PTREC_LAUNCH_CHECK("din_bwd_tc_kernel");
  *n_rows = grid;
  return PTREC_OK;
}

// backward entry: fills g_q, g_keys and `*n_rows` partial rows of weight gradients (the caller reduces them)
int din_bwd_tc(const float* q, int64_t qs, const float* keys, int64_t ksb, int64_t ksl, const int32_t* lens, int64_t B,
               int L, int DQ, int H1, int H2, const float* W1, const float* b1, const float* W2, const float* b2,
               const float* W3, const float* b3, const float* g_pooled, float* g_q, float* g_keys, int64_t gksb,
               int64_t gksl, float* partials, int* n_rows, cudaStream_t st, const DinKeyIds* kid_in) {
  DinKeyIds kid;
  memset(&kid, 0, sizeof(kid));
  if (kid_in != nullptr) kid = *kid_in;
  if (DQ == 32 && H1 == 80 && H2 == 40)
    return din_bwd_tc_launch<32, 80, 40>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3, g_pooled, g_q, g_keys,
                                         gksb, gksl, partials, n_rows, st, kid);
  if (DQ == 32 && H1 == 64 && H2 == 32)
    return din_bwd_tc_launch<32, 64, 32>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3, g_pooled, g_q, g_keys,
                                         gksb, gksl, partials, n_rows, st, kid);
  return PTREC_EUNSUPPORTED;
}

// entry used by din_attn.cu's dispatcher: PTREC_OK, or PTREC_EUNSUPPORTED when this shape has no tensor-core build
int din_fwd_tc(const float* q, int64_t qs, const float* keys, int64_t ksb, int64_t ksl, const int32_t* lens, int64_t B,
               int L, int DQ, int H1, int H2, const float* W1, const float* b1, const float* W2, const float* b2,
               const float* W3, const float* b3, float* out, float* scores, cudaStream_t st, const DinKeyIds* kid_in) {
  DinKeyIds kid;
  memset(&kid, 0, sizeof(kid));
  if (kid_in != nullptr) kid = *kid_in;
  if (DQ == 32 && H1 == 80 && H2 == 40)
    return din_fwd_tc_launch<32, 80, 40>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3, out, scores, st, kid);
  if (DQ == 32 && H1 == 64 && H2 == 32)
    return din_fwd_tc_launch<32, 64, 32>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3, out, scores, st, kid);
  if (DQ == 16 && H1 == 80 && H2 == 40)
    return din_fwd_tc_launch<16, 80, 40>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3, out, scores, st, kid);
  if (DQ == 16 && H1 == 64 && H2 == 32)
    return din_fwd_tc_launch<16, 64, 32>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3, out, scores, st, kid);
  return PTREC_EUNSUPPORTED;
}

}  // namespace ptrec
