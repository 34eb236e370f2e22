// Exact 3-way bf16 split of an fp32 value: v == p0 + p1 + p2 (3 x 8 = 24 mantissa bits).  Shared by K6 and K8.
#pragma once
#include <cuda_bf16.h>

namespace ptrec {

__device__ __forceinline__ void split3(float v, __nv_bfloat16& p0, __nv_bfloat16& p1, __nv_bfloat16& p2) {
  p0 = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(p0);
  p1 = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(p1);
  p2 = __float2bfloat16_rn(r2);
}

// 4 consecutive values -> three 8-byte groups written at o, o + plane, o + 2*plane
__device__ __forceinline__ void split3_store4(const float* v, __nv_bfloat16* o, int64_t plane) {
  __nv_bfloat16 p0[4], p1[4], p2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split3(v[i], p0[i], p1[i], p2[i]);
  *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(p0);
  *reinterpret_cast<uint2*>(o + plane) = *reinterpret_cast<const uint2*>(p1);
  *reinterpret_cast<uint2*>(o + 2 * plane) = *reinterpret_cast<const uint2*>(p2);
}

}  // namespace ptrec

// ---- fp16 x 2 ("h2") split ---------------------------------------------------------------------------------------
// x * s == h0 + h1 / 2048 up to 2^-23 relative, where s is a power of two that puts the tensor's largest magnitude in
// [2^13, 2^14) (fp16 has 5 exponent bits: without s, gradients of ~1e-8 would fall below its subnormals).  The second
// plane is stored multiplied by 2^11 so that it is a normal fp16 number whenever the first one is (Ootomo & Yokota's
// error-corrected tensor-core GEMM); the GEMM accumulates the two cross terms in their own accumulator and applies
// 2^-11 and 1 / (s_a s_b) in its fp32 epilogue.  2 x 11 = 22 mantissa bits, 3 MMAs per product.
#include <cuda_fp16.h>

namespace ptrec {

constexpr float kH2Second = 2048.f;          // 2^11
constexpr float kH2SecondInv = 1.f / 2048.f;

// power-of-two scale for a tensor whose largest magnitude is m: m * s in [2^13, 2^14) (exponent clamped to the normal
// fp32 range; the epilogue undoes the two operand scales with two separate exact multiplications); 0 / inf / nan -> 1.
__device__ __forceinline__ float h2_scale(float m) {
  if (!(m > 0.f) || m > 3.0e38f) return 1.f;
  const int e = (int)((__float_as_uint(m) >> 23) & 0xFFu) - 127;  // floor(log2 m) (-127 for fp32 subnormals)
  int se = 13 - e;
  se = se < -126 ? -126 : (se > 126 ? 126 : se);
  return __uint_as_float((uint32_t)(se + 127) << 23);
}

__device__ __forceinline__ void split2h(float v, float scale, __half& h0, __half& h1) {
  const float xs = v * scale;  // exact (power of two)
  h0 = __float2half_rn(xs);
  const float r = xs - __half2float(h0);  // exact in fp32
  h1 = __float2half_rn(r * kH2Second);
}

// two values at once through the packed conversions (F2FP.F16.F32.PACK_AB / HADD2.F32): the single-value F2F runs on a
// slow pipe and was 256 instructions per thread and tile in the fused GEMM epilogue.  Same roundings, same bits.
// p0 = {h0(v0) low half, h0(v1) high half}, p1 likewise for the second plane.
__device__ __forceinline__ void split2h_pair(float v0, float v1, float scale, uint32_t& p0, uint32_t& p1) {
  const float x0 = v0 * scale, x1 = v1 * scale;  // exact (power of two)
  const __half2 h0 = __floats2half2_rn(x0, x1);
  const float2 f0 = __half22float2(h0);
  const __half2 h1 = __floats2half2_rn((x0 - f0.x) * kH2Second, (x1 - f0.y) * kH2Second);
  p0 = *reinterpret_cast<const uint32_t*>(&h0);
  p1 = *reinterpret_cast<const uint32_t*>(&h1);
}

}  // namespace ptrec
