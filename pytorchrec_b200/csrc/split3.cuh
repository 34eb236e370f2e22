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
