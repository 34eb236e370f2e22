// K3: FM second-order interaction, forward and backward.
//   y[b] = 0.5 * sum_k ( (sum_f v[b,f,k])^2 - sum_f v[b,f,k]^2 )
//   dv[b,f,k] = gy[b] * (S[b,k] - v[b,f,k]) (+ upstream gradient of the other consumer of v)
// Bound: HBM.  Algorithmic bytes: fwd B*F*D*4 + B*4; bwd 2*B*F*D*4 (+ B*F*D*4 when grad_in is fused).
// Warp per sample; lane i owns the float4 chunks i, i+32, ... of the sample's F*D vector.  Because
// D/4 divides 32, a lane always sees the same k-slice, so S is a register accumulator and one xor
// tree finishes it.  The backward keeps v in registers between the S pass and the dv pass.
#include "common.cuh"

namespace ptrec {

constexpr int kFmWarps = 8;
constexpr int kFmMaxChunks = 8;  // register-cached float4 chunks per lane (F*D <= 1024)

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// reduce S over lanes sharing a k-slice (lane bits >= log2(D/4))
__device__ __forceinline__ float4 reduce_S(float4 s, int lanes_per_row) {
  for (int o = lanes_per_row; o < 32; o <<= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
    s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, o);
    s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
  }
  return s;
}

__global__ void __launch_bounds__(kFmWarps * 32)
fm2_fwd_kernel(const float* __restrict__ v, int64_t stride, int64_t B, int n_chunks, int lanes_per_row,
               float* __restrict__ y) {
  const int64_t b = (int64_t)blockIdx.x * kFmWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const float* row = v + b * stride;
  float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
  float Q = 0.f;
  for (int q0 = lane; q0 < n_chunks; q0 += 32 * 4) {
    float4 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * 32;
      t[u] = q < n_chunks ? ldg_stream_f4(row + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      S = f4_add(S, t[u]);
      Q += t[u].x * t[u].x + t[u].y * t[u].y + t[u].z * t[u].z + t[u].w * t[u].w;
    }
  }
  S = reduce_S(S, lanes_per_row);
  float r = (lane < lanes_per_row) ? (S.x * S.x + S.y * S.y + S.z * S.z + S.w * S.w) : 0.f;
  r -= Q;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  if (lane == 0) y[b] = 0.5f * r;
}

__global__ void __launch_bounds__(kFmWarps * 32)
fm2_bwd_kernel(const float* __restrict__ v, int64_t stride, const float* __restrict__ gy,
               const float* __restrict__ grad_in, int64_t gi_stride, int64_t B, int n_chunks,
               int lanes_per_row, float* __restrict__ grad_v, int64_t gv_stride) {
  const int64_t b = (int64_t)blockIdx.x * kFmWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const float* row = v + b * stride;
  float4 c[kFmMaxChunks];
  float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int u = 0; u < kFmMaxChunks; ++u) {
    const int q = lane + u * 32;
    c[u] = q < n_chunks ? ldg_stream_f4(row + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < kFmMaxChunks; ++u) S = f4_add(S, c[u]);
  for (int q = lane + kFmMaxChunks * 32; q < n_chunks; q += 32) S = f4_add(S, ldg_stream_f4(row + q * 4));
  S = reduce_S(S, lanes_per_row);
  const float g = gy[b];
  float* orow = grad_v + b * gv_stride;
  const float* irow = grad_in ? grad_in + b * gi_stride : nullptr;
#pragma unroll
  for (int u = 0; u < kFmMaxChunks; ++u) {
    const int q = lane + u * 32;
    if (q < n_chunks) {
      float4 o = make_float4(g * (S.x - c[u].x), g * (S.y - c[u].y), g * (S.z - c[u].z), g * (S.w - c[u].w));
      if (irow) o = f4_add(o, ldg_stream_f4(irow + q * 4));
      st_f4(orow + q * 4, o);
    }
  }
  for (int q = lane + kFmMaxChunks * 32; q < n_chunks; q += 32) {
    const float4 t = ldg_stream_f4(row + q * 4);
    float4 o = make_float4(g * (S.x - t.x), g * (S.y - t.y), g * (S.z - t.z), g * (S.w - t.w));
    if (irow) o = f4_add(o, ldg_stream_f4(irow + q * 4));
    st_f4(orow + q * 4, o);
  }
}

// generic fallback (D not a power-of-two multiple of 4, or unaligned views): thread per (sample, k)
__global__ void fm2_fwd_generic_kernel(const float* __restrict__ v, int64_t stride, int64_t B, int F, int D,
                                       float* __restrict__ y) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* row = v + b * stride;
  float acc = 0.f;
  for (int k = 0; k < D; ++k) {
    float s = 0.f, q = 0.f;
    for (int f = 0; f < F; ++f) {
      const float x = row[f * D + k];
      s += x;
      q += x * x;
    }
    acc += s * s - q;
  }
  y[b] = 0.5f * acc;
}
__global__ void fm2_bwd_generic_kernel(const float* __restrict__ v, int64_t stride,
                                       const float* __restrict__ gy, const float* __restrict__ grad_in,
                                       int64_t gi_stride, int64_t B, int F, int D,
                                       float* __restrict__ grad_v, int64_t gv_stride) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int64_t b = idx / D;
  const int k = (int)(idx - b * D);
  const float* row = v + b * stride;
  float s = 0.f;
  for (int f = 0; f < F; ++f) s += row[f * D + k];
  const float g = gy[b];
  for (int f = 0; f < F; ++f) {
    float o = g * (s - row[f * D + k]);
    if (grad_in) o += grad_in[b * gi_stride + f * D + k];
    grad_v[b * gv_stride + f * D + k] = o;
  }
}

static bool fm2_vector_ok(int D, const void* p0, int64_t s0, const void* p1, int64_t s1, const void* p2,
                          int64_t s2) {
  if (D < 4 || D > 128 || (D & (D - 1)) != 0) return false;
  if (!aligned16(p0) || (s0 % 4) != 0) return false;
  if (p1 && (!aligned16(p1) || (s1 % 4) != 0)) return false;
  if (p2 && (!aligned16(p2) || (s2 % 4) != 0)) return false;
  return true;
}

}  // namespace ptrec

using namespace ptrec;

extern "C" int ptrec_fm2_fwd(const float* v, int64_t v_row_stride, int64_t B, int32_t F, int32_t D, float* y,
                             void* stream) {
  PTREC_CHECK_ARG(v && y && B >= 0 && F >= 1 && D >= 1, PTREC_EINVAL, "fm2_fwd: bad argument");
  PTREC_CHECK_ARG(v_row_stride >= (int64_t)F * D, PTREC_EINVAL, "fm2_fwd: row stride smaller than F*D");
  if (B == 0) return PTREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (fm2_vector_ok(D, v, v_row_stride, nullptr, 0, nullptr, 0)) {
    fm2_fwd_kernel<<<(unsigned)ceil_div(B, kFmWarps), kFmWarps * 32, 0, st>>>(v, v_row_stride, B, F * D / 4,
                                                                              D / 4, y);
    PTREC_LAUNCH_CHECK("fm2_fwd_kernel");
  } else {
    fm2_fwd_generic_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(v, v_row_stride, B, F, D, y);
    PTREC_LAUNCH_CHECK("fm2_fwd_generic_kernel");
  }
  return PTREC_OK;
}

extern "C" int ptrec_fm2_bwd(const float* v, int64_t v_row_stride, const float* gy, const float* grad_in,
                             int64_t grad_in_row_stride, int64_t B, int32_t F, int32_t D, float* grad_v,
                             int64_t grad_v_row_stride, void* stream) {
  PTREC_CHECK_ARG(v && gy && grad_v && B >= 0 && F >= 1 && D >= 1, PTREC_EINVAL, "fm2_bwd: bad argument");
  PTREC_CHECK_ARG(v_row_stride >= (int64_t)F * D && grad_v_row_stride >= (int64_t)F * D, PTREC_EINVAL,
                  "fm2_bwd: row stride smaller than F*D");
  if (B == 0) return PTREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (fm2_vector_ok(D, v, v_row_stride, grad_in, grad_in_row_stride, grad_v, grad_v_row_stride)) {
    fm2_bwd_kernel<<<(unsigned)ceil_div(B, kFmWarps), kFmWarps * 32, 0, st>>>(
        v, v_row_stride, gy, grad_in, grad_in_row_stride, B, F * D / 4, D / 4, grad_v, grad_v_row_stride);
    PTREC_LAUNCH_CHECK("fm2_bwd_kernel");
  } else {
    fm2_bwd_generic_kernel<<<(unsigned)ceil_div(B * D, 256), 256, 0, st>>>(
        v, v_row_stride, gy, grad_in, grad_in_row_stride, B, F, D, grad_v, grad_v_row_stride);
    PTREC_LAUNCH_CHECK("fm2_bwd_generic_kernel");
  }
  return PTREC_OK;
}
