// K5 companions: the element-wise work around the DCN-v2 cross GEMMs (dcn_cross.cu), one launch each instead of the
// ~12 library launches per cross layer they replace (padding / casting fills and copies, addcmul, float() round trips,
// column sums, transposed weight copies: 2.2 ms of the 5.4 ms cfg3 step, profiles/r2_step_kernels_dcn_before.txt).
// No reference counterpart (the reference has no DCN; torchrec/model/layer/* holds Dense / MLP only).  Bound: HBM.
//
//   prep_weight   W fp32 [d, d], b fp32 [d]  ->  bf16 W [dp, dp], bf16 W^T [dp, dp], fp32 b [dp]   (zero padded)
//   pack_input    x fp32 [B, d] (pitch)      ->  bf16 [B, dp]                                       (zero padded)
//   unpack        bf16 [B, dp]               ->  fp32 [B, d]
//   bwd_init      g fp32 [B, d], x0 bf16     ->  g_out = bf16(g) padded, g_u = g_out * x0
//   bwd_layer     g_x0 (+)= g_out * u  (fp32 [B, dp]);  bias gradient = column sums of g_u (fp32, two fixed-order
//                 levels: per-row-block partials, then one sum per column: bit-reproducible)
//   bwd_final     out fp32 [B, d] = g_x0 + g_out
//   head_fwd      y fp32 [B] = x_L[b, :d] . w           (the cross half of DCN's closing Linear, read from the bf16 x_L:
//                 replaces unpack + a [B, d] x [d] library product)
//   head_bwd      g_out = bf16(g_y (x) w) padded, g_u = g_out * x0, g_w = sum_b g_y[b] x_L[b, :]   (replaces the library
//                 product's two backward kernels, the fp32 [B, d] gradient and bwd_init)
// The fp32 side of pack_input / unpack / bwd_init / bwd_final has rows of d floats (d = 845 at cfg3: 4-byte aligned
// only), so those kernels give every thread ONE column and walk rows: scalar accesses, consecutive lanes on consecutive
// addresses (8 columns per thread made each lane store 32 scattered bytes: 96 - 102 us against a 26 - 42 us floor).
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"

namespace ptrec {

constexpr int kGlThreads = 128;  // 8 bf16 columns per thread: 1024 columns per CTA pass
constexpr int kGlRows = 32;      // rows per CTA of bwd_layer (column sums accumulate in registers over them)

struct Bf8 {
  float v[8];
};
__device__ __forceinline__ Bf8 load_bf8(const __nv_bfloat16* p) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
  Bf8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void store_bf8(__nv_bfloat16* p, const Bf8& r) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}

__global__ void __launch_bounds__(256)
dcn_prep_weight_kernel(const float* __restrict__ W, const float* __restrict__ b, int d, int dp,
                       __nv_bfloat16* __restrict__ w16, __nv_bfloat16* __restrict__ w16t, float* __restrict__ bp) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    const float v = (r < d && c < d) ? W[(int64_t)r * d + c] : 0.f;
    tile[i][threadIdx.x] = v;
    if (r < dp && c < dp) w16[(int64_t)r * dp + c] = __float2bfloat16(v);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;  // transposed element (c, r)
    if (c < dp && r < dp) w16t[(int64_t)c * dp + r] = __float2bfloat16(tile[threadIdx.x][i]);
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    for (int j = threadIdx.y * 32 + threadIdx.x; j < dp; j += 32 * blockDim.y) bp[j] = j < d ? b[j] : 0.f;
  }
}

constexpr int kElRows = 16;  // rows per CTA of the one-column-per-thread kernels: grid (ceil(B / kElRows), ceil(dp / 256))

__global__ void __launch_bounds__(256)
dcn_pack_input_kernel(const float* __restrict__ x, int64_t ldx, int64_t B, int d, int dp, __nv_bfloat16* __restrict__ out) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= dp) return;
  const int64_t r0 = (int64_t)blockIdx.x * kElRows, r1 = min(B, r0 + kElRows);
  float v[kElRows];
#pragma unroll
  for (int k = 0; k < kElRows; ++k) v[k] = (c < d && r0 + k < r1) ? __ldg(x + (r0 + k) * ldx + c) : 0.f;
#pragma unroll
  for (int k = 0; k < kElRows; ++k)
    if (r0 + k < r1) out[(r0 + k) * dp + c] = __float2bfloat16(v[k]);
}

__global__ void __launch_bounds__(256)
dcn_unpack_kernel(const __nv_bfloat16* __restrict__ x, int64_t B, int d, int dp, float* __restrict__ out) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= d) return;
  const int64_t r0 = (int64_t)blockIdx.x * kElRows, r1 = min(B, r0 + kElRows);
  __nv_bfloat16 v[kElRows];
#pragma unroll
  for (int k = 0; k < kElRows; ++k)
    if (r0 + k < r1) v[k] = x[(r0 + k) * dp + c];
#pragma unroll
  for (int k = 0; k < kElRows; ++k)
    if (r0 + k < r1) out[(r0 + k) * d + c] = __bfloat162float(v[k]);
}

__global__ void __launch_bounds__(256)
dcn_bwd_init_kernel(const float* __restrict__ g, int64_t ldg, const __nv_bfloat16* __restrict__ x0, int64_t B, int d,
                    int dp, __nv_bfloat16* __restrict__ g_out, __nv_bfloat16* __restrict__ g_u) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= dp) return;
  const int64_t r0 = (int64_t)blockIdx.x * kElRows, r1 = min(B, r0 + kElRows);
  float gv[kElRows];
  __nv_bfloat16 xv[kElRows];
#pragma unroll
  for (int k = 0; k < kElRows; ++k) {
    gv[k] = (c < d && r0 + k < r1) ? __ldg(g + (r0 + k) * ldg + c) : 0.f;
    if (r0 + k < r1) xv[k] = x0[(r0 + k) * dp + c];
  }
#pragma unroll
  for (int k = 0; k < kElRows; ++k) {
    if (r0 + k < r1) {
      // g_u = bf16(g_out) * x0 with g_out already rounded to bf16, as the unfused `g_out * x0` on bf16 tensors computes it
      const __nv_bfloat16 go = __float2bfloat16(gv[k]);
      g_out[(r0 + k) * dp + c] = go;
      g_u[(r0 + k) * dp + c] = __float2bfloat16(__bfloat162float(go) * __bfloat162float(xv[k]));
    }
  }
}

// grid (ceil(dp / 1024), ceil(B / kGlRows)); partial [gridDim.y][dp]
__global__ void __launch_bounds__(kGlThreads)
dcn_bwd_layer_kernel(const __nv_bfloat16* __restrict__ g_out, const __nv_bfloat16* __restrict__ u,
                     const __nv_bfloat16* __restrict__ g_u, int64_t B, int dp, float* __restrict__ g_x0, int accumulate,
                     float* __restrict__ partial) {
  const int c = (blockIdx.x * kGlThreads + threadIdx.x) * 8;
  if (c >= dp) return;
  const int64_t r0 = (int64_t)blockIdx.y * kGlRows;
  const int64_t r1 = min(B, r0 + kGlRows);
  float cs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[i] = 0.f;
  for (int64_t r = r0; r < r1; r += 2) {
    Bf8 a[2], b[2], gu[2];
    float4 acc[2][2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (r + k < r1) {
        const int64_t off = (r + k) * dp + c;
        a[k] = load_bf8(g_out + off);
        b[k] = load_bf8(u + off);
        gu[k] = load_bf8(g_u + off);
        if (accumulate) {
          acc[k][0] = *reinterpret_cast<const float4*>(g_x0 + off);
          acc[k][1] = *reinterpret_cast<const float4*>(g_x0 + off + 4);
        } else {
          acc[k][0] = acc[k][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (r + k < r1) {
        const int64_t off = (r + k) * dp + c;
        acc[k][0].x += a[k].v[0] * b[k].v[0]; acc[k][0].y += a[k].v[1] * b[k].v[1];
        acc[k][0].z += a[k].v[2] * b[k].v[2]; acc[k][0].w += a[k].v[3] * b[k].v[3];
        acc[k][1].x += a[k].v[4] * b[k].v[4]; acc[k][1].y += a[k].v[5] * b[k].v[5];
        acc[k][1].z += a[k].v[6] * b[k].v[6]; acc[k][1].w += a[k].v[7] * b[k].v[7];
        *reinterpret_cast<float4*>(g_x0 + off) = acc[k][0];
        *reinterpret_cast<float4*>(g_x0 + off + 4) = acc[k][1];
#pragma unroll
        for (int i = 0; i < 8; ++i) cs[i] += gu[k].v[i];
      }
    }
  }
  float* p = partial + (int64_t)blockIdx.y * dp + c;
  *reinterpret_cast<float4*>(p) = make_float4(cs[0], cs[1], cs[2], cs[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(cs[4], cs[5], cs[6], cs[7]);
}

// gb[j] = sum over row blocks (in order) of partial[blk][j]; one warp per 32 columns, lanes stride the blocks, then a
// fixed xor tree
__global__ void __launch_bounds__(256)
dcn_colsum_finish_kernel(const float* __restrict__ partial, int n_blk, int dp, int d, float* __restrict__ gb) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= d) return;
  float s = 0.f;
  for (int b = lane; b < n_blk; b += 32) s += partial[(int64_t)b * dp + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) gb[j] = s;
}

__global__ void __launch_bounds__(256)
dcn_bwd_final_kernel(const float* __restrict__ g_x0, const __nv_bfloat16* __restrict__ g_out, int64_t B, int d, int dp,
                     float* __restrict__ out) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= d) return;
  const int64_t r0 = (int64_t)blockIdx.x * kElRows, r1 = min(B, r0 + kElRows);
  float a[kElRows];
  __nv_bfloat16 b[kElRows];
#pragma unroll
  for (int k = 0; k < kElRows; ++k) {
    if (r0 + k < r1) {
      a[k] = __ldg(g_x0 + (r0 + k) * dp + c);
      b[k] = g_out[(r0 + k) * dp + c];
    }
  }
#pragma unroll
  for (int k = 0; k < kElRows; ++k)
    if (r0 + k < r1) out[(r0 + k) * d + c] = a[k] + __bfloat162float(b[k]);
}

// y[b] = sum_{c < d} x[b, c] w[c]: one warp per row, 8 columns per lane and step, w staged in shared memory (zero padded)
__global__ void __launch_bounds__(256)
dcn_head_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t B, int d, int dp, const float* __restrict__ w,
                    float* __restrict__ y) {
  extern __shared__ float s_w[];  // [dp]
  for (int j = threadIdx.x; j < dp; j += 256) s_w[j] = j < d ? w[j] : 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < B; r += (int64_t)gridDim.x * 8) {
    float acc = 0.f;
    for (int c = lane * 8; c < dp; c += 256) {
      const Bf8 v = load_bf8(x + r * dp + c);
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + c), w1 = *reinterpret_cast<const float4*>(s_w + c + 4);
      acc += v.v[0] * w0.x + v.v[1] * w0.y + v.v[2] * w0.z + v.v[3] * w0.w + v.v[4] * w1.x + v.v[5] * w1.y +
             v.v[6] * w1.z + v.v[7] * w1.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[r] = acc;
  }
}

// grid (ceil(dp / 1024), ceil(B / kGlRows)); partial [gridDim.y][dp]: per-row-block sums of g_y[b] x_L[b, c]
__global__ void __launch_bounds__(kGlThreads)
dcn_head_bwd_kernel(const float* __restrict__ g_y, const float* __restrict__ w, const __nv_bfloat16* __restrict__ x_l,
                    const __nv_bfloat16* __restrict__ x0, int64_t B, int d, int dp, __nv_bfloat16* __restrict__ g_out,
                    __nv_bfloat16* __restrict__ g_u, float* __restrict__ partial) {
  const int c = (blockIdx.x * kGlThreads + threadIdx.x) * 8;
  if (c >= dp) return;
  const int64_t r0 = (int64_t)blockIdx.y * kGlRows;
  const int64_t r1 = min(B, r0 + kGlRows);
  float wv[8], cs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    wv[i] = c + i < d ? w[c + i] : 0.f;
    cs[i] = 0.f;
  }
  for (int64_t r = r0; r < r1; r += 2) {
    Bf8 xl[2], xz[2];
    float gy[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (r + k < r1) {
        gy[k] = g_y[r + k];
        xl[k] = load_bf8(x_l + (r + k) * dp + c);
        xz[k] = load_bf8(x0 + (r + k) * dp + c);
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (r + k < r1) {
        Bf8 go, gu;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          go.v[i] = __bfloat162float(__float2bfloat16(gy[k] * wv[i]));  // g_out is a bf16 tensor: g_u multiplies the rounded value
          gu.v[i] = go.v[i] * xz[k].v[i];
          cs[i] += gy[k] * xl[k].v[i];
        }
        store_bf8(g_out + (r + k) * dp + c, go);
        store_bf8(g_u + (r + k) * dp + c, gu);
      }
    }
  }
  float* p = partial + (int64_t)blockIdx.y * dp + c;
  *reinterpret_cast<float4*>(p) = make_float4(cs[0], cs[1], cs[2], cs[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(cs[4], cs[5], cs[6], cs[7]);
}

}  // namespace ptrec

using namespace ptrec;

static int check_dims(const char* what, int64_t B, int d, int dp) {
  PTREC_CHECK_ARG(B >= 0 && d >= 1 && dp >= d && dp % 8 == 0, PTREC_EINVAL, "%s: bad sizes B=%lld d=%d dp=%d (dp: d rounded up to 8)",
                  what, (long long)B, d, dp);
  return PTREC_OK;
}

extern "C" int ptrec_dcn_prep_weight(const float* W, const float* b, int32_t d, int32_t dp, void* w16, void* w16t,
                                     float* bias_pad, void* stream) {
  int rc = check_dims("dcn_prep_weight", 0, d, dp);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(W && b && w16 && w16t && bias_pad, PTREC_EINVAL, "dcn_prep_weight: null pointer");
  dim3 tb(32, 8), tg((unsigned)ceil_div(dp, 32), (unsigned)ceil_div(dp, 32));
  dcn_prep_weight_kernel<<<tg, tb, 0, (cudaStream_t)stream>>>(W, b, d, dp, reinterpret_cast<__nv_bfloat16*>(w16),
                                                              reinterpret_cast<__nv_bfloat16*>(w16t), bias_pad);
  PTREC_LAUNCH_CHECK("dcn_prep_weight_kernel");
  return PTREC_OK;
}

extern "C" int ptrec_dcn_pack_input(const float* x, int64_t ldx, int64_t B, int32_t d, int32_t dp, void* out, void* stream) {
  int rc = check_dims("dcn_pack_input", B, d, dp);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(x && out && aligned16(out), PTREC_EINVAL, "dcn_pack_input: null / misaligned pointer");
  if (B == 0) return PTREC_OK;
  dcn_pack_input_kernel<<<dim3((unsigned)ceil_div(B, kElRows), (unsigned)ceil_div(dp, 256)), 256, 0, (cudaStream_t)stream>>>(
      x, ldx, B, d, dp, reinterpret_cast<__nv_bfloat16*>(out));
  PTREC_LAUNCH_CHECK("dcn_pack_input_kernel");
  return PTREC_OK;
}

extern "C" int ptrec_dcn_unpack(const void* x, int64_t B, int32_t d, int32_t dp, float* out, void* stream) {
  int rc = check_dims("dcn_unpack", B, d, dp);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(x && out && aligned16(x), PTREC_EINVAL, "dcn_unpack: null / misaligned pointer");
  if (B == 0) return PTREC_OK;
  dcn_unpack_kernel<<<dim3((unsigned)ceil_div(B, kElRows), (unsigned)ceil_div(d, 256)), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), B, d, dp, out);
  PTREC_LAUNCH_CHECK("dcn_unpack_kernel");
  return PTREC_OK;
}

extern "C" int ptrec_dcn_bwd_init(const float* g, int64_t ldg, const void* x0, int64_t B, int32_t d, int32_t dp,
                                  void* g_out, void* g_u, void* stream) {
  int rc = check_dims("dcn_bwd_init", B, d, dp);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(g && x0 && g_out && g_u && aligned16(x0) && aligned16(g_out) && aligned16(g_u), PTREC_EINVAL,
                  "dcn_bwd_init: null / misaligned pointer");
  if (B == 0) return PTREC_OK;
  dcn_bwd_init_kernel<<<dim3((unsigned)ceil_div(B, kElRows), (unsigned)ceil_div(dp, 256)), 256, 0, (cudaStream_t)stream>>>(
      g, ldg, reinterpret_cast<const __nv_bfloat16*>(x0), B, d, dp, reinterpret_cast<__nv_bfloat16*>(g_out),
      reinterpret_cast<__nv_bfloat16*>(g_u));
  PTREC_LAUNCH_CHECK("dcn_bwd_init_kernel");
  return PTREC_OK;
}

extern "C" size_t ptrec_dcn_bwd_layer_workspace_bytes(int64_t B, int32_t dp) {
  return align_up((size_t)ceil_div(B, kGlRows) * (size_t)dp * sizeof(float), 256);
}

extern "C" int ptrec_dcn_bwd_layer(const void* g_out, const void* u, const void* g_u, int64_t B, int32_t d, int32_t dp,
                                   float* g_x0, int32_t accumulate, float* grad_bias, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  int rc = check_dims("dcn_bwd_layer", B, d, dp);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(g_out && u && g_u && g_x0 && grad_bias && workspace, PTREC_EINVAL, "dcn_bwd_layer: null pointer");
  PTREC_CHECK_ARG(aligned16(g_out) && aligned16(u) && aligned16(g_u) && aligned16(g_x0) && aligned16(workspace), PTREC_EALIGN,
                  "dcn_bwd_layer: 16-byte alignment");
  PTREC_CHECK_ARG(workspace_bytes >= ptrec_dcn_bwd_layer_workspace_bytes(B, dp), PTREC_EWORKSPACE,
                  "dcn_bwd_layer: workspace too small");
  if (B == 0) return PTREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int n_blk = (int)ceil_div(B, kGlRows);
  dim3 grid((unsigned)ceil_div(dp, kGlThreads * 8), (unsigned)n_blk);
  float* partial = reinterpret_cast<float*>(workspace);
  dcn_bwd_layer_kernel<<<grid, kGlThreads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g_out),
                                                     reinterpret_cast<const __nv_bfloat16*>(u),
                                                     reinterpret_cast<const __nv_bfloat16*>(g_u), B, dp, g_x0, accumulate,
                                                     partial);
  PTREC_LAUNCH_CHECK("dcn_bwd_layer_kernel");
  dcn_colsum_finish_kernel<<<(unsigned)ceil_div(d, 8), 256, 0, st>>>(partial, n_blk, dp, d, grad_bias);
  PTREC_LAUNCH_CHECK("dcn_colsum_finish_kernel");
  return PTREC_OK;
}

extern "C" int ptrec_dcn_bwd_final(const float* g_x0, const void* g_out, int64_t B, int32_t d, int32_t dp, float* out,
                                   void* stream) {
  int rc = check_dims("dcn_bwd_final", B, d, dp);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(g_x0 && g_out && out && aligned16(g_out), PTREC_EINVAL, "dcn_bwd_final: null / misaligned pointer");
  if (B == 0) return PTREC_OK;
  dcn_bwd_final_kernel<<<dim3((unsigned)ceil_div(B, kElRows), (unsigned)ceil_div(d, 256)), 256, 0, (cudaStream_t)stream>>>(
      g_x0, reinterpret_cast<const __nv_bfloat16*>(g_out), B, d, dp, out);
  PTREC_LAUNCH_CHECK("dcn_bwd_final_kernel");
  return PTREC_OK;
}

extern "C" int ptrec_dcn_head_fwd(const void* x_l, int64_t B, int32_t d, int32_t dp, const float* w, float* y, void* stream) {
  int rc = check_dims("dcn_head_fwd", B, d, dp);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(x_l && w && y && aligned16(x_l), PTREC_EINVAL, "dcn_head_fwd: null / misaligned pointer");
  PTREC_CHECK_ARG(dp <= 8192, PTREC_EUNSUPPORTED, "dcn_head_fwd: d <= 8192");
  if (B == 0) return PTREC_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(B, 8), (int64_t)sms * 8);
  dcn_head_fwd_kernel<<<grid, 256, (size_t)dp * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x_l), B, d, dp, w, y);
  PTREC_LAUNCH_CHECK("dcn_head_fwd_kernel");
  return PTREC_OK;
}

// workspace: ptrec_dcn_bwd_layer_workspace_bytes(B, dp)
extern "C" int ptrec_dcn_head_bwd(const float* g_y, const float* w, const void* x_l, const void* x0, int64_t B, int32_t d,
                                  int32_t dp, void* g_out, void* g_u, float* grad_w, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  int rc = check_dims("dcn_head_bwd", B, d, dp);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(g_y && w && x_l && x0 && g_out && g_u && grad_w && workspace, PTREC_EINVAL, "dcn_head_bwd: null pointer");
  PTREC_CHECK_ARG(aligned16(x_l) && aligned16(x0) && aligned16(g_out) && aligned16(g_u) && aligned16(workspace), PTREC_EALIGN,
                  "dcn_head_bwd: 16-byte alignment");
  PTREC_CHECK_ARG(workspace_bytes >= ptrec_dcn_bwd_layer_workspace_bytes(B, dp), PTREC_EWORKSPACE,
                  "dcn_head_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    PTREC_CUDA(cudaMemsetAsync(grad_w, 0, (size_t)d * sizeof(float), st));
    return PTREC_OK;
  }
  const int n_blk = (int)ceil_div(B, kGlRows);
  dim3 grid((unsigned)ceil_div(dp, kGlThreads * 8), (unsigned)n_blk);
  float* partial = reinterpret_cast<float*>(workspace);
  dcn_head_bwd_kernel<<<grid, kGlThreads, 0, st>>>(g_y, w, reinterpret_cast<const __nv_bfloat16*>(x_l),
                                                    reinterpret_cast<const __nv_bfloat16*>(x0), B, d, dp,
                                                    reinterpret_cast<__nv_bfloat16*>(g_out),
                                                    reinterpret_cast<__nv_bfloat16*>(g_u), partial);
  PTREC_LAUNCH_CHECK("dcn_head_bwd_kernel");
  dcn_colsum_finish_kernel<<<(unsigned)ceil_div(d, 8), 256, 0, st>>>(partial, n_blk, dp, d, grad_w);
  PTREC_LAUNCH_CHECK("dcn_colsum_finish_kernel");
  return PTREC_OK;
}
