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
#include <cuda_bf16.h>

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

// one thread per 8 output columns
__global__ void __launch_bounds__(256)
dcn_pack_input_kernel(const float* __restrict__ x, int64_t ldx, int64_t B, int d, int dp, __nv_bfloat16* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int per_row = dp / 8;
  if (e >= B * per_row) return;
  const int64_t r = e / per_row;
  const int c = (int)(e - r * per_row) * 8;
  Bf8 v;
#pragma unroll
  for (int i = 0; i < 8; ++i) v.v[i] = (c + i < d) ? x[r * ldx + c + i] : 0.f;
  store_bf8(out + r * dp + c, v);
}

__global__ void __launch_bounds__(256)
dcn_unpack_kernel(const __nv_bfloat16* __restrict__ x, int64_t B, int d, int dp, float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int per_row = dp / 8;
  if (e >= B * per_row) return;
  const int64_t r = e / per_row;
  const int c = (int)(e - r * per_row) * 8;
  const Bf8 v = load_bf8(x + r * dp + c);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (c + i < d) out[r * d + c + i] = v.v[i];
}

__global__ void __launch_bounds__(256)
dcn_bwd_init_kernel(const float* __restrict__ g, int64_t ldg, const __nv_bfloat16* __restrict__ x0, int64_t B, int d,
                    int dp, __nv_bfloat16* __restrict__ g_out, __nv_bfloat16* __restrict__ g_u) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int per_row = dp / 8;
  if (e >= B * per_row) return;
  const int64_t r = e / per_row;
  const int c = (int)(e - r * per_row) * 8;
  Bf8 go;
#pragma unroll
  for (int i = 0; i < 8; ++i) go.v[i] = (c + i < d) ? g[r * ldg + c + i] : 0.f;
  store_bf8(g_out + r * dp + c, go);
  // g_u = bf16(g_out) * x0 with g_out already rounded to bf16, as the unfused `g_out * x0` on bf16 tensors computes it
  const Bf8 gr = load_bf8(g_out + r * dp + c), xv = load_bf8(x0 + r * dp + c);
  Bf8 gu;
#pragma unroll
  for (int i = 0; i < 8; ++i) gu.v[i] = gr.v[i] * xv.v[i];
  store_bf8(g_u + r * dp + c, gu);
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
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int per_row = dp / 8;
  if (e >= B * per_row) return;
  const int64_t r = e / per_row;
  const int c = (int)(e - r * per_row) * 8;
  const Bf8 go = load_bf8(g_out + r * dp + c);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (c + i < d) out[r * d + c + i] = g_x0[r * dp + c + i] + go.v[i];
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
  dcn_pack_input_kernel<<<(unsigned)ceil_div(B * (dp / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      x, ldx, B, d, dp, reinterpret_cast<__nv_bfloat16*>(out));
  PTREC_LAUNCH_CHECK("dcn_pack_input_kernel");
  return PTREC_OK;
}

extern "C" int ptrec_dcn_unpack(const void* x, int64_t B, int32_t d, int32_t dp, float* out, void* stream) {
  int rc = check_dims("dcn_unpack", B, d, dp);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(x && out && aligned16(x), PTREC_EINVAL, "dcn_unpack: null / misaligned pointer");
  if (B == 0) return PTREC_OK;
  dcn_unpack_kernel<<<(unsigned)ceil_div(B * (dp / 8), 256), 256, 0, (cudaStream_t)stream>>>(
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
  dcn_bwd_init_kernel<<<(unsigned)ceil_div(B * (dp / 8), 256), 256, 0, (cudaStream_t)stream>>>(
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
  dcn_bwd_final_kernel<<<(unsigned)ceil_div(B * (dp / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      g_x0, reinterpret_cast<const __nv_bfloat16*>(g_out), B, d, dp, out);
  PTREC_LAUNCH_CHECK("dcn_bwd_final_kernel");
  return PTREC_OK;
}
