// K8: the FM "head" around K3 — everything between the embedding gathers and the DNN tower / the loss, fused.
//   forward   logit[b] = 0.5 * sum_k((sum_f v)^2 - sum_f v^2) + sum_f w1[b,f] + x[b].wd + bias
//             and, for DeepFM, the tower input  deep_in[b] = [ v[b, :] | x[b, :] ]  written in the same pass
//   backward  gv[b] = g[b] * (S[b] - v[b]) + g_deep_in[b, :F*D]     (the tower's input gradient is added on the fly)
//             gw1[b,f] = g[b],  gx[b,j] = g[b]*wd[j] + g_deep_in[b, F*D+j],  g_wd[j] = sum_b g[b]*x[b,j],  g_bias = sum_b g[b]
//   rowdot    y[b] = h[b].w   (the Linear(H, 1, bias=False) that closes the tower), backward g_h = g (x) w, g_w = h^T g
// Replaces the chain sum / cat / gemv / add / slice kernels of the model code (SVDPP.py:65-66 idiom: biases + dot,
// NCF.py:68-74: concat -> MLP -> Linear(., 1)) — ~20 latency-bound torch launches per step at cfg2.
// Bound: HBM (each pass streams [B, F*D] once or twice).  Reductions over the batch are two-level with a fixed order.
#include "common.cuh"
#include "split3.cuh"

namespace ptrec {

constexpr int kHdWarps = 8;
constexpr int kHdMaxChunks = 16;  // float4 chunks of v cached per lane in the backward (F*D <= 2048)
constexpr int kHdMaxDense = 128;  // dense features (4 per lane)
constexpr int kHdMaxH = 1024;     // rowdot width (8 float4 chunks per lane)

__device__ __forceinline__ float4 hd_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 hd_reduce_S(float4 s, int lanes_per_row) {
  for (int o = lanes_per_row; o < 32; o <<= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
    s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, o);
    s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
  }
  return s;
}

__global__ void __launch_bounds__(kHdWarps * 32)
fm_head_fwd_kernel(const float* __restrict__ v, int64_t vs, const float* __restrict__ w1, int64_t w1s,
                   const float* __restrict__ x, int64_t xs, const float* __restrict__ wd,
                   const float* __restrict__ bias, int64_t B, int F, int n_chunks, int lanes_per_row, int nd,
                   float* __restrict__ logit, float* __restrict__ deep_in, int64_t ds,
                   __nv_bfloat16* __restrict__ planes, int64_t pl_ld, int64_t pl_plane,
                   const float* __restrict__ h2_scale, uint32_t* __restrict__ h2_max) {
  // h2_scale != null: `planes` are the TWO fp16 planes of the K6 fused tower (tc_linear.cu), split with the carried
  // scale *h2_scale; *h2_max is raised to max |tower input| for the next roll.  Otherwise three exact bf16 planes.
  const int64_t b = (int64_t)blockIdx.x * kHdWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const float* row = v + b * vs;
  float* drow = deep_in ? deep_in + b * ds : nullptr;
  const bool h2 = h2_scale != nullptr;
  const float hs = h2 ? *h2_scale : 1.f;
  float amax = 0.f;
  __nv_bfloat16* prow = (planes && !h2) ? planes + b * pl_ld : nullptr;
  unsigned short* hrow = (planes && h2) ? reinterpret_cast<unsigned short*>(planes) + b * pl_ld : nullptr;
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
      const int q = q0 + u * 32;
      S = hd_add(S, t[u]);
      Q += t[u].x * t[u].x + t[u].y * t[u].y + t[u].z * t[u].z + t[u].w * t[u].w;
      if (drow && q < n_chunks) st_f4(drow + q * 4, t[u]);
      if (prow && q < n_chunks) {
        const float tv[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
        split3_store4(tv, prow + q * 4, pl_plane);
      }
      if (hrow && q < n_chunks) {
        const float tv[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
        uint32_t p0[2], p1[2];
        split2h_pair(tv[0], tv[1], hs, p0[0], p1[0]);
        split2h_pair(tv[2], tv[3], hs, p0[1], p1[1]);
#pragma unroll
        for (int i = 0; i < 4; ++i) amax = fmaxf(amax, fabsf(tv[i]));
        *reinterpret_cast<uint2*>(hrow + q * 4) = make_uint2(p0[0], p0[1]);
        *reinterpret_cast<uint2*>(hrow + pl_plane + q * 4) = make_uint2(p1[0], p1[1]);
      }
    }
  }
  S = hd_reduce_S(S, lanes_per_row);
  float r = (lane < lanes_per_row) ? (S.x * S.x + S.y * S.y + S.z * S.z + S.w * S.w) : 0.f;
  float acc = 0.5f * (r - Q);
  if (w1 != nullptr)
    for (int f = lane; f < F; f += 32) acc += w1[b * w1s + f];
  if (x != nullptr) {
    for (int j = lane; j < nd; j += 32) {
      const float xv = x[b * xs + j];
      if (wd != nullptr) acc += xv * wd[j];
      if (drow) drow[n_chunks * 4 + j] = xv;
    }
  }
  if (prow) {  // dense columns and the zero pad up to the plane pitch
    for (int j = lane; n_chunks * 4 + j < pl_ld; j += 32) {
      const float xv = (x != nullptr && j < nd) ? x[b * xs + j] : 0.f;
      __nv_bfloat16 p0, p1, p2;
      split3(xv, p0, p1, p2);
      __nv_bfloat16* o = prow + n_chunks * 4 + j;
      o[0] = p0;
      o[pl_plane] = p1;
      o[2 * pl_plane] = p2;
    }
  }
  if (hrow) {  // dense columns and the zero pad up to the plane pitch
    for (int j = lane; n_chunks * 4 + j < pl_ld; j += 32) {
      const float xv = (x != nullptr && j < nd) ? x[b * xs + j] : 0.f;
      __half a0, a1;
      split2h(xv, hs, a0, a1);
      hrow[n_chunks * 4 + j] = __half_as_ushort(a0);
      hrow[pl_plane + n_chunks * 4 + j] = __half_as_ushort(a1);
      amax = fmaxf(amax, fabsf(xv));
    }
    float m = amax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m > 0.f && h2_max != nullptr) atomicMax(h2_max, __float_as_uint(m));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) logit[b] = acc + (bias ? bias[0] : 0.f);
}

// grid-stride over samples; per-CTA partial sums of g*x[.,j] (j < nd) and of g (column nd) -> part[blockIdx][nd+1]
// MAXC = float4 chunks of a row cached per lane (n_chunks <= 32*MAXC): v and the tower gradient are requested together
template <int MAXC>
__global__ void __launch_bounds__(kHdWarps * 32)
fm_head_bwd_kernel(const float* __restrict__ v, int64_t vs, const float* __restrict__ x, int64_t xs,
                   const float* __restrict__ wd, const float* __restrict__ g, const float* __restrict__ gdi,
                   int64_t gds, int64_t B, int F, int n_chunks, int lanes_per_row, int nd, float* __restrict__ gv,
                   int64_t gvs, float* __restrict__ gw1, float* __restrict__ gx, int64_t gxs,
                   float* __restrict__ part) {
  __shared__ float s_part[kHdWarps][kHdMaxDense + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float pw[4] = {0.f, 0.f, 0.f, 0.f};  // columns lane, lane+32, lane+64, lane+96 of g*x
  float pb = 0.f;
  for (int64_t b = (int64_t)blockIdx.x * kHdWarps + warp; b < B; b += (int64_t)gridDim.x * kHdWarps) {
    const float* row = v + b * vs;
    const float* irow = gdi ? gdi + b * gds : nullptr;
    float4 c[MAXC], d[MAXC];
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      const int q = lane + u * 32;
      const bool in = q < n_chunks;
      c[u] = in ? ldg_stream_f4(row + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      d[u] = (in && irow) ? ldg_stream_f4(irow + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float gb = g[b];
    float xv[4] = {0.f, 0.f, 0.f, 0.f};
    if (x != nullptr) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = lane + u * 32;
        if (j < nd) xv[u] = x[b * xs + j];
      }
    }
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < MAXC; ++u) S = hd_add(S, c[u]);
    S = hd_reduce_S(S, lanes_per_row);
    float* orow = gv + b * gvs;
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      const int q = lane + u * 32;
      if (q < n_chunks) {
        st_f4(orow + q * 4, make_float4(gb * (S.x - c[u].x) + d[u].x, gb * (S.y - c[u].y) + d[u].y,
                                        gb * (S.z - c[u].z) + d[u].z, gb * (S.w - c[u].w) + d[u].w));
      }
    }
    if (gw1 != nullptr)
      for (int f = lane; f < F; f += 32) gw1[b * F + f] = gb;
    if (x != nullptr) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = lane + u * 32;
        if (j < nd) {
          pw[u] += gb * xv[u];
          if (gx != nullptr) gx[b * gxs + j] = (wd ? gb * wd[j] : 0.f) + (irow ? irow[n_chunks * 4 + j] : 0.f);
        }
      }
    }
    pb += gb;  // same value in every lane
  }
  if (part == nullptr) return;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = lane + u * 32;
    if (j < nd) s_part[warp][j] = pw[u];
  }
  if (lane == 0) s_part[warp][nd] = pb;
  __syncthreads();
  for (int j = threadIdx.x; j <= nd; j += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kHdWarps; ++w) s += s_part[w][j];  // fixed order
    part[(int64_t)blockIdx.x * (nd + 1) + j] = s;
  }
}

// out[c] = sum_t part[t][c], c < C: 32 warps stride over the partials, then a fixed-order sum over the warps
__global__ void __launch_bounds__(1024) head_reduce_kernel(const float* __restrict__ part, int tiles, int C,
                                                           float* __restrict__ out0, int C0,
                                                           float* __restrict__ out1) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sub = threadIdx.x >> 5;
  __shared__ float s[32][33];
  float acc = 0.f;
  if (c < C)
    for (int t = sub; t < tiles; t += 32) acc += part[(int64_t)t * C + c];
  s[sub][threadIdx.x & 31] = acc;
  __syncthreads();
  if (sub == 0 && c < C) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) r += s[k][threadIdx.x];
    if (c < C0) {
      if (out0) out0[c] = r;
    } else if (out1) {
      out1[c - C0] = r;  // trailing columns (e.g. the bias gradient)
    }
  }
}

// ---- rowdot ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kHdWarps * 32)
rowdot_fwd_kernel(const float* __restrict__ h, int64_t hs, const float* __restrict__ w, int64_t B, int H, int vec,
                  float* __restrict__ y) {
  const int64_t b = (int64_t)blockIdx.x * kHdWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const float* row = h + b * hs;
  float acc = 0.f;
  if (vec) {
    for (int q = lane; q < H / 4; q += 32) {
      const float4 a = ldg_stream_f4(row + q * 4);
      const float4 c = *reinterpret_cast<const float4*>(w + q * 4);
      acc += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
    }
  } else {
    for (int k = lane; k < H; k += 32) acc += row[k] * w[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[b] = acc;
}

// g_h[b,k] = g[b]*w[k];  per-CTA partials of g_w[k] = sum_b g[b]*h[b,k]  (H % 4 == 0, H <= 128*MAXC, aligned)
// H2 (the K6 fused tower's hand-off, tc_linear.cu): h is the ReLU output of the tower's last layer, so the gradient of
// its pre-activation is g (x) w where h > 0; that product leaves as two fp16 planes split with *h2_scale (no fp32
// g_h, no split pass), its column sums (the last bias gradient) as per-CTA partials, its |max| into *h2_max.
template <int MAXC, bool H2>
__global__ void __launch_bounds__(kHdWarps * 32)
rowdot_bwd_kernel(const float* __restrict__ h, int64_t hs, const float* __restrict__ w, const float* __restrict__ g,
                  int64_t B, int H, float* __restrict__ gh, int64_t ghs, float* __restrict__ part,
                  unsigned short* __restrict__ planes, int64_t pl_ld, int64_t pl_plane,
                  const float* __restrict__ h2_scale, uint32_t* __restrict__ h2_max, float* __restrict__ cs_part) {
  extern __shared__ float s_acc[];  // [kHdWarps][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = H / 4;
  // blockIdx.y: a block of 32 * MAXC column quads (a row wider than 128 * MAXC columns is walked by several CTAs: with one
  // warp per whole 1024-column row the kernel held 4 x 32 accumulators per thread and ran at a quarter of the HBM rate)
  const int qbase = blockIdx.y * 32 * MAXC;
  const int k_lo = qbase * 4, k_hi = min(H, (qbase + 32 * MAXC) * 4);
  float4 a[MAXC], wv[MAXC], cs[H2 ? MAXC : 1];
  const float hsc = H2 ? *h2_scale : 1.f;
  float amax = 0.f;
#pragma unroll
  for (int u = 0; u < MAXC; ++u) {
    a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (H2) cs[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int q = qbase + lane + u * 32;
    wv[u] = q < nq ? *reinterpret_cast<const float4*>(w + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t b = (int64_t)blockIdx.x * kHdWarps + warp; b < B; b += (int64_t)gridDim.x * kHdWarps) {
    const float gb = g[b];
    const float* row = h + b * hs;
    float* orow = gh ? gh + b * ghs : nullptr;
    float4 t[MAXC];
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      const int q = qbase + lane + u * 32;
      t[u] = q < nq ? ldg_stream_f4(row + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      const int q = qbase + lane + u * 32;
      a[u].x += gb * t[u].x; a[u].y += gb * t[u].y; a[u].z += gb * t[u].z; a[u].w += gb * t[u].w;
      if (orow && q < nq) st_f4(orow + q * 4, make_float4(gb * wv[u].x, gb * wv[u].y, gb * wv[u].z, gb * wv[u].w));
      if (H2 && q < nq) {
        const float tv[4] = {t[u].x > 0.f ? gb * wv[u].x : 0.f, t[u].y > 0.f ? gb * wv[u].y : 0.f,
                             t[u].z > 0.f ? gb * wv[u].z : 0.f, t[u].w > 0.f ? gb * wv[u].w : 0.f};
        uint32_t p0[2], p1[2];
        split2h_pair(tv[0], tv[1], hsc, p0[0], p1[0]);
        split2h_pair(tv[2], tv[3], hsc, p0[1], p1[1]);
#pragma unroll
        for (int i = 0; i < 4; ++i) amax = fmaxf(amax, fabsf(tv[i]));
        unsigned short* o = planes + b * pl_ld + q * 4;
        *reinterpret_cast<uint2*>(o) = make_uint2(p0[0], p0[1]);
        *reinterpret_cast<uint2*>(o + pl_plane) = make_uint2(p1[0], p1[1]);
        cs[u].x += tv[0]; cs[u].y += tv[1]; cs[u].z += tv[2]; cs[u].w += tv[3];
      }
    }
  }
  if (H2) {
    // (pad columns [H, pl_ld) are never read: the GEMMs' tensor maps end at H)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0 && amax > 0.f && h2_max != nullptr) atomicMax(h2_max, __float_as_uint(amax));
    if (cs_part != nullptr) {  // fixed-order per-CTA column sums of the masked gradient (bias gradient)
#pragma unroll
      for (int u = 0; u < MAXC; ++u) {
        const int q = qbase + lane + u * 32;
        if (q < nq) *reinterpret_cast<float4*>(&s_acc[(size_t)warp * H + q * 4]) = cs[u];
      }
      __syncthreads();
      for (int k = k_lo + threadIdx.x; k < k_hi; k += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int wi = 0; wi < kHdWarps; ++wi) s += s_acc[(size_t)wi * H + k];
        cs_part[(int64_t)blockIdx.x * H + k] = s;
      }
      __syncthreads();
    }
  }
  if (part == nullptr) return;
#pragma unroll
  for (int u = 0; u < MAXC; ++u) {
    const int q = qbase + lane + u * 32;
    if (q < nq) *reinterpret_cast<float4*>(&s_acc[(size_t)warp * H + q * 4]) = a[u];
  }
  __syncthreads();
  for (int k = k_lo + threadIdx.x; k < k_hi; k += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int wi = 0; wi < kHdWarps; ++wi) s += s_acc[(size_t)wi * H + k];
    part[(int64_t)blockIdx.x * H + k] = s;
  }
}

static int head_grid() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms * 6;
}

}  // namespace ptrec

using namespace ptrec;

static bool head_vec_ok(int D, int F) {
  return D >= 4 && D <= 128 && (D & (D - 1)) == 0 && (int64_t)F * D <= 2048;
}

extern "C" int ptrec_fm_head_supported(int32_t F, int32_t D, int32_t nd) {
  return head_vec_ok(D, F) && nd >= 0 && nd <= kHdMaxDense ? 1 : 0;
}

extern "C" int ptrec_fm_head_fwd(const float* v, int64_t v_row_stride, const float* w1, int64_t w1_row_stride,
                                 const float* x, int64_t x_row_stride, const float* wd, const float* bias, int64_t B,
                                 int32_t F, int32_t D, int32_t nd, float* logit, float* deep_in,
                                 int64_t deep_in_row_stride, void* deep_in_planes, int64_t deep_in_planes_ld,
                                 void* stream) {
  PTREC_CHECK_ARG(v && logit && B >= 0 && F >= 1, PTREC_EINVAL, "fm_head_fwd: bad argument");
  PTREC_CHECK_ARG(ptrec_fm_head_supported(F, D, nd), PTREC_EUNSUPPORTED,
                  "fm_head_fwd: needs D a power of two in [4, 128], F*D <= 2048, <= %d dense features", kHdMaxDense);
  PTREC_CHECK_ARG(aligned16(v) && v_row_stride % 4 == 0 && v_row_stride >= (int64_t)F * D, PTREC_EALIGN,
                  "fm_head_fwd: v misaligned");
  PTREC_CHECK_ARG(!deep_in || (aligned16(deep_in) && deep_in_row_stride % 4 == 0 &&
                               deep_in_row_stride >= (int64_t)F * D + nd),
                  PTREC_EALIGN, "fm_head_fwd: deep_in misaligned");
  PTREC_CHECK_ARG(nd == 0 || x != nullptr, PTREC_EINVAL, "fm_head_fwd: dense features without x");
  PTREC_CHECK_ARG(!deep_in_planes || (aligned16(deep_in_planes) && deep_in_planes_ld % 8 == 0 &&
                                      deep_in_planes_ld >= (int64_t)F * D + nd),
                  PTREC_EALIGN, "fm_head_fwd: plane pitch must be a multiple of 8 >= F*D + nd");
  if (B == 0) return PTREC_OK;
  fm_head_fwd_kernel<<<(unsigned)ceil_div(B, kHdWarps), kHdWarps * 32, 0, (cudaStream_t)stream>>>(
      v, v_row_stride, w1, w1_row_stride, nd ? x : nullptr, x_row_stride, wd, bias, B, F, F * D / 4, D / 4, nd, logit,
      deep_in, deep_in_row_stride, reinterpret_cast<__nv_bfloat16*>(deep_in_planes), deep_in_planes_ld,
      B * deep_in_planes_ld, nullptr, nullptr);
  PTREC_LAUNCH_CHECK("fm_head_fwd_kernel");
  return PTREC_OK;
}

// the same pass writing the tower input as the fp16 x 2 planes of the K6 fused tower (carried scale, tc_linear.cu);
// deep_in (fp32) may be NULL when the tower reads the planes only
extern "C" int ptrec_fm_head_fwd_h2(const float* v, int64_t v_row_stride, const float* w1, int64_t w1_row_stride,
                                    const float* x, int64_t x_row_stride, const float* wd, const float* bias, int64_t B,
                                    int32_t F, int32_t D, int32_t nd, float* logit, float* deep_in,
                                    int64_t deep_in_row_stride, void* planes, int64_t planes_ld, const float* scale,
                                    float* max_out, void* stream) {
  PTREC_CHECK_ARG(v && logit && planes && scale && B >= 0 && F >= 1, PTREC_EINVAL, "fm_head_fwd_h2: bad argument");
  PTREC_CHECK_ARG(ptrec_fm_head_supported(F, D, nd), PTREC_EUNSUPPORTED, "fm_head_fwd_h2: unsupported shape");
  PTREC_CHECK_ARG(aligned16(v) && v_row_stride % 4 == 0 && v_row_stride >= (int64_t)F * D, PTREC_EALIGN,
                  "fm_head_fwd_h2: v misaligned");
  PTREC_CHECK_ARG(!deep_in || (aligned16(deep_in) && deep_in_row_stride % 4 == 0 &&
                               deep_in_row_stride >= (int64_t)F * D + nd),
                  PTREC_EALIGN, "fm_head_fwd_h2: deep_in misaligned");
  PTREC_CHECK_ARG(nd == 0 || x != nullptr, PTREC_EINVAL, "fm_head_fwd_h2: dense features without x");
  PTREC_CHECK_ARG(aligned16(planes) && planes_ld % 8 == 0 && planes_ld >= (int64_t)F * D + nd, PTREC_EALIGN,
                  "fm_head_fwd_h2: plane pitch must be a multiple of 8 >= F*D + nd");
  if (B == 0) return PTREC_OK;
  fm_head_fwd_kernel<<<(unsigned)ceil_div(B, kHdWarps), kHdWarps * 32, 0, (cudaStream_t)stream>>>(
      v, v_row_stride, w1, w1_row_stride, nd ? x : nullptr, x_row_stride, wd, bias, B, F, F * D / 4, D / 4, nd, logit,
      deep_in, deep_in_row_stride, reinterpret_cast<__nv_bfloat16*>(planes), planes_ld, B * planes_ld, scale,
      reinterpret_cast<uint32_t*>(max_out));
  PTREC_LAUNCH_CHECK("fm_head_fwd_kernel");
  return PTREC_OK;
}

extern "C" size_t ptrec_fm_head_bwd_workspace_bytes(int32_t nd) {
  return align_up((size_t)head_grid() * (nd + 1) * sizeof(float), 256);
}

extern "C" int ptrec_fm_head_bwd(const float* v, int64_t v_row_stride, const float* x, int64_t x_row_stride,
                                 const float* wd, const float* g, const float* g_deep_in, int64_t g_deep_in_row_stride,
                                 int64_t B, int32_t F, int32_t D, int32_t nd, float* grad_v, int64_t grad_v_row_stride,
                                 float* grad_w1, float* grad_x, int64_t grad_x_row_stride, float* grad_wd,
                                 float* grad_bias, void* workspace, size_t workspace_bytes, void* stream) {
  PTREC_CHECK_ARG(v && g && grad_v && B >= 0 && F >= 1, PTREC_EINVAL, "fm_head_bwd: bad argument");
  PTREC_CHECK_ARG(ptrec_fm_head_supported(F, D, nd), PTREC_EUNSUPPORTED, "fm_head_bwd: unsupported shape");
  PTREC_CHECK_ARG(aligned16(v) && v_row_stride % 4 == 0 && aligned16(grad_v) && grad_v_row_stride % 4 == 0 &&
                      (!g_deep_in || (aligned16(g_deep_in) && g_deep_in_row_stride % 4 == 0)),
                  PTREC_EALIGN, "fm_head_bwd: misaligned");
  const bool need_part = grad_wd != nullptr || grad_bias != nullptr;
  PTREC_CHECK_ARG(!need_part || (workspace && workspace_bytes >= ptrec_fm_head_bwd_workspace_bytes(nd)),
                  PTREC_EWORKSPACE, "fm_head_bwd: workspace too small");
  if (B == 0) return PTREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>(head_grid(), ceil_div(B, kHdWarps));
  float* part = need_part ? reinterpret_cast<float*>(workspace) : nullptr;
#define PTREC_HEAD_BWD(MC)                                                                                          \
  fm_head_bwd_kernel<MC><<<grid, kHdWarps * 32, 0, st>>>(v, v_row_stride, nd ? x : nullptr, x_row_stride, wd, g,     \
                                                         g_deep_in, g_deep_in_row_stride, B, F, F * D / 4, D / 4, nd, \
                                                         grad_v, grad_v_row_stride, grad_w1, grad_x, grad_x_row_stride, part)
  const int n_chunks = F * D / 4;
  if (n_chunks <= 128) PTREC_HEAD_BWD(4);
  else if (n_chunks <= 256) PTREC_HEAD_BWD(8);
  else PTREC_HEAD_BWD(16);
#undef PTREC_HEAD_BWD
  PTREC_LAUNCH_CHECK("fm_head_bwd_kernel");
  if (need_part) {
    cudaStream_t rs = reduce_stream_after(st);
    head_reduce_kernel<<<(unsigned)ceil_div(nd + 1, 32), 1024, 0, rs>>>(part, grid, nd + 1, grad_wd, nd, grad_bias);
    PTREC_LAUNCH_CHECK("head_reduce_kernel");
  }
  return PTREC_OK;
}

extern "C" int ptrec_rowdot_fwd(const float* h, int64_t h_row_stride, const float* w, int64_t B, int32_t H, float* y,
                                void* stream) {
  PTREC_CHECK_ARG(h && w && y && B >= 0 && H >= 1 && h_row_stride >= H, PTREC_EINVAL, "rowdot_fwd: bad argument");
  if (B == 0) return PTREC_OK;
  const int vec = (H % 4 == 0 && aligned16(h) && aligned16(w) && h_row_stride % 4 == 0) ? 1 : 0;
  rowdot_fwd_kernel<<<(unsigned)ceil_div(B, kHdWarps), kHdWarps * 32, 0, (cudaStream_t)stream>>>(h, h_row_stride, w, B,
                                                                                                 H, vec, y);
  PTREC_LAUNCH_CHECK("rowdot_fwd_kernel");
  return PTREC_OK;
}

extern "C" int ptrec_rowdot_supported(int32_t H) { return H % 4 == 0 && H >= 4 && H <= kHdMaxH ? 1 : 0; }

extern "C" size_t ptrec_rowdot_bwd_workspace_bytes(int32_t H) {
  return align_up((size_t)head_grid() * H * sizeof(float), 256);
}

extern "C" int ptrec_rowdot_bwd(const float* h, int64_t h_row_stride, const float* w, const float* g, int64_t B,
                                int32_t H, float* grad_h, int64_t grad_h_row_stride, float* grad_w, void* workspace,
                                size_t workspace_bytes, void* stream) {
  PTREC_CHECK_ARG(h && w && g && B >= 0, PTREC_EINVAL, "rowdot_bwd: bad argument");
  PTREC_CHECK_ARG(ptrec_rowdot_supported(H), PTREC_EUNSUPPORTED, "rowdot_bwd: H must be a multiple of 4 <= %d", kHdMaxH);
  PTREC_CHECK_ARG(aligned16(h) && aligned16(w) && h_row_stride % 4 == 0 &&
                      (!grad_h || (aligned16(grad_h) && grad_h_row_stride % 4 == 0)),
                  PTREC_EALIGN, "rowdot_bwd: misaligned");
  PTREC_CHECK_ARG(!grad_w || (workspace && workspace_bytes >= ptrec_rowdot_bwd_workspace_bytes(H)), PTREC_EWORKSPACE,
                  "rowdot_bwd: workspace too small");
  if (B == 0) return PTREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>(head_grid(), ceil_div(B, kHdWarps));
  const size_t smem = (size_t)kHdWarps * H * sizeof(float);
  float* part = grad_w ? reinterpret_cast<float*>(workspace) : nullptr;
  const dim3 grid2((unsigned)grid, (unsigned)ceil_div(H, 512));
  rowdot_bwd_kernel<4, false><<<grid2, kHdWarps * 32, smem, st>>>(h, h_row_stride, w, g, B, H, grad_h, grad_h_row_stride,
                                                                  part, nullptr, 0, 0, nullptr, nullptr, nullptr);
  PTREC_LAUNCH_CHECK("rowdot_bwd_kernel");
  if (grad_w) {
    head_reduce_kernel<<<(unsigned)ceil_div(H, 32), 1024, 0, st>>>(part, grid, H, grad_w, H, nullptr);
    PTREC_LAUNCH_CHECK("head_reduce_kernel");
  }
  return PTREC_OK;
}

// rowdot backward handing the masked gradient to the K6 fused tower as planes (see rowdot_bwd_kernel<., true>):
// workspace = 2 x ptrec_rowdot_bwd_workspace_bytes(H) (g_w partials, column-sum partials)
extern "C" int ptrec_rowdot_bwd_h2(const float* h, int64_t h_row_stride, const float* w, const float* g, int64_t B,
                                   int32_t H, void* planes, int64_t planes_ld, const float* scale, float* max_out,
                                   float* colsum, float* grad_w, void* workspace, size_t workspace_bytes, void* stream) {
  PTREC_CHECK_ARG(h && w && g && planes && scale && B >= 0, PTREC_EINVAL, "rowdot_bwd_h2: bad argument");
  PTREC_CHECK_ARG(ptrec_rowdot_supported(H), PTREC_EUNSUPPORTED, "rowdot_bwd_h2: H must be a multiple of 4 <= %d", kHdMaxH);
  PTREC_CHECK_ARG(aligned16(h) && aligned16(w) && h_row_stride % 4 == 0 && aligned16(planes) && planes_ld % 8 == 0 &&
                      planes_ld >= H, PTREC_EALIGN, "rowdot_bwd_h2: misaligned");
  const size_t one = ptrec_rowdot_bwd_workspace_bytes(H);
  PTREC_CHECK_ARG(workspace && workspace_bytes >= 2 * one, PTREC_EWORKSPACE, "rowdot_bwd_h2: workspace too small");
  if (B == 0) return PTREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>(head_grid(), ceil_div(B, kHdWarps));
  const size_t smem = (size_t)kHdWarps * H * sizeof(float);
  float* part = grad_w ? reinterpret_cast<float*>(workspace) : nullptr;
  float* cs_part = colsum ? reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + one) : nullptr;
  unsigned short* pl = reinterpret_cast<unsigned short*>(planes);
  const dim3 grid2((unsigned)grid, (unsigned)ceil_div(H, 512));
  rowdot_bwd_kernel<4, true><<<grid2, kHdWarps * 32, smem, st>>>(h, h_row_stride, w, g, B, H, nullptr, 0, part, pl,
                                                                 planes_ld, B * planes_ld, scale,
                                                                 reinterpret_cast<uint32_t*>(max_out), cs_part);
  PTREC_LAUNCH_CHECK("rowdot_bwd_kernel");
  cudaStream_t rs = (grad_w || colsum) ? reduce_stream_after(st) : st;
  if (grad_w) {
    head_reduce_kernel<<<(unsigned)ceil_div(H, 32), 1024, 0, rs>>>(part, grid, H, grad_w, H, nullptr);
    PTREC_LAUNCH_CHECK("head_reduce_kernel");
  }
  if (colsum) {
    head_reduce_kernel<<<(unsigned)ceil_div(H, 32), 1024, 0, rs>>>(cs_part, grid, H, colsum, H, nullptr);
    PTREC_LAUNCH_CHECK("head_reduce_kernel");
  }
  return PTREC_OK;
}
