// K4: DIN attention pooling (activation unit + masked weighted sum), forward and backward, fp32.
//
//   a_l    = W3 relu(W2 relu(W1 [q, k_l, q-k_l, q*k_l] + b1) + b2) + b3        (no softmax), l < len_b
//   pooled = sum_{l < len_b} a_l k_l
// Algebraic fusion: W1 [q,k,q-k,q*k] = (W1q+W1d) q + ((W1k-W1d) + W1p diag(q)) k = c_b + M_b k, so the first
// layer costs H1*DQ MACs per position instead of 4*H1*DQ, and nothing of shape [B, L, 4*DQ] or [B, L, H1] is
// ever written to HBM.  Algorithmic bytes: B*L*DQ*4 (+ B*L*DQ*4 key gradients in the backward);
// FLOPs fwd: B*L*2*(H1*DQ + H2*H1 + H2).  On fp32 CUDA cores the unit is COMPUTE-bound: reported
// against both the HBM roofline and the fp32-SIMT pipe.
// A persistent CTA keeps the weights in shared memory and walks samples; one thread owns one history position.
// The backward recomputes the unit, emits key / query gradients, and accumulates the weight gradients of all its
// samples in registers (4x4 register tiles of two small GEMMs per sample); per-CTA partials are reduced in a
// fixed order by a second kernel, so results are run-to-run bit-reproducible (no floating-point atomics).
#include "din_keys.cuh"
#include "common.cuh"

namespace ptrec {

constexpr int kDinPos = 128;  // history positions per pass (one thread each)

template <int DQ, int H1, int H2>
struct DinWeights {
  alignas(16) float Wkd[H1][DQ];  // W1k - W1d
  alignas(16) float W1p[H1][DQ];
  alignas(16) float Wq[H1][DQ];   // W1q + W1d
  alignas(16) float b1[H1];
  alignas(16) float W2t[H1][H2];  // W2 transposed
  alignas(16) float b2[H2];
  alignas(16) float W3[H2];
  alignas(16) float q[DQ];
  alignas(16) float c[H1];
  alignas(16) float M[H1][DQ];
  float b3;
};

template <int DQ, int H1, int H2>
__device__ __forceinline__ void din_load_weights(DinWeights<DQ, H1, H2>* s, const float* W1, const float* b1,
                                                 const float* W2, const float* b2, const float* W3, const float* b3) {
  for (int e = threadIdx.x; e < H1 * DQ; e += blockDim.x) {
    const int j = e / DQ, i = e - j * DQ;
    const float* row = W1 + (int64_t)j * 4 * DQ;
    const float wq = row[i], wk = row[DQ + i], wd = row[2 * DQ + i], wp = row[3 * DQ + i];
    s->Wkd[j][i] = wk - wd;
    s->W1p[j][i] = wp;
    s->Wq[j][i] = wq + wd;
  }
  for (int e = threadIdx.x; e < H1 * H2; e += blockDim.x) {
    const int m = e / H1, j = e - m * H1;
    s->W2t[j][m] = W2[e];
  }
  for (int e = threadIdx.x; e < H1; e += blockDim.x) s->b1[e] = b1[e];
  for (int e = threadIdx.x; e < H2; e += blockDim.x) {
    s->b2[e] = b2[e];
    s->W3[e] = W3[e];
  }
  if (threadIdx.x == 0) s->b3 = b3[0];
}

// per-sample constants: c = Wq q + b1, M = Wkd + W1p diag(q)   (call between two __syncthreads)
template <int DQ, int H1, int H2>
__device__ __forceinline__ void din_sample_setup(DinWeights<DQ, H1, H2>* s) {
  for (int j = threadIdx.x; j < H1; j += blockDim.x) {
    float acc = s->b1[j];
#pragma unroll
    for (int i = 0; i < DQ; ++i) acc += s->Wq[j][i] * s->q[i];
    s->c[j] = acc;
  }
  for (int e = threadIdx.x; e < H1 * DQ; e += blockDim.x) {
    const int j = e / DQ, i = e - j * DQ;
    s->M[j][i] = s->Wkd[j][i] + s->W1p[j][i] * s->q[i];
  }
}

// activation unit for one position: returns a, leaves pre-activations of layer 2 in h2pre; optionally stores relu(h1)
template <int DQ, int H1, int H2, bool STORE_H1>
__device__ __forceinline__ float din_unit(const DinWeights<DQ, H1, H2>* s, const float* kk, float* h2pre, float* h1_row) {
#pragma unroll
  for (int m = 0; m < H2; ++m) h2pre[m] = s->b2[m];
#pragma unroll 2
  for (int j = 0; j < H1; ++j) {
    float h = s->c[j];
#pragma unroll
    for (int i = 0; i < DQ; i += 4) {
      const float4 w = *reinterpret_cast<const float4*>(&s->M[j][i]);
      h += w.x * kk[i] + w.y * kk[i + 1] + w.z * kk[i + 2] + w.w * kk[i + 3];
    }
    h = fmaxf(h, 0.f);
    if (STORE_H1) h1_row[j] = h;
#pragma unroll
    for (int m = 0; m < H2; m += 4) {
      const float4 w = *reinterpret_cast<const float4*>(&s->W2t[j][m]);
      h2pre[m] += w.x * h;
      h2pre[m + 1] += w.y * h;
      h2pre[m + 2] += w.z * h;
      h2pre[m + 3] += w.w * h;
    }
  }
  float a = s->b3;
#pragma unroll
  for (int m = 0; m < H2; ++m) a += s->W3[m] * fmaxf(h2pre[m], 0.f);
  return a;
}

template <int DQ, int H1, int H2>
struct DinFwdSmem {
  DinWeights<DQ, H1, H2> w;
  float k[kDinPos][DQ + 1];
  float a[kDinPos];
  float red[kDinPos / DQ][DQ];
};

template <int DQ, int H1, int H2>
__global__ void __launch_bounds__(kDinPos)
din_fwd_kernel(const float* __restrict__ q, int64_t q_stride, const float* __restrict__ keys, int64_t ksb, int64_t ksl,
               const int32_t* __restrict__ lens, int64_t B, int L, const float* W1, const float* b1, const float* W2,
               const float* b2, const float* W3, const float* b3, float* __restrict__ out, float* __restrict__ scores) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto* s = reinterpret_cast<DinFwdSmem<DQ, H1, H2>*>(smem_raw);
  constexpr int PARTS = kDinPos / DQ;
  const int t = threadIdx.x;
  din_load_weights(&s->w, W1, b1, W2, b2, W3, b3);
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const int len = lens ? min(max(lens[b], 0), L) : L;
    __syncthreads();
    if (t < DQ) s->w.q[t] = q[b * q_stride + t];
    __syncthreads();
    din_sample_setup(&s->w);
    __syncthreads();
    const int i = t % DQ, part = t / DQ;
    float acc = 0.f;
    for (int l0 = 0; l0 < len; l0 += kDinPos) {
      const int l = l0 + t;
      float a = 0.f;
      if (l < len) {
        float kk[DQ];
        const float* kp = keys + b * ksb + (int64_t)l * ksl;
#pragma unroll
        for (int x = 0; x < DQ; x += 4) {
          const float4 v = ldg_stream_f4(kp + x);
          kk[x] = v.x; kk[x + 1] = v.y; kk[x + 2] = v.z; kk[x + 3] = v.w;
        }
#pragma unroll
        for (int x = 0; x < DQ; ++x) s->k[t][x] = kk[x];
        float h2pre[H2];
        a = din_unit<DQ, H1, H2, false>(&s->w, kk, h2pre, nullptr);
        if (scores) scores[b * L + l] = a;
      }
      s->a[t] = a;
      __syncthreads();
      const int n = min(kDinPos, len - l0);
      for (int ll = part; ll < n; ll += PARTS) acc += s->a[ll] * s->k[ll][i];
      __syncthreads();
    }
    if (scores) {
      for (int l = len + t; l < L; l += kDinPos) scores[b * L + l] = 0.f;
    }
    s->red[part][i] = acc;
    __syncthreads();
    if (t < DQ) {
      float r = 0.f;
#pragma unroll
      for (int p = 0; p < PARTS; ++p) r += s->red[p][t];
      out[b * DQ + t] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int kDinBwdThreads = 256;
constexpr int kDinBwdGroup = 8;  // samples per CTA of the backward

template <int DQ, int H1, int H2>
struct DinBwdSmem {
  DinWeights<DQ, H1, H2> w;
  float gp[DQ];                    // upstream gradient of pooled
  float h1[kDinPos][H1 + 1];       // relu(h1) per position
  float gh1[kDinPos][H1 + 1];      // d loss / d h1pre
  float gh2[kDinPos][H2 + 1];      // d loss / d h2pre
  float gah2[kDinPos][H2 + 1];     // g_a * relu(h2)  (contribution to grad W3)
  float k[kDinPos][DQ + 1];
  float ga[kDinPos];
  float gqp[H1 / 4][DQ];           // per M-block-row partial of grad q
};

template <int DQ, int H1, int H2>
constexpr int din_grad_floats() { return H1 * 4 * DQ + H1 + H2 * H1 + H2 + H2 + 1; }

template <int DQ, int H1, int H2>
__global__ void __launch_bounds__(kDinBwdThreads)
din_bwd_kernel(const float* __restrict__ q, int64_t q_stride, const float* __restrict__ keys, int64_t ksb, int64_t ksl,
               const int32_t* __restrict__ lens, int64_t B, int L, const float* W1, const float* b1, const float* W2,
               const float* b2, const float* W3, const float* b3, const float* __restrict__ g_pooled,
               float* __restrict__ g_q, float* __restrict__ g_keys, int64_t gksb, int64_t gksl,
               float* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto* s = reinterpret_cast<DinBwdSmem<DQ, H1, H2>*>(smem_raw);
  constexpr int MB_J = H1 / 4, MB_I = DQ / 4, N_MBLK = MB_J * MB_I;     // grad-M register tiles (4 x 4)
  constexpr int WB_M = H2 / 4, WB_J = H1 / 4, N_WBLK = WB_M * WB_J;     // grad-W2 register tiles
  static_assert(N_MBLK <= kDinBwdThreads && N_WBLK <= kDinBwdThreads, "activation unit too wide for one CTA");
  static_assert(H1 % 4 == 0 && H2 % 4 == 0 && DQ % 4 == 0 && kDinPos % DQ == 0, "dims must be multiples of 4");
  const int t = threadIdx.x;
  din_load_weights(&s->w, W1, b1, W2, b2, W3, b3);

  // weight-gradient accumulators, summed over every sample this CTA walks
  float A1[4][4] = {}, A2[4][4] = {}, A3[4][4] = {};  // grad wrt Wkd, W1p, Wq tiles (M-block owners)
  float AW2[4][4] = {};                                 // grad W2 tile (W2-block owners)
  float Ab1[4] = {};                                    // grad b1 for the tile's 4 rows (M-block owners with ib == 0)
  float Ab2 = 0.f, AW3 = 0.f, Ab3 = 0.f;                // threads t < H2 (b2, W3), t == 0 (b3)
  const bool m_owner = t < N_MBLK, w_owner = t < N_WBLK;
  const int mjb = t / MB_I, mib = t % MB_I;             // M tile: rows 4*mjb.., cols 4*mib..
  const int wmb = t / WB_J, wjb = t % WB_J;             // W2 tile: rows (m) 4*wmb.., cols (j) 4*wjb..

  // A CTA owns kDinBwdGroup CONSECUTIVE samples: its weight-gradient accumulators see chains of at most L positions
  // (per-sample tiles) + kDinBwdGroup samples, and the per-CTA partials are summed two-level by the reduce kernel.
  // (One CTA per SM walking B / 148 samples kept chains of ~5500 fp32 additions: the resulting gradient noise, turned
  // into weight noise by Adagrad's g / sqrt(sum g^2), was 100x the CPU oracle's at the 99th percentile —
  // tests/test_gpu_fullsize.py with the fp64 oracle as referee.)
  const int64_t b_end = min(B, ((int64_t)blockIdx.x + 1) * kDinBwdGroup);
  for (int64_t b = (int64_t)blockIdx.x * kDinBwdGroup; b < b_end; ++b) {
    const int len = lens ? min(max(lens[b], 0), L) : L;
    __syncthreads();
    if (t < DQ) {
      s->w.q[t] = q[b * q_stride + t];
      s->gp[t] = g_pooled[b * DQ + t];
    }
    __syncthreads();
    din_sample_setup(&s->w);
    __syncthreads();
    float gq_tile[4] = {0.f, 0.f, 0.f, 0.f};  // this tile's contribution to grad q (cols 4*mib..), summed over chunks
    for (int l0 = 0; l0 < max(len, 1); l0 += kDinPos) {
      const int n = max(0, min(kDinPos, len - l0));
      // ---- phase 1: one thread per position ------------------------------------------------------------------
      if (t < kDinPos) {
        const int l = l0 + t;
        if (t < n) {
          float kk[DQ];
          const float* kp = keys + b * ksb + (int64_t)l * ksl;
#pragma unroll
          for (int x = 0; x < DQ; x += 4) {
            const float4 v = ldg_stream_f4(kp + x);
            kk[x] = v.x; kk[x + 1] = v.y; kk[x + 2] = v.z; kk[x + 3] = v.w;
          }
          float h2pre[H2];
          const float a = din_unit<DQ, H1, H2, true>(&s->w, kk, h2pre, s->h1[t]);
          float ga = 0.f;
#pragma unroll
          for (int x = 0; x < DQ; ++x) {
            ga += s->gp[x] * kk[x];
            s->k[t][x] = kk[x];
          }
          s->ga[t] = ga;
          float gh2[H2];
#pragma unroll
          for (int m = 0; m < H2; ++m) {
            gh2[m] = h2pre[m] > 0.f ? ga * s->w.W3[m] : 0.f;
            s->gh2[t][m] = gh2[m];
            s->gah2[t][m] = ga * fmaxf(h2pre[m], 0.f);
          }
          float gk[DQ];
#pragma unroll
          for (int x = 0; x < DQ; ++x) gk[x] = a * s->gp[x];
#pragma unroll 2
          for (int j = 0; j < H1; ++j) {
            float sum = 0.f;
#pragma unroll
            for (int m = 0; m < H2; m += 4) {
              const float4 w = *reinterpret_cast<const float4*>(&s->w.W2t[j][m]);
              sum += w.x * gh2[m] + w.y * gh2[m + 1] + w.z * gh2[m + 2] + w.w * gh2[m + 3];
            }
            const float g1 = s->h1[t][j] > 0.f ? sum : 0.f;
            s->gh1[t][j] = g1;
#pragma unroll
            for (int x = 0; x < DQ; x += 4) {
              const float4 w = *reinterpret_cast<const float4*>(&s->w.M[j][x]);
              gk[x] += w.x * g1; gk[x + 1] += w.y * g1; gk[x + 2] += w.z * g1; gk[x + 3] += w.w * g1;
            }
          }
          float* gkp = g_keys + b * gksb + (int64_t)l * gksl;
#pragma unroll
          for (int x = 0; x < DQ; x += 4) st_f4(gkp + x, make_float4(gk[x], gk[x + 1], gk[x + 2], gk[x + 3]));
        }
      }
      __syncthreads();
      // ---- phase 2: per-sample weight-gradient GEMMs over the n valid positions (4x4 register tiles) -----------
      if (m_owner) {
        float gM[4][4] = {};
        float gc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int l = 0; l < n; ++l) {
          float g[4], kv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) g[r] = s->gh1[l][4 * mjb + r];
#pragma unroll
          for (int c = 0; c < 4; ++c) kv[c] = s->k[l][4 * mib + c];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            gc[r] += g[r];
#pragma unroll
            for (int c = 0; c < 4; ++c) gM[r][c] += g[r] * kv[c];
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (mib == 0) Ab1[r] += gc[r];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float qv = s->w.q[4 * mib + c];
            A1[r][c] += gM[r][c];
            A2[r][c] += gM[r][c] * qv;
            A3[r][c] += gc[r] * qv;
            // d c / d q = Wq ; d M / d q = W1p (column-wise)
            gq_tile[c] += s->w.Wq[4 * mjb + r][4 * mib + c] * gc[r] + s->w.W1p[4 * mjb + r][4 * mib + c] * gM[r][c];
          }
        }
      }
      if (w_owner) {
        for (int l = 0; l < n; ++l) {
          float g[4], h[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) g[r] = s->gh2[l][4 * wmb + r];
#pragma unroll
          for (int c = 0; c < 4; ++c) h[c] = s->h1[l][4 * wjb + c];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) AW2[r][c] += g[r] * h[c];
        }
      }
      if (t < H2) {
        for (int l = 0; l < n; ++l) {
          Ab2 += s->gh2[l][t];
          AW3 += s->gah2[l][t];
        }
      }
      if (t == kDinBwdThreads - 1) {
        for (int l = 0; l < n; ++l) Ab3 += s->ga[l];
      }
      __syncthreads();
    }
    // zero gradient for the padded tail of the history
    for (int64_t e = len * (int64_t)DQ + t; e < (int64_t)L * DQ; e += kDinBwdThreads) {
      const int64_t l = e / DQ;
      g_keys[b * gksb + l * gksl + (e - l * DQ)] = 0.f;
    }
    // grad q: sum the tile partials over the H1/4 tile rows (fixed order)
    if (m_owner) {
#pragma unroll
      for (int c = 0; c < 4; ++c) s->gqp[mjb][4 * mib + c] = gq_tile[c];
    }
    __syncthreads();
    if (t < DQ) {
      float r = 0.f;
#pragma unroll 4
      for (int jb = 0; jb < MB_J; ++jb) r += s->gqp[jb][t];
      g_q[b * DQ + t] = r;
    }
  }

  // per-CTA partial weight gradients: [W1 (H1 x 4DQ) | b1 | W2 (H2 x H1) | b2 | W3 | b3]
  float* P = partials + (int64_t)blockIdx.x * din_grad_floats<DQ, H1, H2>();
  if (m_owner) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float* row = P + (int64_t)(4 * mjb + r) * 4 * DQ;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = 4 * mib + c;
        row[i] = A3[r][c];                       // W1q
        row[DQ + i] = A1[r][c];                  // W1k
        row[2 * DQ + i] = A3[r][c] - A1[r][c];   // W1d  (Wq = W1q + W1d, Wkd = W1k - W1d)
        row[3 * DQ + i] = A2[r][c];              // W1p
      }
      if (mib == 0) P[H1 * 4 * DQ + 4 * mjb + r] = Ab1[r];
    }
  }
  float* PW2 = P + H1 * 4 * DQ + H1;
  if (w_owner) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) PW2[(4 * wmb + r) * H1 + 4 * wjb + c] = AW2[r][c];
  }
  if (t < H2) {
    PW2[H2 * H1 + t] = Ab2;
    PW2[H2 * H1 + H2 + t] = AW3;
  }
  if (t == kDinBwdThreads - 1) PW2[H2 * H1 + 2 * H2] = Ab3;
}

// out[e] = sum over CTAs of partials[cta][e], two fixed-order levels (runs of 32 rows, then the run sums): coalesced
// over e, chains of 32 + n_cta / 32 additions
__global__ void din_reduce_partials_kernel(const float* __restrict__ partials, int n_cta, int n, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float total = 0.f;
  for (int c0 = 0; c0 < n_cta; c0 += 32) {
    float acc = 0.f;
    const int c1 = min(n_cta, c0 + 32);
    for (int c = c0; c < c1; ++c) acc += partials[(int64_t)c * n + e];
    total += acc;
  }
  out[e] = total;
}

static int din_grid() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

template <int DQ, int H1, int H2>
static int din_fwd_launch(const float* q, int64_t qs, const float* keys, int64_t ksb, int64_t ksl, const int32_t* lens,
                          int64_t B, int L, const float* W1, const float* b1, const float* W2, const float* b2,
                          const float* W3, const float* b3, float* out, float* scores, cudaStream_t st) {
  const size_t smem = sizeof(DinFwdSmem<DQ, H1, H2>);
  PTREC_CUDA(cudaFuncSetAttribute(din_fwd_kernel<DQ, H1, H2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)(B < (int64_t)din_grid() * 2 ? B : (int64_t)din_grid() * 2);
  din_fwd_kernel<DQ, H1, H2><<<grid, kDinPos, smem, st>>>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3, out,
                                                           scores);
  PTREC_LAUNCH_CHECK("din_fwd_kernel");
  return PTREC_OK;
}

template <int DQ, int H1, int H2>
static int din_bwd_launch(const float* q, int64_t qs, const float* keys, int64_t ksb, int64_t ksl, const int32_t* lens,
                          int64_t B, int L, const float* W1, const float* b1, const float* W2, const float* b2,
                          const float* W3, const float* b3, const float* g_pooled, float* g_q, float* g_keys,
                          int64_t gksb, int64_t gksl, float* grads, float* partials, cudaStream_t st) {
  const size_t smem = sizeof(DinBwdSmem<DQ, H1, H2>);
  PTREC_CUDA(cudaFuncSetAttribute(din_bwd_kernel<DQ, H1, H2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)((B + kDinBwdGroup - 1) / kDinBwdGroup);
  din_bwd_kernel<DQ, H1, H2><<<grid, kDinBwdThreads, smem, st>>>(q, qs, keys, ksb, ksl, lens, B, L, W1, b1, W2, b2, W3, b3,
                                                                  g_pooled, g_q, g_keys, gksb, gksl, partials);
  PTREC_LAUNCH_CHECK("din_bwd_kernel");
  const int n = din_grad_floats<DQ, H1, H2>();
  din_reduce_partials_kernel<<<(n + 255) / 256, 256, 0, st>>>(partials, grid, n, grads);
  PTREC_LAUNCH_CHECK("din_reduce_partials_kernel");
  return PTREC_OK;
}

}  // namespace ptrec

namespace ptrec {
// din_attn_tc.cu: the forward on the tensor cores (PTREC_EUNSUPPORTED when the shape has no such build)
int din_fwd_tc(const float* q, int64_t qs, const float* keys, int64_t ksb, int64_t ksl, const int32_t* lens, int64_t B,
               int L, int DQ, int H1, int H2, const float* W1, const float* b1, const float* W2, const float* b2,
               const float* W3, const float* b3, float* out, float* scores, cudaStream_t st,
               const DinKeyIds* kid = nullptr);
int din_bwd_tc(const float* q, int64_t qs, const float* keys, int64_t ksb, int64_t ksl, const int32_t* lens, int64_t B,
               int L, int DQ, int H1, int H2, const float* W1, const float* b1, const float* W2, const float* b2,
               const float* W3, const float* b3, const float* g_pooled, float* g_q, float* g_keys, int64_t gksb,
               int64_t gksl, float* partials, int* n_rows, cudaStream_t st, const DinKeyIds* kid = nullptr);
}  // namespace ptrec

using namespace ptrec;

// bit 0: forward through the tcgen05 kernel where a build exists; bit 1: backward as well (default: both); 0: fp32 SIMT
static int g_din_tc = 3;
extern "C" void ptrec_set_din_tc(int32_t mode) { g_din_tc = mode & 3; }
extern "C" int32_t ptrec_din_tc_enabled(void) { return g_din_tc; }

static int din_check(const void* q, const void* keys, int64_t B, int32_t L, int32_t DQ, int32_t H1, int32_t H2,
                     int64_t qs, int64_t ksb, int64_t ksl) {
  PTREC_CHECK_ARG(q && keys && B >= 0 && L >= 1, PTREC_EINVAL, "din_attn_pool: bad argument");
  const bool dims_ok = (DQ == 16 || DQ == 32) && ((H1 == 80 && H2 == 40) || (H1 == 64 && H2 == 32));
  PTREC_CHECK_ARG(dims_ok, PTREC_EUNSUPPORTED,
                  "din_attn_pool: built for DQ in {16,32} and (H1,H2) in {(80,40),(64,32)}; got DQ=%d H1=%d H2=%d", DQ, H1, H2);
  PTREC_CHECK_ARG(aligned16(q) && aligned16(keys) && qs % 4 == 0 && ksb % 4 == 0 && ksl % 4 == 0, PTREC_EALIGN,
                  "din_attn_pool: q / keys must be 16-byte aligned with strides that are multiples of 4 floats");
  return PTREC_OK;
}

#define PTREC_DIN_DISPATCH(FN, ...)                                   \
  do {                                                                \
    if (H1 == 80) {                                                   \
      if (DQ == 16) return FN<16, 80, 40>(__VA_ARGS__);               \
      return FN<32, 80, 40>(__VA_ARGS__);                             \
    }                                                                 \
    if (DQ == 16) return FN<16, 64, 32>(__VA_ARGS__);                 \
    return FN<32, 64, 32>(__VA_ARGS__);                               \
  } while (0)

extern "C" int ptrec_din_attn_pool_fwd(const float* q, int64_t q_stride, const float* keys, int64_t k_stride_b,
                                       int64_t k_stride_l, const int32_t* lens, int64_t B, int32_t L, int32_t DQ,
                                       int32_t H1, int32_t H2, const float* W1, const float* b1, const float* W2,
                                       const float* b2, const float* W3, const float* b3, float* out, float* scores,
                                       void* stream) {
  int rc = din_check(q, keys, B, L, DQ, H1, H2, q_stride, k_stride_b, k_stride_l);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(W1 && b1 && W2 && b2 && W3 && b3 && out, PTREC_EINVAL, "din_attn_pool_fwd: null pointer");
  if (B == 0) return PTREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_din_tc & 1) {
    rc = din_fwd_tc(q, q_stride, keys, k_stride_b, k_stride_l, lens, B, L, DQ, H1, H2, W1, b1, W2, b2, W3, b3, out, scores,
                    st);
    if (rc != PTREC_EUNSUPPORTED) return rc;
  }
  PTREC_DIN_DISPATCH(din_fwd_launch, q, q_stride, keys, k_stride_b, k_stride_l, lens, B, L, W1, b1, W2, b2, W3, b3, out,
                     scores, st);
}

extern "C" int32_t ptrec_din_attn_pool_grad_floats(int32_t DQ, int32_t H1, int32_t H2) {
  return H1 * 4 * DQ + H1 + H2 * H1 + H2 + H2 + 1;
}

extern "C" size_t ptrec_din_attn_pool_bwd_workspace_bytes(int64_t B, int32_t DQ, int32_t H1, int32_t H2) {
  const size_t rows = (size_t)((B > 0 ? B : 1) + kDinBwdGroup - 1) / kDinBwdGroup;  // one partial row per CTA
  return align_up(rows * (size_t)ptrec_din_attn_pool_grad_floats(DQ, H1, H2) * sizeof(float), 256);
}

extern "C" int ptrec_din_attn_pool_bwd(const float* q, int64_t q_stride, const float* keys, int64_t k_stride_b,
                                       int64_t k_stride_l, const int32_t* lens, int64_t B, int32_t L, int32_t DQ,
                                       int32_t H1, int32_t H2, const float* W1, const float* b1, const float* W2,
                                       const float* b2, const float* W3, const float* b3, const float* g_pooled,
                                       float* g_q, float* g_keys, int64_t gk_stride_b, int64_t gk_stride_l,
                                       float* grad_params, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = din_check(q, keys, B, L, DQ, H1, H2, q_stride, k_stride_b, k_stride_l);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(W1 && b1 && W2 && b2 && W3 && b3 && g_pooled && g_q && g_keys && grad_params && workspace, PTREC_EINVAL,
                  "din_attn_pool_bwd: null pointer");
  PTREC_CHECK_ARG(aligned16(g_keys) && gk_stride_b % 4 == 0 && gk_stride_l % 4 == 0, PTREC_EALIGN, "din_attn_pool_bwd: g_keys alignment");
  PTREC_CHECK_ARG(workspace_bytes >= ptrec_din_attn_pool_bwd_workspace_bytes(B, DQ, H1, H2), PTREC_EWORKSPACE,
                  "din_attn_pool_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    PTREC_CUDA(cudaMemsetAsync(grad_params, 0, (size_t)ptrec_din_attn_pool_grad_floats(DQ, H1, H2) * sizeof(float), st));
    return PTREC_OK;
  }
  float* partials = reinterpret_cast<float*>(workspace);
  if (g_din_tc & 2) {
    int n_rows = 0;
    rc = din_bwd_tc(q, q_stride, keys, k_stride_b, k_stride_l, lens, B, L, DQ, H1, H2, W1, b1, W2, b2, W3, b3, g_pooled,
                    g_q, g_keys, gk_stride_b, gk_stride_l, partials, &n_rows, st);
    if (rc == PTREC_OK) {
      const int n = ptrec_din_attn_pool_grad_floats(DQ, H1, H2);
      din_reduce_partials_kernel<<<(n + 255) / 256, 256, 0, st>>>(partials, n_rows, n, grad_params);
      PTREC_LAUNCH_CHECK("din_reduce_partials_kernel");
      return PTREC_OK;
    }
    if (rc != PTREC_EUNSUPPORTED) return rc;
  }
  PTREC_DIN_DISPATCH(din_bwd_launch, q, q_stride, keys, k_stride_b, k_stride_l, lens, B, L, W1, b1, W2, b2, W3, b3,
                     g_pooled, g_q, g_keys, gk_stride_b, gk_stride_l, grad_params, partials, st);
}

// ---- K4 with the key gather fused in (tensor-core builds only; din_keys.cuh) ---------------------------------------
static int din_ids_fill(DinKeyIds* kid, const float* table0, const float* table1, int64_t row_stride0, int64_t row_stride1,
                        int64_t rows0, int64_t rows1, const int64_t* ids0, const int64_t* ids1, int64_t ids_stride_b,
                        int64_t ids_offset, int32_t* err_flag, int32_t DQ) {
  PTREC_CHECK_ARG(table0 && table1 && ids0 && ids1 && rows0 >= 1 && rows1 >= 1 && ids_stride_b >= 1 && ids_offset >= 0,
                  PTREC_EINVAL, "din_attn_pool_ids: bad argument");
  PTREC_CHECK_ARG(aligned16(table0) && aligned16(table1) && row_stride0 % 4 == 0 && row_stride1 % 4 == 0 &&
                      row_stride0 >= DQ / 2 && row_stride1 >= DQ / 2, PTREC_EALIGN,
                  "din_attn_pool_ids: tables must be 16-byte aligned fp32 with a row pitch that is a multiple of 4 floats");
  kid->base[0] = table0; kid->base[1] = table1;
  kid->stride[0] = row_stride0; kid->stride[1] = row_stride1;
  kid->rows[0] = rows0; kid->rows[1] = rows1;
  kid->ids[0] = ids0; kid->ids[1] = ids1;
  kid->ids_sb = ids_stride_b; kid->ids_off = ids_offset;
  kid->err = err_flag;
  return PTREC_OK;
}

extern "C" int ptrec_din_attn_pool_fwd_ids(const float* q, int64_t q_stride, const float* table0, const float* table1,
                                           int64_t row_stride0, int64_t row_stride1, int64_t rows0, int64_t rows1,
                                           const int64_t* ids0, const int64_t* ids1, int64_t ids_stride_b,
                                           int64_t ids_offset, int32_t* err_flag, const int32_t* lens, int64_t B, int32_t L,
                                           int32_t DQ, int32_t H1, int32_t H2, const float* W1, const float* b1,
                                           const float* W2, const float* b2, const float* W3, const float* b3, float* out,
                                           float* scores, void* stream) {
  int rc = din_check(q, table0, B, L, DQ, H1, H2, q_stride, 4, 4);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(W1 && b1 && W2 && b2 && W3 && b3 && out, PTREC_EINVAL, "din_attn_pool_fwd_ids: null pointer");
  if (!(g_din_tc & 1)) return PTREC_EUNSUPPORTED;
  DinKeyIds kid;
  rc = din_ids_fill(&kid, table0, table1, row_stride0, row_stride1, rows0, rows1, ids0, ids1, ids_stride_b, ids_offset,
                    err_flag, DQ);
  if (rc != PTREC_OK) return rc;
  if (B == 0) return PTREC_OK;
  return din_fwd_tc(q, q_stride, nullptr, 0, 0, lens, B, L, DQ, H1, H2, W1, b1, W2, b2, W3, b3, out, scores,
                    (cudaStream_t)stream, &kid);
}

extern "C" int ptrec_din_attn_pool_bwd_ids(const float* q, int64_t q_stride, const float* table0, const float* table1,
                                           int64_t row_stride0, int64_t row_stride1, int64_t rows0, int64_t rows1,
                                           const int64_t* ids0, const int64_t* ids1, int64_t ids_stride_b,
                                           int64_t ids_offset, const int32_t* lens, int64_t B, int32_t L, int32_t DQ,
                                           int32_t H1, int32_t H2, const float* W1, const float* b1, const float* W2,
                                           const float* b2, const float* W3, const float* b3, const float* g_pooled,
                                           float* g_q, float* g_keys, int64_t gk_stride_b, int64_t gk_stride_l,
                                           float* grad_params, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = din_check(q, table0, B, L, DQ, H1, H2, q_stride, 4, 4);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(W1 && b1 && W2 && b2 && W3 && b3 && g_pooled && g_q && g_keys && grad_params && workspace, PTREC_EINVAL,
                  "din_attn_pool_bwd_ids: null pointer");
  PTREC_CHECK_ARG(aligned16(g_keys) && gk_stride_b % 4 == 0 && gk_stride_l % 4 == 0, PTREC_EALIGN,
                  "din_attn_pool_bwd_ids: g_keys alignment");
  PTREC_CHECK_ARG(workspace_bytes >= ptrec_din_attn_pool_bwd_workspace_bytes(B, DQ, H1, H2), PTREC_EWORKSPACE,
                  "din_attn_pool_bwd_ids: workspace too small");
  if (!(g_din_tc & 2)) return PTREC_EUNSUPPORTED;
  DinKeyIds kid;
  rc = din_ids_fill(&kid, table0, table1, row_stride0, row_stride1, rows0, rows1, ids0, ids1, ids_stride_b, ids_offset,
                    nullptr, DQ);
  if (rc != PTREC_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    PTREC_CUDA(cudaMemsetAsync(grad_params, 0, (size_t)ptrec_din_attn_pool_grad_floats(DQ, H1, H2) * sizeof(float), st));
    return PTREC_OK;
  }
  float* partials = reinterpret_cast<float*>(workspace);
  int n_rows = 0;
  rc = din_bwd_tc(q, q_stride, nullptr, 0, 0, lens, B, L, DQ, H1, H2, W1, b1, W2, b2, W3, b3, g_pooled, g_q, g_keys,
                  gk_stride_b, gk_stride_l, partials, &n_rows, st, &kid);
  if (rc != PTREC_OK) return rc;
  const int n = ptrec_din_attn_pool_grad_floats(DQ, H1, H2);
  din_reduce_partials_kernel<<<(n + 255) / 256, 256, 0, st>>>(partials, n_rows, n, grad_params);
  PTREC_LAUNCH_CHECK("din_reduce_partials_kernel");
  return PTREC_OK;
}
