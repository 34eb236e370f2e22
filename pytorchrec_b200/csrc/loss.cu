// K9: BCEWithLogitsLoss(reduction='mean') forward + gradient in one pass (the loss at the end of the reference's
// train_step, torchrec/model/IModel.py:121: `loss = self.compiled_loss(prediction, target)`; ATen spends ~8 launches on
// it per step: the stable formula's element-wise pieces, the mean, and their backward).
//   l_i = max(x_i, 0) - x_i t_i + log1p(exp(-|x_i|))        (= ATen's (1 - t) x + max(-x, 0) + log(exp(-m) + exp(-x - m)))
//   loss = (1 / n) sum_i l_i          grad_i = (sigmoid(x_i) - t_i) / n
// Deterministic: fixed per-thread strides, fixed tree per CTA, per-CTA partials summed in CTA order by the last CTA
// to finish (ticket counter in the caller's workspace, left at zero).  Bound: launch latency (n = batch).
#include "common.cuh"

namespace ptrec {

constexpr int kLossThreads = 1024;
constexpr int kLossMaxCtas = 64;

__global__ void __launch_bounds__(kLossThreads)
bce_logits_mean_kernel(const float* __restrict__ x, const float* __restrict__ t, int64_t n, float inv_n,
                       float* __restrict__ loss, float* __restrict__ grad, float* __restrict__ part,
                       unsigned int* __restrict__ ticket) {
  __shared__ float s_red[kLossThreads / 32];
  __shared__ bool s_last;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kLossThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kLossThreads) {
    const float xi = x[i], ti = t[i];
    const float e = expf(-fabsf(xi));                   // in (0, 1]
    acc += fmaxf(xi, 0.f) - xi * ti + log1pf(e);
    // sigmoid(x) - t without cancellation when the logit saturates towards its label: 1 - sigmoid(x) = e / (1 + e), x >= 0
    const float q = e / (1.f + e);
    if (grad != nullptr) grad[i] = (xi >= 0.f ? (1.f - ti) - q : q - ti) * inv_n;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = s_red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) {
      part[blockIdx.x] = v;
      __threadfence();
      s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    float tot = 0.f;
    for (unsigned c = 0; c < gridDim.x; ++c) tot += reinterpret_cast<volatile float*>(part)[c];  // CTA order
    *loss = tot * inv_n;
    *ticket = 0u;
  }
}

}  // namespace ptrec

using namespace ptrec;

extern "C" size_t ptrec_bce_logits_workspace_bytes(void) { return 512; }  // 64 partials + the ticket (zero on first use)

extern "C" int ptrec_bce_logits_mean(const float* logits, const float* target, int64_t n, float* loss, float* grad,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  PTREC_CHECK_ARG(logits && target && loss && n >= 1, PTREC_EINVAL, "bce_logits_mean: bad argument");
  PTREC_CHECK_ARG(workspace && workspace_bytes >= ptrec_bce_logits_workspace_bytes(), PTREC_EWORKSPACE,
                  "bce_logits_mean: workspace too small");
  float* part = reinterpret_cast<float*>(workspace);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(part + kLossMaxCtas);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(kLossMaxCtas, ceil_div(n, (int64_t)4 * kLossThreads)));
  bce_logits_mean_kernel<<<grid, kLossThreads, 0, (cudaStream_t)stream>>>(logits, target, n, 1.0f / (float)n, loss, grad,
                                                                           part, ticket);
  PTREC_LAUNCH_CHECK("bce_logits_mean_kernel");
  return PTREC_OK;
}
