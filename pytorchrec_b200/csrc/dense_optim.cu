// K7: one-launch update of the (small) dense parameters that ride along with the fused table optimizers.
// The reference steps them with the stock dense optimizer (torchrec/optim/optimizers.py:7-11 -> torch.optim.*); the
// stock multi-tensor path costs ~10 latency-bound launches (~100 us at cfg2 for 0.5 M parameters).  Same arithmetic,
// in the order torch.optim uses, one launch for every tensor of a parameter group.   Bound: launch latency / HBM.
#include "common.cuh"

namespace ptrec {

constexpr int kDoThreads = 256;
constexpr int kDoChunk = 1024;  // elements per CTA
constexpr int kDoMaxTensors = 256;

struct DenseHyper {
  int kind;
  float lr, clr, eps, beta1, beta2, wd, step_size, bc2_sqrt;
};

__global__ void __launch_bounds__(kDoThreads)
dense_optim_kernel(const ptrec_dense_tensor* __restrict__ tensors, const int32_t* __restrict__ chunk_start,
                   int n_tensors, DenseHyper h, const float* const* __restrict__ peer_stage, int G, float gscale) {
  __shared__ int s_t;
  if (threadIdx.x == 0) {
    int t = 0;
    while (t + 1 < n_tensors && chunk_start[t + 1] <= (int)blockIdx.x) ++t;
    s_t = t;
  }
  __syncthreads();
  const ptrec_dense_tensor d = tensors[s_t];
  const int64_t base = (int64_t)((int)blockIdx.x - chunk_start[s_t]) * kDoChunk;
  float* p = reinterpret_cast<float*>(d.param);
  const float* g = reinterpret_cast<const float*>(d.grad);
  float* s1 = reinterpret_cast<float*>(d.state1);
  float* s2 = reinterpret_cast<float*>(d.state2);
#pragma unroll
  for (int i = 0; i < kDoChunk / kDoThreads; ++i) {
    const int64_t e = base + i * kDoThreads + threadIdx.x;
    if (e >= d.numel) break;
    float w = p[e];
    float gr;
    if (peer_stage != nullptr) {
      // reduce mode (peer_sync.cu): the gradient is the mean over the ranks' packed stages, summed in rank order so
      // that every replica computes the same bits; element e of tensor t sits at stage[blockIdx.x * chunk + ...]
      const int64_t off = (int64_t)blockIdx.x * kDoChunk + i * kDoThreads + threadIdx.x;
      gr = 0.f;
      for (int r = 0; r < G; ++r) gr += ldg_stream_f1(peer_stage[r] + off);
      gr *= gscale;
    } else {
      gr = g[e];
    }
    if (h.wd != 0.f) gr = gr + h.wd * w;           // grad.add(param, alpha=weight_decay)
    if (h.kind == PTREC_OPT_SGD) {
      w = w + (-h.lr) * gr;                        // param.add_(grad, alpha=-lr)
    } else if (h.kind == PTREC_OPT_LAZY_ADAM) {    // torch.optim.Adam (dense), amsgrad off
      float m = s1[e], v = s2[e];
      m = m + (1.f - h.beta1) * (gr - m);          // exp_avg.lerp_(grad, 1 - beta1)
      v = v * h.beta2 + (1.f - h.beta2) * gr * gr; // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
      s1[e] = m;
      s2[e] = v;
      const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
      w = w + (-h.step_size) * (m / denom);        // param.addcdiv_(exp_avg, denom, value=-step_size)
    } else {                                       // torch.optim.Adagrad
      float s = s1[e];
      s = s + gr * gr;                             // state_sum.addcmul_(grad, grad, value=1)
      s1[e] = s;
      const float sd = sqrtf(s) + h.eps;
      w = w + (-h.clr) * (gr / sd);                // param.addcdiv_(grad, std, value=-clr)
    }
    p[e] = w;
  }
}

}  // namespace ptrec

using namespace ptrec;

static int dense_step_impl(const ptrec_dense_tensor* tensors, const int32_t* chunk_start, int32_t n_tensors,
                           int32_t n_chunks, const ptrec_optim_args* args, const float* const* peer_stage, int32_t G,
                           float gscale, void* stream) {
  PTREC_CHECK_ARG(tensors && chunk_start && args, PTREC_EINVAL, "dense_optim: null pointer");
  PTREC_CHECK_ARG(n_tensors >= 1 && n_tensors <= kDoMaxTensors && n_chunks >= 0, PTREC_EINVAL,
                  "dense_optim: n_tensors=%d out of range (max %d)", n_tensors, kDoMaxTensors);
  PTREC_CHECK_ARG(args->kind == PTREC_OPT_SGD || args->kind == PTREC_OPT_ADAGRAD || args->kind == PTREC_OPT_LAZY_ADAM,
                  PTREC_EUNSUPPORTED, "dense_optim: kind %d has no dense form", args->kind);
  PTREC_CHECK_ARG(args->step >= 1, PTREC_EINVAL, "dense_optim: step must be >= 1");
  if (n_chunks == 0) return PTREC_OK;
  DenseHyper h;
  h.kind = args->kind;
  h.lr = args->lr;
  h.eps = args->eps;
  h.beta1 = args->beta1;
  h.beta2 = args->beta2;
  h.wd = args->weight_decay;
  h.clr = (float)((double)args->lr / (1.0 + (double)(args->step - 1) * (double)args->lr_decay));
  const double bc1 = 1.0 - pow((double)args->beta1, (double)args->step);
  const double bc2 = 1.0 - pow((double)args->beta2, (double)args->step);
  h.step_size = (float)((double)args->lr / (bc1 != 0.0 ? bc1 : 1.0));
  h.bc2_sqrt = (float)sqrt(bc2 > 0.0 ? bc2 : 1.0);
  dense_optim_kernel<<<(unsigned)n_chunks, kDoThreads, 0, (cudaStream_t)stream>>>(tensors, chunk_start, n_tensors, h,
                                                                                    peer_stage, G, gscale);
  PTREC_LAUNCH_CHECK("dense_optim_kernel");
  return PTREC_OK;
}

extern "C" int ptrec_dense_optim_step(const ptrec_dense_tensor* tensors, const int32_t* chunk_start, int32_t n_tensors,
                                      int32_t n_chunks, const ptrec_optim_args* args, void* stream) {
  return dense_step_impl(tensors, chunk_start, n_tensors, n_chunks, args, nullptr, 1, 1.f, stream);
}

extern "C" int ptrec_dense_optim_step_reduce(const ptrec_dense_tensor* tensors, const int32_t* chunk_start,
                                             int32_t n_tensors, int32_t n_chunks, const ptrec_optim_args* args,
                                             const float* const* peer_stage, int32_t G, float grad_scale, void* stream) {
  PTREC_CHECK_ARG(peer_stage && G >= 1 && G <= 64, PTREC_EINVAL, "dense_optim_reduce: bad peer stage / G=%d", G);
  return dense_step_impl(tensors, chunk_start, n_tensors, n_chunks, args, peer_stage, G, grad_scale, stream);
}

extern "C" int32_t ptrec_dense_optim_chunk(void) { return kDoChunk; }
