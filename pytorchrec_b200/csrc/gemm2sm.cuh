// The CTA-pair (cta_group::2, 256 x 256 pair tiles) GEMM of tc_linear.cu with ONE operand plane: a plain bf16 GEMM with
// fp32 accumulation.  K5 (dcn_cross.cu) runs the DCN-v2 cross layers on it: K-major operands with the cross-layer
// epilogues fused, MN-major operands (the reduction runs over the ROWS of both stored matrices) for the weight gradient —
// no transposed copies.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace ptrec {

struct Bf16Gemm {
  bool mn_major;      // false: A [M, K], B [N, K] (K contiguous);  true: A [K, M], B [K, N] (M / N contiguous)
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  int64_t M, N, K;
  const float* bias;  // [N] or null, added to the accumulator
  int dcn_mode;       // 0: fp32 result `out` (split-K allowed);  1: cross forward;  2: cross input gradient (LinEpi)
  const void* p0;     // epilogue operands, bf16 [M][pld]
  const void* p1;
  int64_t pld;
  void* o0;           // epilogue results, bf16 [M][pld]; o1 may be null
  void* o1;
  float* out;         // dcn_mode 0: fp32 [M][ldo]
  int64_t ldo;
  int splits;         // dcn_mode 0: K cut into ranges, partials in `workspace`, summed in a fixed order
  void* workspace;
  size_t workspace_bytes;
};

int gemm_bf16_2sm(const Bf16Gemm& g, cudaStream_t st);

}  // namespace ptrec
