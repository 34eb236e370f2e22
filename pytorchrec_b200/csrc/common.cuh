// Shared device/host helpers for libptrec_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "ptrec_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libptrec_b200 is written for sm_100a (B200) only"
#endif

namespace ptrec {

// ----------------------------------------------------------------------------- host error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();  // every kernel launch of this library is counted (ptrec_launch_count)

#define PTREC_CHECK_ARG(cond, code, ...)  \
  do {                                     \
    if (!(cond)) {                         \
      ::ptrec::set_error(__VA_ARGS__);     \
      return (code);                       \
    }                                      \
  } while (0)

#define PTREC_CUDA(call)                                         \
  do {                                                           \
    cudaError_t _e = (call);                                     \
    if (_e != cudaSuccess) return ::ptrec::cuda_fail(_e, #call); \
  } while (0)

// Gradient-finalising reductions (per-CTA partials -> bias / weight-vector gradients) feed nothing but the optimizer.
// With ptrec_set_reduce_stream(s) set, the entry points that end in such a reduction launch it on `s` behind an event
// recorded on the producer's stream, so it leaves the backward's critical path; the caller joins `s` before the
// optimizer step and gives every call in flight its own partial buffer.  Returns `producer` when no stream is set.
cudaStream_t reduce_stream_after(cudaStream_t producer);

#define PTREC_LAUNCH_CHECK(name)                                          \
  do {                                                                    \
    ::ptrec::count_launch();                                              \
    cudaError_t _e = cudaPeekAtLastError();                               \
    if (_e != cudaSuccess) return ::ptrec::cuda_fail(_e, "launch " name); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kMaxFeatures = 256;  // descriptors are cached in shared memory
constexpr int kMaxTables = 128;
constexpr uint32_t kMaskedKey = 0xFFFFFFFFu;

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// 128-bit streaming loads/stores.  Rows are read once per step: keep them out of L1.
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ldg_stream_f2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
// plain (coherent) 128-bit load: used for rows that this same kernel also writes
__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// A lane's slice of an embedding row: VEC contiguous floats (1, 2 or 4).
template <int VEC>
struct RowVec {
  float v[VEC];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = 0.f;
  }
  __device__ __forceinline__ void add(const RowVec& o) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] += o.v[i];
  }
  __device__ __forceinline__ void scale(float s) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] *= s;
  }
};

template <int VEC>
__device__ __forceinline__ RowVec<VEC> load_row_stream(const float* p) {
  RowVec<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = ldg_stream_f4(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = ldg_stream_f2(p);
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = ldg_stream_f1(p);
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ RowVec<VEC> load_row(const float* p) {
  RowVec<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = *p;
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void store_row(float* p, const RowVec<VEC>& r) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
  } else {
    *p = r.v[0];
  }
}

// ---- bf16 tables (dtype PTREC_BF16): rows stored as bf16, every computation in fp32.  VEC elements = VEC * 2 bytes.
// bf16 -> fp32 is exact (a 16-bit shift); fp32 -> bf16 rounds to nearest even.
__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }
__device__ __forceinline__ uint32_t float_to_bf16_bits(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (u >> 16) | 0x40u;  // NaN stays NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return u >> 16;
}
template <int VEC, bool STREAM>
__device__ __forceinline__ RowVec<VEC> load_row_bf16(const uint16_t* p) {
  RowVec<VEC> r;
  if constexpr (VEC == 4) {
    uint2 t;
    if constexpr (STREAM) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(t.x), "=r"(t.y) : "l"(p));
    else t = *reinterpret_cast<const uint2*>(p);
    r.v[0] = bf16_bits_to_float(t.x & 0xffffu); r.v[1] = bf16_bits_to_float(t.x >> 16);
    r.v[2] = bf16_bits_to_float(t.y & 0xffffu); r.v[3] = bf16_bits_to_float(t.y >> 16);
  } else if constexpr (VEC == 2) {
    uint32_t t;
    if constexpr (STREAM) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(t) : "l"(p));
    else t = *reinterpret_cast<const uint32_t*>(p);
    r.v[0] = bf16_bits_to_float(t & 0xffffu); r.v[1] = bf16_bits_to_float(t >> 16);
  } else {
    uint16_t t;
    if constexpr (STREAM) asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(t) : "l"(p));
    else t = *p;
    r.v[0] = bf16_bits_to_float(t);
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void store_row_bf16(uint16_t* p, const RowVec<VEC>& r) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<uint2*>(p) = make_uint2(float_to_bf16_bits(r.v[0]) | (float_to_bf16_bits(r.v[1]) << 16),
                                              float_to_bf16_bits(r.v[2]) | (float_to_bf16_bits(r.v[3]) << 16));
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<uint32_t*>(p) = float_to_bf16_bits(r.v[0]) | (float_to_bf16_bits(r.v[1]) << 16);
  } else {
    *p = (uint16_t)float_to_bf16_bits(r.v[0]);
  }
}
// table rows of either element type behind one pointer: WB = weights are bf16
template <int VEC, bool WB, bool STREAM>
__device__ __forceinline__ RowVec<VEC> load_table_row(const void* base, int64_t elem_off) {
  if constexpr (WB) return load_row_bf16<VEC, STREAM>(reinterpret_cast<const uint16_t*>(base) + elem_off);
  else if constexpr (STREAM) return load_row_stream<VEC>(reinterpret_cast<const float*>(base) + elem_off);
  else return load_row<VEC>(reinterpret_cast<const float*>(base) + elem_off);
}
template <int VEC, bool WB>
__device__ __forceinline__ void store_table_row(void* base, int64_t elem_off, const RowVec<VEC>& r) {
  if constexpr (WB) store_row_bf16<VEC>(reinterpret_cast<uint16_t*>(base) + elem_off, r);
  else store_row<VEC>(reinterpret_cast<float*>(base) + elem_off, r);
}

// ---- mbarrier + 1-D TMA bulk copy (global -> shared), used to stage index lists -------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a lost transaction traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins > (1u << 24)) __trap();
  }
}
// bytes must be a multiple of 16; src and dst 16-byte aligned.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- feature descriptors cached in shared memory -------------------------------------------------
struct FeatTable {
  ptrec_feature_desc f[kMaxFeatures];
};

__device__ __forceinline__ void load_feats(ptrec_feature_desc* s, const ptrec_feature_desc* g, int F) {
  // 40-byte structs copied as 8-byte words
  const uint64_t* src = reinterpret_cast<const uint64_t*>(g);
  uint64_t* dst = reinterpret_cast<uint64_t*>(s);
  for (int i = threadIdx.x; i < F * 5; i += blockDim.x) dst[i] = src[i];
}

// feature owning slot p (slots are feature-major: feature f owns [id_base_f*B, (id_base_f+L_f)*B))
__device__ __forceinline__ int find_feature(const ptrec_feature_desc* s, int F, int64_t B, int64_t p) {
  int lo = 0, hi = F - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (s[mid].id_base * B <= p) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ bool slot_valid(int mask_mode, int64_t id, int l, const int32_t* lens,
                                           int lens_col, int64_t B, int64_t b) {
  switch (mask_mode) {
    case PTREC_MASK_PAD: return id != 0;
    case PTREC_MASK_PAD_KEEP_FIRST: return id != 0 || l == 0;
    case PTREC_MASK_LENS: return l < lens[(int64_t)lens_col * B + b];
    default: return true;
  }
}

__device__ __forceinline__ float pool_scale(int pooling, int count) {
  float c = (float)(count > 0 ? count : 1);
  if (pooling == PTREC_POOL_MEAN) return 1.0f / c;
  if (pooling == PTREC_POOL_SQRTN) return 1.0f / sqrtf(c);
  return 1.0f;
}

#endif  // __CUDACC__

}  // namespace ptrec
