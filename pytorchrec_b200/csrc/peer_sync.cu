// C2: cross-GPU ordering and the dense-gradient all-reduce over NVLink / NVSwitch peer memory.
//
// No reference counterpart: the reference is single-device (torchrec/task/Task.py:187-190).  In the row-wise
// sharded path every rank stores lookup lists and gradient rows straight into the owners' buffers and reads the
// owners' rows with plain loads (a2a_pack.cu, gather_pool.cu), so what the step still needs from a "collective" is
// (a) ordering — "everything every rank enqueued before this point has landed" — and (b) the average of the
// replicated dense-tower gradients.  Both are done here by kernels of this library on symmetric memory instead of
// NCCL launches:
//
//  peer_barrier_kernel     one CTA, one thread per peer: system-scope fence, release-store of this call's epoch into
//                          the peer's flag word, acquire-spin on this rank's own words until every peer's epoch has
//                          arrived.  The epoch counter lives in device memory, so a captured step graph replays it.
//                          ~3 us against ~20 us for a 1-element ncclAllReduce (ring LL) that did the same job.
//  dense_pack_kernel       the step's dense gradients (scattered autograd tensors) -> this rank's contiguous stage
//                          in symmetric memory, one launch for all of them (same descriptor table as K7).
//  K7 reduce mode          dense_optim_kernel (dense_optim.cu) reads each gradient element from EVERY rank's stage
//                          (peer loads), sums them in rank order — identical bits on every rank, so the replicas
//                          cannot drift — scales by 1/G and applies the optimizer: the all-reduce is fused into the
//                          update, no reduced gradient is ever written.
// Bound: launch / NVLink latency (the dense tower of the CTR configs has 0.5 - 1 M parameters).
#include "common.cuh"

namespace ptrec {

constexpr int kSyncMaxRanks = 64;
constexpr int kSyncSlots = 8;

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// flags of rank r: uint32 [kSyncSlots][kSyncMaxRanks] in symmetric memory; peer_flags[p] = rank p's array as seen
// from this GPU.  Word [slot][src] of rank r holds the last epoch rank src announced to r on that slot.
__global__ void __launch_bounds__(kSyncMaxRanks)
peer_barrier_kernel(uint32_t* const* __restrict__ peer_flags, uint32_t* __restrict__ local_epoch, int slot, int G,
                    int my_rank) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = local_epoch[slot] + 1u;
  __syncthreads();
  const uint32_t e = s_epoch;
  const int p = threadIdx.x;
  if (p < G) {
    // stores of the kernels before this one on the stream (to local and to peer memory) are ordered before the flag
    __threadfence_system();
    st_release_sys_u32(peer_flags[p] + slot * kSyncMaxRanks + my_rank, e);
    const uint32_t* mine = peer_flags[my_rank] + slot * kSyncMaxRanks + p;
    // a peer may be a whole step behind at start-up (allocation, graph capture): generous bound, then trap
    for (uint64_t spins = 0; (int32_t)(ld_acquire_sys_u32(mine) - e) < 0; ++spins) {
      if (spins > (1ull << 28)) __trap();
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) local_epoch[slot] = e;
}

constexpr int kPkThreads = 256;
constexpr int kPkChunk = 1024;  // == K7's chunk: tensor t starts at element chunk_start[t] * kPkChunk of the stage

__global__ void __launch_bounds__(kPkThreads)
dense_pack_kernel(const ptrec_dense_tensor* __restrict__ tensors, const int32_t* __restrict__ chunk_start,
                  int n_tensors, float* __restrict__ stage) {
  __shared__ int s_t;
  if (threadIdx.x == 0) {
    int t = 0;
    while (t + 1 < n_tensors && chunk_start[t + 1] <= (int)blockIdx.x) ++t;
    s_t = t;
  }
  __syncthreads();
  const ptrec_dense_tensor d = tensors[s_t];
  const int64_t base = (int64_t)((int)blockIdx.x - chunk_start[s_t]) * kPkChunk;
  const float* g = reinterpret_cast<const float*>(d.grad);
  float* dst = stage + (int64_t)blockIdx.x * kPkChunk;
#pragma unroll
  for (int i = 0; i < kPkChunk / kPkThreads; ++i) {
    const int k = i * kPkThreads + threadIdx.x;
    dst[k] = (base + k < d.numel) ? g[base + k] : 0.f;
  }
}

}  // namespace ptrec

using namespace ptrec;

extern "C" int32_t ptrec_peer_sync_max_ranks(void) { return kSyncMaxRanks; }
extern "C" int32_t ptrec_peer_sync_slots(void) { return kSyncSlots; }

extern "C" int ptrec_peer_barrier(uint32_t* const* peer_flags, uint32_t* local_epoch, int32_t slot, int32_t G,
                                  int32_t my_rank, void* stream) {
  PTREC_CHECK_ARG(peer_flags && local_epoch, PTREC_EINVAL, "peer_barrier: null pointer");
  PTREC_CHECK_ARG(G >= 1 && G <= kSyncMaxRanks && my_rank >= 0 && my_rank < G && slot >= 0 && slot < kSyncSlots,
                  PTREC_EINVAL, "peer_barrier: bad G=%d rank=%d slot=%d", G, my_rank, slot);
  peer_barrier_kernel<<<1, kSyncMaxRanks, 0, (cudaStream_t)stream>>>(peer_flags, local_epoch, slot, G, my_rank);
  PTREC_LAUNCH_CHECK("peer_barrier_kernel");
  return PTREC_OK;
}

extern "C" int ptrec_dense_pack(const ptrec_dense_tensor* tensors, const int32_t* chunk_start, int32_t n_tensors,
                                int32_t n_chunks, float* stage, void* stream) {
  PTREC_CHECK_ARG(tensors && chunk_start && stage, PTREC_EINVAL, "dense_pack: null pointer");
  PTREC_CHECK_ARG(n_tensors >= 1 && n_tensors <= 256 && n_chunks >= 0, PTREC_EINVAL, "dense_pack: n_tensors=%d", n_tensors);
  if (n_chunks == 0) return PTREC_OK;
  dense_pack_kernel<<<(unsigned)n_chunks, kPkThreads, 0, (cudaStream_t)stream>>>(tensors, chunk_start, n_tensors, stage);
  PTREC_LAUNCH_CHECK("dense_pack_kernel");
  return PTREC_OK;
}
