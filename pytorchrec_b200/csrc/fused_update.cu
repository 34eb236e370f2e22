// K2b: segment-reduce of the pooled-output gradient fused with the sparse row update.
//
// Bound: HBM.  Algorithmic bytes per launch =
//   lookups*(4 perm + D*4 grad row) + segments*(8 key/start + 2*D*4 weight RMW + 2*S state RMW),
//   S = 0 (SGD), 4 (row-wise Adagrad), D*4 (Adagrad), 2*D*4 (lazy Adam).
// One sub-warp of D/4 lanes owns one unique row: it sums the row's gradient slots in sorted
// (= original batch) order with 128-bit loads, then read-modify-writes weight + state once.
// Rows hit more than `kLongSeg` times in the batch (Zipf heads) are deferred to a second kernel in
// which a whole CTA reduces the run with a fixed-shape shared-memory tree.  Both orders are fixed
// by the sort, so results are run-to-run bit-reproducible; no atomics touch floating-point data.
#include "common.cuh"

namespace ptrec {

constexpr int kUpdThreads = 128;
constexpr int kLongThreads = 256;
constexpr int kLongSeg = 32;
constexpr int kLongChunk = 2048;  // slots one CTA reduces; longer runs are split over several CTAs (below)
constexpr int kOptNone = -1;  // segment-sum only

// Scratch of the long-segment path (in the caller's workspace).  Runs of more than kLongSeg slots are queued here by
// the main kernel; one 64-bit atomic hands out BOTH the queue position (high word) and the first partial-sum row of
// the run (low word), so the two stay monotone together.  Runs longer than kLongChunk (a padding id in every history
// of a DIN batch is one run of ~4e5 slots; one CTA walking it alone took 2.2 ms) are reduced chunk by chunk by many
// CTAs into `partial`, and the run's owner adds the chunks in order: the summation order is fixed, no fp atomics.
struct LongWs {
  unsigned long long* counter;  // (runs queued << 32) | partial rows handed out
  int32_t* list;                // [cap] segment index of each queued run
  int32_t* chunk_base;          // [cap] first partial row of the run (only runs longer than kLongChunk own rows)
  float* partial;               // [partial rows][D]
};

struct OptParams {
  int64_t row_stride;  // floats between consecutive rows of a table AND of its element-wise state
  float lr;         // already lr-decayed (Adagrad) / bias-corrected step size (Adam)
  float eps;
  float beta1;
  float beta2;
  float weight_decay;
  float inv_D;
};

// lanes of the sub-warp that owns one row (all of them take the same branches)
template <int LPR>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (LPR == 32) {
    return 0xffffffffu;
  } else {
    return ((1u << LPR) - 1u) << (((threadIdx.x & 31) / LPR) * LPR);
  }
}
template <int VEC, int LPR>
__device__ __forceinline__ float group_sum_sq(const RowVec<VEC>& g) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < VEC; ++k) s += g.v[k] * g.v[k];
  const unsigned m = group_mask<LPR>();
#pragma unroll
  for (int o = 1; o < LPR; o <<= 1) s += __shfl_xor_sync(m, s, o);
  return s;
}

// Row state issued BEFORE the gradient gather so that the weight / state reads overlap it
// (they depend only on the segment key, not on the gradient).
// WB: the weights are bf16 (dtype PTREC_BF16): widened exactly on load, rounded to nearest even on store; optimizer
// state and all arithmetic stay fp32 (state tensors then have the weight's row stride IN ELEMENTS, no interleaving).
template <int VEC, int OPT, bool WB = false>
struct RowState {
  RowVec<VEC> w, a, b;  // weight, state1 slice, state2 slice
  float r;              // row-wise state
  __device__ __forceinline__ void clear() { w.zero(); a.zero(); b.zero(); r = 0.f; }
  __device__ __forceinline__ void load(const void* wrow, const float* s1, const float* s2, int64_t row,
                                       int64_t row_stride, int lane, bool lane_on) {
    const int64_t off = row * row_stride + lane * VEC;
    clear();
    if constexpr (OPT == kOptNone) return;
    if (lane_on) w = load_table_row<VEC, WB, false>(wrow, off);
    if constexpr (OPT == PTREC_OPT_ADAGRAD) { if (lane_on) a = load_row<VEC>(s1 + off); }
    if constexpr (OPT == PTREC_OPT_ROWWISE_ADAGRAD) { r = s1[row]; }
    if constexpr (OPT == PTREC_OPT_LAZY_ADAM) {
      if (lane_on) { a = load_row<VEC>(s1 + off); b = load_row<VEC>(s2 + off); }
    }
  }
};

// apply the optimizer to this lane's VEC-float slice of row `row`
template <int VEC, int LPR, int OPT, bool WB = false>
__device__ __forceinline__ void apply_update(RowVec<VEC> g, const RowState<VEC, OPT, WB>& st, void* __restrict__ wrow,
                                             float* __restrict__ s1, float* __restrict__ s2,
                                             int64_t row, int D, int lane, bool lane_on,
                                             const OptParams& op, float* row_grad_out) {
  const int64_t off = row * op.row_stride + lane * VEC;
  if constexpr (OPT == kOptNone) {
    if (lane_on) store_row<VEC>(row_grad_out + lane * VEC, g);
    return;
  }
  RowVec<VEC> w = st.w;
  if (op.weight_decay != 0.f) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) g.v[k] += op.weight_decay * w.v[k];
  }
  if constexpr (OPT == PTREC_OPT_SGD) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) w.v[k] -= op.lr * g.v[k];
  } else if constexpr (OPT == PTREC_OPT_ADAGRAD) {
    // torch.optim.Adagrad: state_sum.addcmul_(g, g); std = sqrt(state_sum) + eps; p.addcdiv_(g, std, -clr)
    RowVec<VEC> s = st.a;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      s.v[k] += g.v[k] * g.v[k];
      w.v[k] -= op.lr * (g.v[k] / (sqrtf(s.v[k]) + op.eps));
    }
    if (lane_on) store_row<VEC>(s1 + off, s);
  } else if constexpr (OPT == PTREC_OPT_ROWWISE_ADAGRAD) {
    // one accumulator per row: state += mean_k(g_k^2)
    if (!lane_on) g.zero();
    const float ss = group_sum_sq<VEC, LPR>(g) * op.inv_D;
    const float acc = st.r + ss;
    const float inv = op.lr / (sqrtf(acc) + op.eps);
#pragma unroll
    for (int k = 0; k < VEC; ++k) w.v[k] -= inv * g.v[k];
    if (lane == 0) s1[row] = acc;
  } else if constexpr (OPT == PTREC_OPT_LAZY_ADAM) {
    // torch.optim.SparseAdam (_functional.sparse_adam): moments move only on touched rows
    RowVec<VEC> m = st.a, v = st.b;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      m.v[k] += (g.v[k] - m.v[k]) * (1.f - op.beta1);
      v.v[k] += (g.v[k] * g.v[k] - v.v[k]) * (1.f - op.beta2);
      w.v[k] -= op.lr * (m.v[k] / (sqrtf(v.v[k]) + op.eps));
    }
    if (lane_on) {
      store_row<VEC>(s1 + off, m);
      store_row<VEC>(s2 + off, v);
    }
  }
  if (lane_on) store_table_row<VEC, WB>(wrow, off, w);
}

// Per-CTA shared lookup: features of each table (contiguous range) and each feature's first slot.
struct SlotMap {
  ptrec_feature_desc feats[kMaxFeatures];
  uint32_t base[kMaxFeatures];  // first slot of feature f = id_base * B (N < 2^31)
  int16_t f0[kMaxTables];       // first feature of table t
  int16_t nf[kMaxTables];       // number of features of table t
};
__device__ __forceinline__ void build_slot_map(SlotMap* m, const ptrec_feature_desc* g, int F, int T, int64_t B) {
  load_feats(m->feats, g, F);
  for (int t = threadIdx.x; t < T; t += blockDim.x) { m->f0[t] = 0; m->nf[t] = 0; }
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    m->base[f] = (uint32_t)(m->feats[f].id_base * B);
    const int t = m->feats[f].table;
    if (f == 0 || m->feats[f - 1].table != t) m->f0[t] = (int16_t)f;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int n = 0;
    for (int f = m->f0[t]; f < F && m->feats[f].table == t; ++f) ++n;
    m->nf[t] = (int16_t)n;
  }
  __syncthreads();
}

// gradient row of slot p (which belongs to table t), scaled by its bag's pooling factor
template <int VEC>
__device__ __forceinline__ RowVec<VEC> slot_grad(const SlotMap* m, int t, int64_t B,
                                                 int32_t p, const float* __restrict__ grad_out,
                                                 int64_t stride, const float* __restrict__ bag_scale,
                                                 int lane, bool lane_on) {
  int f = m->f0[t];
  for (int e = f + m->nf[t] - 1; f < e && m->base[f + 1] <= (uint32_t)p; ++f) {}
  const ptrec_feature_desc& fd = m->feats[f];
  const uint32_t rel = (uint32_t)p - m->base[f];
  const int64_t b = fd.bag_len == 1 ? rel : rel / (uint32_t)fd.bag_len;
  RowVec<VEC> g;
  g.zero();
  if (lane_on) g = load_row_stream<VEC>(grad_out + b * stride + fd.out_col + lane * VEC);
  if (fd.pooling != PTREC_POOL_SUM && bag_scale != nullptr &&
      !(fd.bag_len == 1 && fd.mask_mode == PTREC_MASK_NONE)) {
    g.scale(bag_scale[(int64_t)f * B + b]);
  }
  return g;
}

template <int VEC, int LPR, int OPT, bool WB>
__global__ void __launch_bounds__(kUpdThreads)
fused_update_kernel(void* const* __restrict__ table_ptrs, void* const* __restrict__ state1_ptrs,
                    void* const* __restrict__ state2_ptrs, int T, int D,
                    const ptrec_feature_desc* __restrict__ feats, int F, int64_t B,
                    const uint32_t* __restrict__ sorted_keys, const int32_t* __restrict__ perm,
                    const int32_t* __restrict__ seg_start, const ptrec_segment_meta* __restrict__ seg_meta,
                    const int32_t* __restrict__ n_seg_ptr, const float* __restrict__ grad_out,
                    int64_t stride, const float* __restrict__ bag_scale, OptParams op, LongWs lw,
                    float* __restrict__ row_grad) {
  __shared__ SlotMap s_map;
  build_slot_map(&s_map, feats, F, T, B);
  constexpr int NSG = kUpdThreads / LPR;
  constexpr int SEGS = 4;  // unique rows in flight per sub-warp: 4 x (weight + state + gradient) rows requested together
  const int sg = threadIdx.x / LPR, lane = threadIdx.x % LPR;
  const bool lane_on = lane * VEC < D;
  const int n_seg = *n_seg_ptr;
  const int n_round = (n_seg + NSG * SEGS - 1) / (NSG * SEGS);
  // Segment metadata is ONE round of independent loads (bounds + a 16-byte record), and the metadata of the next
  // round is requested before this round's rows: the only exposed dependency per iteration is the row fetch.
  int nstart[SEGS], nend[SEGS];
  int4 nmeta[SEGS];
  auto fetch_meta = [&](int r) {
#pragma unroll
    for (int q = 0; q < SEGS; ++q) {
      const int u = (r * SEGS + q) * NSG + sg;
      nstart[q] = nend[q] = 0;
      nmeta[q] = make_int4((int)kMaskedKey, 0, 0, 0);
      if (r < n_round && u < n_seg) {
        nstart[q] = seg_start[u];
        nend[q] = seg_start[u + 1];
        nmeta[q] = *reinterpret_cast<const int4*>(seg_meta + u);
      }
    }
  };
  fetch_meta(blockIdx.x);
  for (int r = blockIdx.x; r < n_round; r += gridDim.x) {
    int start[SEGS], end[SEGS], tab[SEGS];
    uint32_t key[SEGS];
    int32_t p0[SEGS];
    bool work[SEGS], live[SEGS];
#pragma unroll
    for (int q = 0; q < SEGS; ++q) {
      const int u = (r * SEGS + q) * NSG + sg;
      live[q] = u < n_seg;
      start[q] = nstart[q];
      end[q] = nend[q];
      key[q] = (uint32_t)nmeta[q].x;
      p0[q] = nmeta[q].y;
      tab[q] = nmeta[q].z;
    }
    fetch_meta(r + gridDim.x);
    // round 3: weight / state rows and the first gradient row of every segment
    void* wrow[SEGS];
    float* s1[SEGS];
    float* s2[SEGS];
    RowState<VEC, OPT, WB> st[SEGS];
    RowVec<VEC> acc[SEGS];
#pragma unroll
    for (int q = 0; q < SEGS; ++q) {
      const int u = (r * SEGS + q) * NSG + sg;
      const bool masked = key[q] == kMaskedKey;
      const bool is_long = live[q] && !masked && (end[q] - start[q]) > kLongSeg && OPT != kOptNone;
      if (is_long && lane == 0) {
        const int len = end[q] - start[q];
        const unsigned nch = len > kLongChunk ? (unsigned)((len + kLongChunk - 1) / kLongChunk) : 0u;
        const unsigned long long old = atomicAdd(lw.counter, (1ull << 32) | nch);
        lw.list[old >> 32] = u;
        lw.chunk_base[old >> 32] = (int32_t)(old & 0xffffffffu);
      }
      work[q] = live[q] && !masked && !is_long;
      wrow[q] = nullptr;
      s1[q] = s2[q] = nullptr;
      st[q].clear();
      acc[q].zero();
      if (work[q]) {
        if constexpr (OPT != kOptNone) {
          wrow[q] = table_ptrs[tab[q]];
          if (state1_ptrs) s1[q] = reinterpret_cast<float*>(state1_ptrs[tab[q]]);
          if (state2_ptrs) s2[q] = reinterpret_cast<float*>(state2_ptrs[tab[q]]);
          st[q].load(wrow[q], s1[q], s2[q], (int64_t)key[q], op.row_stride, lane, lane_on);
        }
        acc[q] = slot_grad<VEC>(&s_map, tab[q], B, p0[q], grad_out, stride, bag_scale, lane, lane_on);
      }
    }
    // remaining gradient slots of longer segments (duplicates in the batch)
#pragma unroll
    for (int q = 0; q < SEGS; ++q) {
      if (work[q]) {
        for (int j0 = start[q] + 1; j0 < end[q]; j0 += 4) {
          RowVec<VEC> g[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            g[x].zero();
            if (j0 + x < end[q])
              g[x] = slot_grad<VEC>(&s_map, tab[q], B, perm[j0 + x], grad_out, stride, bag_scale, lane, lane_on);
          }
#pragma unroll
          for (int x = 0; x < 4; ++x) acc[q].add(g[x]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < SEGS; ++q) {
      const int u = (r * SEGS + q) * NSG + sg;
      if constexpr (OPT == kOptNone) {
        if (live[q]) apply_update<VEC, LPR, OPT, WB>(acc[q], st[q], nullptr, nullptr, nullptr, 0, D, lane, lane_on, op,
                                                     row_grad + (int64_t)u * D);
      } else {
        if (work[q]) apply_update<VEC, LPR, OPT, WB>(acc[q], st[q], wrow[q], s1[q], s2[q], (int64_t)key[q], D, lane,
                                                     lane_on, op, nullptr);
      }
    }
  }
}

// sum of the gradient rows of sorted slots [beg, end) of table t by one CTA: sub-warp sg walks slots beg + sg,
// beg + sg + 2 NSG, ... two at a time, then a fixed-shape tree over the sub-warps.  The total ends up in the lanes of
// sub-warp 0 (as `mine`, in shared memory).  Must be called by every thread of the CTA.
template <int VEC, int LPR>
__device__ __forceinline__ float* cta_slot_sum(const SlotMap* s_map, float* s_part, int t, int64_t B, int beg, int end,
                                               const int32_t* __restrict__ perm, const float* __restrict__ grad_out,
                                               int64_t stride, const float* __restrict__ bag_scale, int sg, int lane,
                                               bool lane_on) {
  constexpr int NSG = kLongThreads / LPR;
  RowVec<VEC> acc;
  acc.zero();
  for (int j0 = beg + sg; j0 < end; j0 += NSG * 2) {
    RowVec<VEC> g0, g1;
    g0 = slot_grad<VEC>(s_map, t, B, perm[j0], grad_out, stride, bag_scale, lane, lane_on);
    g1.zero();
    if (j0 + NSG < end)
      g1 = slot_grad<VEC>(s_map, t, B, perm[j0 + NSG], grad_out, stride, bag_scale, lane, lane_on);
    acc.add(g0);
    acc.add(g1);
  }
  float* mine = s_part + (sg * LPR + lane) * VEC;
#pragma unroll
  for (int k = 0; k < VEC; ++k) mine[k] = acc.v[k];
  __syncthreads();
  for (int h = NSG / 2; h > 0; h >>= 1) {
    if (sg < h) {
      const float* other = s_part + ((sg + h) * LPR + lane) * VEC;
#pragma unroll
      for (int k = 0; k < VEC; ++k) mine[k] += other[k];
    }
    __syncthreads();
  }
  return mine;
}

// pass A: CTA per kLongChunk-slot chunk of the runs longer than kLongChunk -> one partial row each
template <int VEC, int LPR>
__global__ void __launch_bounds__(kLongThreads)
fused_update_chunk_kernel(int T, int D, const ptrec_feature_desc* __restrict__ feats, int F, int64_t B,
                          const int32_t* __restrict__ perm, const int32_t* __restrict__ seg_start,
                          const ptrec_segment_meta* __restrict__ seg_meta, const float* __restrict__ grad_out,
                          int64_t stride, const float* __restrict__ bag_scale, LongWs lw) {
  const unsigned long long packed = *lw.counter;
  const int n_long = (int)(packed >> 32), n_chunks = (int)(packed & 0xffffffffu);
  if ((int)blockIdx.x >= n_chunks) return;  // the common case (no giant runs): exit before any setup
  __shared__ SlotMap s_map;
  constexpr int NSG = kLongThreads / LPR;
  __shared__ float s_part[NSG * LPR * VEC];
  build_slot_map(&s_map, feats, F, T, B);
  const int sg = threadIdx.x / LPR, lane = threadIdx.x % LPR;
  const bool lane_on = lane * VEC < D;
  for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
    // the queued run owning partial row ci: the last entry whose first row is <= ci (entries without rows share the
    // first row of the next entry that has some, and come before it)
    int lo = 0, hi = n_long - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (lw.chunk_base[mid] <= ci) lo = mid; else hi = mid - 1;
    }
    const int u = lw.list[lo];
    const int c = ci - lw.chunk_base[lo];
    const int start = seg_start[u], end = seg_start[u + 1];
    const int beg = start + c * kLongChunk;
    const float* mine = cta_slot_sum<VEC, LPR>(&s_map, s_part, seg_meta[u].table, B, beg, min(end, beg + kLongChunk), perm,
                                               grad_out, stride, bag_scale, sg, lane, lane_on);
    if (sg == 0 && lane_on) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) lw.partial[(int64_t)ci * D + lane * VEC + k] = mine[k];
    }
    __syncthreads();
  }
}

// pass B: CTA per queued run: reduce it (directly, or from its chunks' partial rows in chunk order) and update the row
template <int VEC, int LPR, int OPT, bool WB>
__global__ void __launch_bounds__(kLongThreads)
fused_update_long_kernel(void* const* __restrict__ table_ptrs, void* const* __restrict__ state1_ptrs,
                         void* const* __restrict__ state2_ptrs, int T, int D,
                         const ptrec_feature_desc* __restrict__ feats, int F, int64_t B,
                         const uint32_t* __restrict__ sorted_keys, const int32_t* __restrict__ perm,
                         const int32_t* __restrict__ seg_start, const ptrec_segment_meta* __restrict__ seg_meta,
                         const float* __restrict__ grad_out, int64_t stride,
                         const float* __restrict__ bag_scale, OptParams op, LongWs lw) {
  const int n_long = (int)(*lw.counter >> 32);
  if ((int)blockIdx.x >= n_long) return;  // the common case (no hot rows): exit before any setup
  __shared__ SlotMap s_map;
  constexpr int NSG = kLongThreads / LPR;
  __shared__ float s_part[NSG * LPR * VEC];
  build_slot_map(&s_map, feats, F, T, B);
  const int sg = threadIdx.x / LPR, lane = threadIdx.x % LPR;
  const bool lane_on = lane * VEC < D;
  for (int i = blockIdx.x; i < n_long; i += gridDim.x) {
    const int u = lw.list[i];
    const int start = seg_start[u], end = seg_start[u + 1];
    const int t = seg_meta[u].table;
    RowVec<VEC> tot;
    tot.zero();
    if (end - start > kLongChunk) {
      if (sg == 0 && lane_on) {
        const int nch = (end - start + kLongChunk - 1) / kLongChunk;
        const float* p = lw.partial + (int64_t)lw.chunk_base[i] * D + lane * VEC;
        for (int c = 0; c < nch; ++c) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) tot.v[k] += p[(int64_t)c * D + k];
        }
      }
    } else {
      const float* mine = cta_slot_sum<VEC, LPR>(&s_map, s_part, t, B, start, end, perm, grad_out, stride, bag_scale, sg,
                                                 lane, lane_on);
      if (sg == 0) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) tot.v[k] = mine[k];
      }
    }
    if (sg == 0) {  // sub-warp 0 owns the row
      const uint32_t key = sorted_keys[start];
      void* wrow = table_ptrs[t];
      float* s1 = state1_ptrs ? reinterpret_cast<float*>(state1_ptrs[t]) : nullptr;
      float* s2 = state2_ptrs ? reinterpret_cast<float*>(state2_ptrs[t]) : nullptr;
      RowState<VEC, OPT, WB> st;
      st.load(wrow, s1, s2, (int64_t)key, op.row_stride, lane, lane_on);
      apply_update<VEC, LPR, OPT, WB>(tot, st, wrow, s1, s2, (int64_t)key, D, lane, lane_on, op, nullptr);
    }
    __syncthreads();
  }
}

template <int VEC, int LPR, int OPT, bool WB>
static int launch_update(void* const* table_ptrs, void* const* s1, void* const* s2, int T, int D,
                         const ptrec_feature_desc* feats, int F, int64_t B, int64_t N,
                         const uint32_t* sorted_keys, const int32_t* perm, const int32_t* seg_start,
                         const ptrec_segment_meta* seg_meta, const int32_t* n_seg, const float* grad_out,
                         int64_t stride, const float* bag_scale, const OptParams& op, const LongWs& long_ws,
                         float* row_grad, cudaStream_t st) {
  constexpr int NSG = kUpdThreads / LPR;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t rounds = ceil_div(N, NSG * 4);
  // 4 CTAs per SM (register-limited): one wave, several rounds per CTA so that the metadata prefetch has a next round
  const unsigned grid = (unsigned)(rounds < (int64_t)sms * 4 ? (rounds > 0 ? rounds : 1) : (int64_t)sms * 4);
  if (OPT != kOptNone) PTREC_CUDA(cudaMemsetAsync(long_ws.counter, 0, sizeof(unsigned long long), st));
  fused_update_kernel<VEC, LPR, OPT, WB><<<grid, kUpdThreads, 0, st>>>(
      table_ptrs, s1, s2, T, D, feats, F, B, sorted_keys, perm, seg_start, seg_meta, n_seg, grad_out,
      stride, bag_scale, op, long_ws, row_grad);
  PTREC_LAUNCH_CHECK("fused_update_kernel");
  if (OPT != kOptNone) {
    fused_update_chunk_kernel<VEC, LPR><<<sms * 2, kLongThreads, 0, st>>>(T, D, feats, F, B, perm, seg_start, seg_meta,
                                                                        grad_out, stride, bag_scale, long_ws);
    PTREC_LAUNCH_CHECK("fused_update_chunk_kernel");
    fused_update_long_kernel<VEC, LPR, OPT, WB><<<sms * 2, kLongThreads, 0, st>>>(
        table_ptrs, s1, s2, T, D, feats, F, B, sorted_keys, perm, seg_start, seg_meta, grad_out, stride,
        bag_scale, op, long_ws);
    PTREC_LAUNCH_CHECK("fused_update_long_kernel");
  }
  return PTREC_OK;
}

template <int OPT, bool WB = false>
static int dispatch_D(void* const* table_ptrs, void* const* s1, void* const* s2, int T, int D,
                      const ptrec_feature_desc* feats, int F, int64_t B, int64_t N,
                      const uint32_t* sorted_keys, const int32_t* perm, const int32_t* seg_start,
                      const ptrec_segment_meta* seg_meta, const int32_t* n_seg, const float* grad_out,
                      int64_t stride, const float* bag_scale, const OptParams& op, const LongWs& long_ws,
                      float* row_grad, cudaStream_t st) {
#define PTREC_UPD(V, P) \
  return launch_update<V, P, OPT, WB>(table_ptrs, s1, s2, T, D, feats, F, B, N, sorted_keys, perm, seg_start, \
                                  seg_meta, n_seg, grad_out, stride, bag_scale, op, long_ws, row_grad, st)
  if (D == 1) PTREC_UPD(1, 1);
  if (D == 2) PTREC_UPD(2, 1);
  const int lanes = D / 4;
  if (lanes <= 1) PTREC_UPD(4, 1);
  if (lanes <= 2) PTREC_UPD(4, 2);
  if (lanes <= 4) PTREC_UPD(4, 4);
  if (lanes <= 8) PTREC_UPD(4, 8);
  if (lanes <= 16) PTREC_UPD(4, 16);
  PTREC_UPD(4, 32);
#undef PTREC_UPD
}

}  // namespace ptrec

using namespace ptrec;

static size_t long_cap(int64_t N) { return (size_t)N / kLongSeg + 2; }               // runs that can be queued
static size_t long_partial_rows(int64_t N) { return 2 * ((size_t)N / kLongChunk) + 2; }  // sum ceil(len / chunk), len > chunk
extern "C" size_t ptrec_embedding_bwd_workspace_bytes(int64_t N, int32_t D) {
  return 256 + 2 * align_up(long_cap(N) * sizeof(int32_t), 256) +
         align_up(long_partial_rows(N) * (size_t)(D > 0 ? D : 1) * sizeof(float), 256);
}

static int64_t total_slots(const ptrec_feature_desc* feats_host, int F, int64_t B) {
  int64_t L = 0;
  for (int f = 0; f < F; ++f) L += feats_host[f].bag_len;
  return L * B;
}

static int check_common(int32_t T, int32_t D, int32_t dtype, int32_t F, const void* grad_out,
                        int64_t stride) {
  PTREC_CHECK_ARG(dtype == PTREC_F32 || dtype == PTREC_BF16, PTREC_EUNSUPPORTED, "bwd: unknown table dtype %d", dtype);
  PTREC_CHECK_ARG(T >= 1 && T <= kMaxTables && F >= 1 && F <= kMaxFeatures, PTREC_EINVAL, "bwd: T=%d F=%d out of range", T, F);
  const bool d_ok = D == 1 || D == 2 || (D >= 4 && D <= 128 && D % 4 == 0);
  PTREC_CHECK_ARG(d_ok, PTREC_EUNSUPPORTED, "bwd: D=%d unsupported", D);
  const int vec = D >= 4 ? 4 : D;
  PTREC_CHECK_ARG(((uintptr_t)grad_out % (vec * 4)) == 0 && (stride % vec) == 0, PTREC_EALIGN,
                  "bwd: grad_out / stride not aligned to %d bytes", vec * 4);
  return PTREC_OK;
}

extern "C" int ptrec_embedding_bwd_fused(void* const* table_ptrs, void* const* state1_ptrs,
                                         void* const* state2_ptrs, int32_t T, int32_t D, int64_t row_stride,
                                         int32_t dtype, const ptrec_feature_desc* feats,
                                         const ptrec_feature_desc* feats_host, int32_t F, int64_t B,
                                         const uint32_t* sorted_keys, const int32_t* perm,
                                         const int32_t* seg_start, const ptrec_segment_meta* seg_meta,
                                         const int32_t* n_seg, const float* grad_out,
                                         int64_t grad_row_stride, const float* bag_scale,
                                         const ptrec_optim_args* opt_host, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  int rc = check_common(T, D, dtype, F, grad_out, grad_row_stride);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(table_ptrs && feats && feats_host && sorted_keys && perm && seg_start && seg_meta && n_seg &&
                      grad_out && opt_host, PTREC_EINVAL, "bwd_fused: null pointer");
  const int64_t N = total_slots(feats_host, F, B);
  PTREC_CHECK_ARG(workspace && workspace_bytes >= ptrec_embedding_bwd_workspace_bytes(N, D), PTREC_EWORKSPACE,
                  "bwd_fused: workspace too small");
  if (N == 0) return PTREC_OK;
  for (int f = 0; f < F; ++f) {
    const int vec = D >= 4 ? 4 : D;
    PTREC_CHECK_ARG(feats_host[f].out_col % vec == 0, PTREC_EALIGN, "bwd_fused: feature %d out_col misaligned", f);
  }
  OptParams op;
  PTREC_CHECK_ARG(row_stride >= D && row_stride % (D >= 4 ? 4 : D) == 0, PTREC_EALIGN, "bwd_fused: bad table row_stride");
  op.row_stride = row_stride;
  op.eps = opt_host->eps;
  op.beta1 = opt_host->beta1;
  op.beta2 = opt_host->beta2;
  op.weight_decay = opt_host->weight_decay;
  op.inv_D = 1.0f / (float)D;
  op.lr = opt_host->lr;
  const int step = opt_host->step < 1 ? 1 : opt_host->step;
  LongWs long_ws;
  {
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    long_ws.counter = reinterpret_cast<unsigned long long*>(w);
    w += 256;
    long_ws.list = reinterpret_cast<int32_t*>(w);
    w += align_up(long_cap(N) * sizeof(int32_t), 256);
    long_ws.chunk_base = reinterpret_cast<int32_t*>(w);
    w += align_up(long_cap(N) * sizeof(int32_t), 256);
    long_ws.partial = reinterpret_cast<float*>(w);
  }
  cudaStream_t st = (cudaStream_t)stream;
#define PTREC_ARGS                                                                                      \
  table_ptrs, state1_ptrs, state2_ptrs, T, D, feats, F, B, N, sorted_keys, perm, seg_start, seg_meta, n_seg, \
      grad_out, grad_row_stride, bag_scale, op, long_ws, nullptr, st
  const bool wb = dtype == PTREC_BF16;
  switch (opt_host->kind) {
    case PTREC_OPT_SGD:
      return wb ? dispatch_D<PTREC_OPT_SGD, true>(PTREC_ARGS) : dispatch_D<PTREC_OPT_SGD>(PTREC_ARGS);
    case PTREC_OPT_ADAGRAD:
      PTREC_CHECK_ARG(state1_ptrs, PTREC_EINVAL, "bwd_fused: Adagrad needs state1");
      op.lr = (float)((double)opt_host->lr / (1.0 + (double)(step - 1) * (double)opt_host->lr_decay));
      return wb ? dispatch_D<PTREC_OPT_ADAGRAD, true>(PTREC_ARGS) : dispatch_D<PTREC_OPT_ADAGRAD>(PTREC_ARGS);
    case PTREC_OPT_ROWWISE_ADAGRAD:
      PTREC_CHECK_ARG(state1_ptrs, PTREC_EINVAL, "bwd_fused: row-wise Adagrad needs state1");
      op.lr = (float)((double)opt_host->lr / (1.0 + (double)(step - 1) * (double)opt_host->lr_decay));
      return wb ? dispatch_D<PTREC_OPT_ROWWISE_ADAGRAD, true>(PTREC_ARGS) : dispatch_D<PTREC_OPT_ROWWISE_ADAGRAD>(PTREC_ARGS);
    case PTREC_OPT_LAZY_ADAM: {
      PTREC_CHECK_ARG(state1_ptrs && state2_ptrs, PTREC_EINVAL, "bwd_fused: lazy Adam needs state1 and state2");
      const double bc1 = 1.0 - pow((double)opt_host->beta1, (double)step);
      const double bc2 = 1.0 - pow((double)opt_host->beta2, (double)step);
      op.lr = (float)((double)opt_host->lr * sqrt(bc2) / bc1);
      return wb ? dispatch_D<PTREC_OPT_LAZY_ADAM, true>(PTREC_ARGS) : dispatch_D<PTREC_OPT_LAZY_ADAM>(PTREC_ARGS);
    }
    default:
      PTREC_CHECK_ARG(false, PTREC_EINVAL, "bwd_fused: unknown optimizer kind %d", opt_host->kind);
  }
#undef PTREC_ARGS
  return PTREC_OK;
}

#define PTREC_NAMED(NAME, KIND)                                                                          \
  extern "C" int NAME(void* const* a, void* const* b, void* const* c, int32_t T, int32_t D, int64_t rs, int32_t dt, \
                      const ptrec_feature_desc* f, const ptrec_feature_desc* fh, int32_t F, int64_t B,   \
                      const uint32_t* k, const int32_t* p, const int32_t* s, const ptrec_segment_meta* st,          \
                      const int32_t* n, const float* g, int64_t gs, const float* bs,                     \
                      const ptrec_optim_args* o, void* ws, size_t wsb, void* stream) {                   \
    PTREC_CHECK_ARG(o && o->kind == KIND, PTREC_EINVAL, #NAME ": opt_host->kind must be " #KIND);        \
    return ptrec_embedding_bwd_fused(a, b, c, T, D, rs, dt, f, fh, F, B, k, p, s, st, n, g, gs, bs, o, ws, \
                                     wsb, stream);                                                        \
  }
PTREC_NAMED(ptrec_embedding_bwd_fused_sgd, PTREC_OPT_SGD)
PTREC_NAMED(ptrec_embedding_bwd_fused_adagrad, PTREC_OPT_ADAGRAD)
PTREC_NAMED(ptrec_embedding_bwd_fused_rowwise_adagrad, PTREC_OPT_ROWWISE_ADAGRAD)
PTREC_NAMED(ptrec_embedding_bwd_fused_lazy_adam, PTREC_OPT_LAZY_ADAM)
#undef PTREC_NAMED

extern "C" int ptrec_embedding_bwd_segment_sum(int32_t T, int32_t D, const ptrec_feature_desc* feats,
                                               const ptrec_feature_desc* feats_host, int32_t F,
                                               int64_t B, const uint32_t* sorted_keys,
                                               const int32_t* perm, const int32_t* seg_start,
                                               const ptrec_segment_meta* seg_meta, const int32_t* n_seg,
                                               const float* grad_out, int64_t grad_row_stride,
                                               const float* bag_scale, float* row_grad, void* stream) {
  int rc = check_common(T, D, PTREC_F32, F, grad_out, grad_row_stride);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(feats && feats_host && sorted_keys && perm && seg_start && seg_meta && n_seg && row_grad,
                  PTREC_EINVAL, "segment_sum: null pointer");
  PTREC_CHECK_ARG(aligned16(row_grad), PTREC_EALIGN, "segment_sum: row_grad must be 16-byte aligned");
  const int64_t N = total_slots(feats_host, F, B);
  if (N == 0) return PTREC_OK;
  OptParams op{};
  LongWs none{};
  return dispatch_D<kOptNone>(nullptr, nullptr, nullptr, T, D, feats, F, B, N, sorted_keys, perm, seg_start,
                              seg_meta, n_seg, grad_out, grad_row_stride, bag_scale, op, none, row_grad,
                              (cudaStream_t)stream);
}
