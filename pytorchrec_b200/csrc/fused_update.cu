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
constexpr int kOptNone = -1;  // segment-sum only

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
template <int VEC, int OPT>
struct RowState {
  RowVec<VEC> w, a, b;  // weight, state1 slice, state2 slice
  float r;              // row-wise state
  __device__ __forceinline__ void clear() { w.zero(); a.zero(); b.zero(); r = 0.f; }
  __device__ __forceinline__ void load(const float* wrow, const float* s1, const float* s2, int64_t row,
                                       int64_t row_stride, int lane, bool lane_on) {
    const int64_t off = row * row_stride + lane * VEC;
    clear();
    if constexpr (OPT == kOptNone) return;
    if (lane_on) w = load_row<VEC>(wrow + off);
    if constexpr (OPT == PTREC_OPT_ADAGRAD) { if (lane_on) a = load_row<VEC>(s1 + off); }
    if constexpr (OPT == PTREC_OPT_ROWWISE_ADAGRAD) { r = s1[row]; }
    if constexpr (OPT == PTREC_OPT_LAZY_ADAM) {
      if (lane_on) { a = load_row<VEC>(s1 + off); b = load_row<VEC>(s2 + off); }
    }
  }
};

// apply the optimizer to this lane's VEC-float slice of row `row`
template <int VEC, int LPR, int OPT>
__device__ __forceinline__ void apply_update(RowVec<VEC> g, const RowState<VEC, OPT>& st, float* __restrict__ wrow,
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
  if (lane_on) store_row<VEC>(wrow + off, w);
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

template <int VEC, int LPR, int OPT>
__global__ void __launch_bounds__(kUpdThreads)
fused_update_kernel(void* const* __restrict__ table_ptrs, void* const* __restrict__ state1_ptrs,
                    void* const* __restrict__ state2_ptrs, int T, int D,
                    const ptrec_feature_desc* __restrict__ feats, int F, int64_t B,
                    const uint32_t* __restrict__ sorted_keys, const int32_t* __restrict__ perm,
                    const int32_t* __restrict__ seg_start, const ptrec_segment_meta* __restrict__ seg_meta,
                    const int32_t* __restrict__ n_seg_ptr, const float* __restrict__ grad_out,
                    int64_t stride, const float* __restrict__ bag_scale, OptParams op,
                    int* __restrict__ long_count, int32_t* __restrict__ long_list,
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
    float* wrow[SEGS];
    float* s1[SEGS];
    float* s2[SEGS];
    RowState<VEC, OPT> st[SEGS];
    RowVec<VEC> acc[SEGS];
#pragma unroll
    for (int q = 0; q < SEGS; ++q) {
      const int u = (r * SEGS + q) * NSG + sg;
      const bool masked = key[q] == kMaskedKey;
      const bool is_long = live[q] && !masked && (end[q] - start[q]) > kLongSeg && OPT != kOptNone;
      if (is_long && lane == 0) long_list[atomicAdd(long_count, 1)] = u;
      work[q] = live[q] && !masked && !is_long;
      wrow[q] = s1[q] = s2[q] = nullptr;
      st[q].clear();
      acc[q].zero();
      if (work[q]) {
        if constexpr (OPT != kOptNone) {
          wrow[q] = reinterpret_cast<float*>(table_ptrs[tab[q]]);
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
        if (live[q]) apply_update<VEC, LPR, OPT>(acc[q], st[q], nullptr, nullptr, nullptr, 0, D, lane, lane_on, op,
                                                 row_grad + (int64_t)u * D);
      } else {
        if (work[q]) apply_update<VEC, LPR, OPT>(acc[q], st[q], wrow[q], s1[q], s2[q], (int64_t)key[q], D, lane, lane_on,
                                                 op, nullptr);
      }
    }
  }
}

// ---- Adagrad, v2 thread mapping ---------------------------------------------------------------------------------
// 2*LPR lanes own one unique row: lanes [0, LPR) its weight slice, lanes [LPR, 2*LPR) the matching slice of the
// Adagrad sum.  With the interleaved layout (state = weight + D, row_stride = 2*D) one 128-bit load per lane reads
// the row's whole 2*D*4-byte bundle as ONE contiguous request (a full 128-byte line at D = 16) and one 128-bit
// store per lane writes it back as a full line; v1 issues two half-line loads and two half-line stores per row.
// Both halves load the (same) gradient slice — a broadcast inside the load instruction — exchange w / sum with one
// shuffle per float and compute the update redundantly; each stores its own half.  No metadata is carried across
// rounds (v1 prefetches the next round's 6 words per segment into registers): the kernel stays under
// 65536 / (128 * MINB) registers so that MINB CTAs are resident per SM and the latency of a round is hidden by other
// warps instead of by registers.  Same arithmetic and the same (sorted = batch) summation order as v1.
template <int VEC, int LPR, int SEGS, int MINB>
__global__ void __launch_bounds__(kUpdThreads, MINB)
fused_adagrad_pair_kernel(void* const* __restrict__ table_ptrs, void* const* __restrict__ state1_ptrs, int T, int D,
                          const ptrec_feature_desc* __restrict__ feats, int F, int64_t B,
                          const int32_t* __restrict__ perm, const int32_t* __restrict__ seg_start,
                          const ptrec_segment_meta* __restrict__ seg_meta, const int32_t* __restrict__ n_seg_ptr,
                          const float* __restrict__ grad_out, int64_t stride, const float* __restrict__ bag_scale,
                          OptParams op, int* __restrict__ long_count, int32_t* __restrict__ long_list) {
  __shared__ SlotMap s_map;
  build_slot_map(&s_map, feats, F, T, B);
  constexpr int SW = 2 * LPR;  // lanes per row
  constexpr int NSG = kUpdThreads / SW;
  const int sg = threadIdx.x / SW, sl = threadIdx.x % SW;
  const int half = sl / LPR, lane = sl % LPR;
  const bool lane_on = lane * VEC < D;
  const int n_seg = *n_seg_ptr;
  const int n_round = (n_seg + NSG * SEGS - 1) / (NSG * SEGS);
  for (int r = blockIdx.x; r < n_round; r += gridDim.x) {
    int start[SEGS], end[SEGS];
    int4 meta[SEGS];
#pragma unroll
    for (int q = 0; q < SEGS; ++q) {
      const int u = (r * SEGS + q) * NSG + sg;
      start[q] = end[q] = 0;
      meta[q] = make_int4((int)kMaskedKey, 0, 0, 0);
      if (u < n_seg) {
        start[q] = seg_start[u];
        end[q] = seg_start[u + 1];
        meta[q] = *reinterpret_cast<const int4*>(seg_meta + u);
      }
    }
    float* rowp[SEGS];
    RowVec<VEC> x[SEGS], acc[SEGS];
    bool work[SEGS];
#pragma unroll
    for (int q = 0; q < SEGS; ++q) {
      const int u = (r * SEGS + q) * NSG + sg;
      const bool live = u < n_seg && (uint32_t)meta[q].x != kMaskedKey;
      const bool is_long = live && (end[q] - start[q]) > kLongSeg;
      if (is_long && sl == 0) long_list[atomicAdd(long_count, 1)] = u;
      work[q] = live && !is_long;
      rowp[q] = nullptr;
      x[q].zero();
      acc[q].zero();
      if (work[q]) {
        float* base = reinterpret_cast<float*>(half == 0 ? table_ptrs[meta[q].z] : state1_ptrs[meta[q].z]);
        rowp[q] = base + (int64_t)(uint32_t)meta[q].x * op.row_stride + lane * VEC;
        if (lane_on) x[q] = load_row<VEC>(rowp[q]);
        acc[q] = slot_grad<VEC>(&s_map, meta[q].z, B, meta[q].y, grad_out, stride, bag_scale, lane, lane_on);
      }
    }
    // remaining gradient slots of longer segments (duplicates in the batch), in sorted order
#pragma unroll
    for (int q = 0; q < SEGS; ++q) {
      if (work[q]) {
        for (int j0 = start[q] + 1; j0 < end[q]; j0 += 4) {
          RowVec<VEC> g[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            g[k].zero();
            if (j0 + k < end[q])
              g[k] = slot_grad<VEC>(&s_map, meta[q].z, B, perm[j0 + k], grad_out, stride, bag_scale, lane, lane_on);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[q].add(g[k]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < SEGS; ++q) {
      RowVec<VEC> out;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        // the partner lane (same slice, other half) holds sum if this lane holds w, and vice versa; the shuffle is
        // executed by every lane of the warp (work[] only gates the store)
        const float other = __shfl_xor_sync(0xffffffffu, x[q].v[k], LPR);
        const float w = half == 0 ? x[q].v[k] : other;
        float sum = half == 0 ? other : x[q].v[k];
        float g = acc[q].v[k];
        if (op.weight_decay != 0.f) g += op.weight_decay * w;
        sum += g * g;
        out.v[k] = half == 0 ? w - op.lr * (g / (sqrtf(sum) + op.eps)) : sum;
      }
      if (work[q] && lane_on) store_row<VEC>(rowp[q], out);
    }
  }
}

// CTA per long segment
template <int VEC, int LPR, int OPT>
__global__ void __launch_bounds__(kLongThreads)
fused_update_long_kernel(void* const* __restrict__ table_ptrs, void* const* __restrict__ state1_ptrs,
                         void* const* __restrict__ state2_ptrs, int T, int D,
                         const ptrec_feature_desc* __restrict__ feats, int F, int64_t B,
                         const uint32_t* __restrict__ sorted_keys, const int32_t* __restrict__ perm,
                         const int32_t* __restrict__ seg_start, const ptrec_segment_meta* __restrict__ seg_meta,
                         const float* __restrict__ grad_out, int64_t stride,
                         const float* __restrict__ bag_scale, OptParams op,
                         const int* __restrict__ long_count, const int32_t* __restrict__ long_list) {
  const int n_long = *long_count;
  if ((int)blockIdx.x >= n_long) return;  // the common case (no hot rows): exit before any setup
  __shared__ SlotMap s_map;
  constexpr int NSG = kLongThreads / LPR;
  __shared__ float s_part[NSG * LPR * VEC];
  build_slot_map(&s_map, feats, F, T, B);
  const int sg = threadIdx.x / LPR, lane = threadIdx.x % LPR;
  const bool lane_on = lane * VEC < D;
  for (int i = blockIdx.x; i < n_long; i += gridDim.x) {
    const int u = long_list[i];
    const int start = seg_start[u], end = seg_start[u + 1];
    const int t = seg_meta[u].table;
    RowVec<VEC> acc;
    acc.zero();
    for (int j0 = start + sg; j0 < end; j0 += NSG * 2) {
      RowVec<VEC> g0, g1;
      g0 = slot_grad<VEC>(&s_map, t, B, perm[j0], grad_out, stride, bag_scale, lane, lane_on);
      g1.zero();
      if (j0 + NSG < end)
        g1 = slot_grad<VEC>(&s_map, t, B, perm[j0 + NSG], grad_out, stride, bag_scale, lane, lane_on);
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
    if (sg == 0) {  // sub-warp 0 owns the row
      RowVec<VEC> tot;
#pragma unroll
      for (int k = 0; k < VEC; ++k) tot.v[k] = mine[k];
      const uint32_t key = sorted_keys[start];
      float* wrow = reinterpret_cast<float*>(table_ptrs[t]);
      float* s1 = state1_ptrs ? reinterpret_cast<float*>(state1_ptrs[t]) : nullptr;
      float* s2 = state2_ptrs ? reinterpret_cast<float*>(state2_ptrs[t]) : nullptr;
      RowState<VEC, OPT> st;
      st.load(wrow, s1, s2, (int64_t)key, op.row_stride, lane, lane_on);
      apply_update<VEC, LPR, OPT>(tot, st, wrow, s1, s2, (int64_t)key, D, lane, lane_on, op, nullptr);
    }
    __syncthreads();
  }
}

static int g_update_variant = 0;  // 0 = v1 for every optimizer; Adagrad: 1 = pair<SEGS 2, 8 CTAs/SM>, 2 = pair<4, 6>, 3 = pair<4, 4>
}  // namespace ptrec
extern "C" void ptrec_set_update_variant(int32_t v) { ptrec::g_update_variant = v < 0 ? 0 : (v > 3 ? 3 : v); }
extern "C" int32_t ptrec_update_variant(void) { return ptrec::g_update_variant; }
namespace ptrec {

template <int VEC, int LPR, int SEGS, int MINB>
static void launch_pair(void* const* table_ptrs, void* const* s1, int T, int D, const ptrec_feature_desc* feats, int F,
                        int64_t B, int64_t N, const int32_t* perm, const int32_t* seg_start,
                        const ptrec_segment_meta* seg_meta, const int32_t* n_seg, const float* grad_out, int64_t stride,
                        const float* bag_scale, const OptParams& op, int* long_count, int32_t* long_list, int sms,
                        cudaStream_t st) {
  constexpr int NSG = kUpdThreads / (2 * LPR);
  const int64_t rounds = ceil_div(N, NSG * SEGS);
  const int64_t cap = (int64_t)sms * MINB;
  const unsigned grid = (unsigned)(rounds < cap ? (rounds > 0 ? rounds : 1) : cap);
  fused_adagrad_pair_kernel<VEC, LPR, SEGS, MINB><<<grid, kUpdThreads, 0, st>>>(
      table_ptrs, s1, T, D, feats, F, B, perm, seg_start, seg_meta, n_seg, grad_out, stride, bag_scale, op, long_count,
      long_list);
}

template <int VEC, int LPR, int OPT>
static int launch_update(void* const* table_ptrs, void* const* s1, void* const* s2, int T, int D,
                         const ptrec_feature_desc* feats, int F, int64_t B, int64_t N,
                         const uint32_t* sorted_keys, const int32_t* perm, const int32_t* seg_start,
                         const ptrec_segment_meta* seg_meta, const int32_t* n_seg, const float* grad_out,
                         int64_t stride, const float* bag_scale, const OptParams& op, int* long_count,
                         int32_t* long_list, float* row_grad, cudaStream_t st) {
  constexpr int NSG = kUpdThreads / LPR;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t rounds = ceil_div(N, NSG * 4);
  // 4 CTAs per SM (register-limited): one wave, several rounds per CTA so that the metadata prefetch has a next round
  const unsigned grid = (unsigned)(rounds < (int64_t)sms * 4 ? (rounds > 0 ? rounds : 1) : (int64_t)sms * 4);
  if (OPT != kOptNone) PTREC_CUDA(cudaMemsetAsync(long_count, 0, sizeof(int), st));
  bool done = false;
  if constexpr (OPT == PTREC_OPT_ADAGRAD && 2 * LPR <= 32) {
    if (g_update_variant != 0) {
#define PTREC_PAIR(S, M) \
  launch_pair<VEC, LPR, S, M>(table_ptrs, s1, T, D, feats, F, B, N, perm, seg_start, seg_meta, n_seg, grad_out, stride, \
                              bag_scale, op, long_count, long_list, sms, st)
      if (g_update_variant == 1) PTREC_PAIR(2, 8);
      else if (g_update_variant == 2) PTREC_PAIR(4, 6);
      else PTREC_PAIR(4, 4);
#undef PTREC_PAIR
      PTREC_LAUNCH_CHECK("fused_adagrad_pair_kernel");
      done = true;
    }
  }
  if (!done) {
    fused_update_kernel<VEC, LPR, OPT><<<grid, kUpdThreads, 0, st>>>(
        table_ptrs, s1, s2, T, D, feats, F, B, sorted_keys, perm, seg_start, seg_meta, n_seg, grad_out,
        stride, bag_scale, op, long_count, long_list, row_grad);
    PTREC_LAUNCH_CHECK("fused_update_kernel");
  }
  if (OPT != kOptNone) {
    fused_update_long_kernel<VEC, LPR, OPT><<<sms * 2, kLongThreads, 0, st>>>(
        table_ptrs, s1, s2, T, D, feats, F, B, sorted_keys, perm, seg_start, seg_meta, grad_out, stride,
        bag_scale, op, long_count, long_list);
    PTREC_LAUNCH_CHECK("fused_update_long_kernel");
  }
  return PTREC_OK;
}

template <int OPT>
static int dispatch_D(void* const* table_ptrs, void* const* s1, void* const* s2, int T, int D,
                      const ptrec_feature_desc* feats, int F, int64_t B, int64_t N,
                      const uint32_t* sorted_keys, const int32_t* perm, const int32_t* seg_start,
                      const ptrec_segment_meta* seg_meta, const int32_t* n_seg, const float* grad_out,
                      int64_t stride, const float* bag_scale, const OptParams& op, int* long_count,
                      int32_t* long_list, float* row_grad, cudaStream_t st) {
#define PTREC_UPD(V, P) \
  return launch_update<V, P, OPT>(table_ptrs, s1, s2, T, D, feats, F, B, N, sorted_keys, perm, seg_start, \
                                  seg_meta, n_seg, grad_out, stride, bag_scale, op, long_count,        \
                                  long_list, row_grad, st)
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

extern "C" size_t ptrec_embedding_bwd_workspace_bytes(int64_t N, int32_t D) {
  (void)D;
  return align_up(256 + ((size_t)N / kLongSeg + 2) * sizeof(int32_t), 256);
}

static int64_t total_slots(const ptrec_feature_desc* feats_host, int F, int64_t B) {
  int64_t L = 0;
  for (int f = 0; f < F; ++f) L += feats_host[f].bag_len;
  return L * B;
}

static int check_common(int32_t T, int32_t D, int32_t dtype, int32_t F, const void* grad_out,
                        int64_t stride) {
  PTREC_CHECK_ARG(dtype == PTREC_F32, PTREC_EUNSUPPORTED, "bwd: only fp32 tables are built (dtype=%d)", dtype);
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
  int* long_count = reinterpret_cast<int*>(workspace);
  int32_t* long_list = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(workspace) + 256);
  cudaStream_t st = (cudaStream_t)stream;
#define PTREC_ARGS                                                                                      \
  table_ptrs, state1_ptrs, state2_ptrs, T, D, feats, F, B, N, sorted_keys, perm, seg_start, seg_meta, n_seg, \
      grad_out, grad_row_stride, bag_scale, op, long_count, long_list, nullptr, st
  switch (opt_host->kind) {
    case PTREC_OPT_SGD:
      return dispatch_D<PTREC_OPT_SGD>(PTREC_ARGS);
    case PTREC_OPT_ADAGRAD:
      PTREC_CHECK_ARG(state1_ptrs, PTREC_EINVAL, "bwd_fused: Adagrad needs state1");
      op.lr = (float)((double)opt_host->lr / (1.0 + (double)(step - 1) * (double)opt_host->lr_decay));
      return dispatch_D<PTREC_OPT_ADAGRAD>(PTREC_ARGS);
    case PTREC_OPT_ROWWISE_ADAGRAD:
      PTREC_CHECK_ARG(state1_ptrs, PTREC_EINVAL, "bwd_fused: row-wise Adagrad needs state1");
      op.lr = (float)((double)opt_host->lr / (1.0 + (double)(step - 1) * (double)opt_host->lr_decay));
      return dispatch_D<PTREC_OPT_ROWWISE_ADAGRAD>(PTREC_ARGS);
    case PTREC_OPT_LAZY_ADAM: {
      PTREC_CHECK_ARG(state1_ptrs && state2_ptrs, PTREC_EINVAL, "bwd_fused: lazy Adam needs state1 and state2");
      const double bc1 = 1.0 - pow((double)opt_host->beta1, (double)step);
      const double bc2 = 1.0 - pow((double)opt_host->beta2, (double)step);
      op.lr = (float)((double)opt_host->lr * sqrt(bc2) / bc1);
      return dispatch_D<PTREC_OPT_LAZY_ADAM>(PTREC_ARGS);
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
  return dispatch_D<kOptNone>(nullptr, nullptr, nullptr, T, D, feats, F, B, N, sorted_keys, perm, seg_start,
                              seg_meta, n_seg, grad_out, grad_row_stride, bag_scale, op, nullptr,
                              nullptr, row_grad, (cudaStream_t)stream);
}
