// C1: pack / unpack kernels either side of the row-wise-sharding all-to-all.
//
// Tables are sharded row-wise: owner(id) = id mod G, local_row = id div G.  Every rank sends, for each
// (destination, field), a FIXED-capacity list of local rows (capacity C >= expected B/G plus slack, unused
// slots = -1) so that the collective has equal splits and needs no host synchronisation.  The same slot
// numbering routes embedding rows back and row gradients forward:
//     ret_pos[f, b] = (owner * F + f) * C + slot      (or -1 if the list overflowed; flagged)
// Slots are assigned in batch order (stable), which keeps the owner's sort / segment sums bit-reproducible.
// Bound: these move 8-16 B per lookup; they are launch-latency bound at the configs' sizes.
//
// Peer mode (ptrec_a2a_*_peer): the lists and the gradient rows are stored DIRECTLY into the owner's receive buffers
// through peer pointers (NVLink / NVSwitch stores), in the layout the owner-side sort and fused update read:
//     owner_ids [f, src, slot]          owner_grads [(src*F + f)*C + slot, :]
// so the dispatch needs no all-to-all; the caller orders "all pushes done" before the owner consumes them with a
// barrier (peer_sync.cu); every list is written in full (lookups, then -1 up to its capacity), so the owner never
// resets anything.
#include <algorithm>

#include "common.cuh"

namespace ptrec {

constexpr int kPackThreads = 256;
constexpr int kPackItems = 8;
constexpr int kPackTile = kPackThreads * kPackItems;  // 2048 lookups per CTA
constexpr int kMaxRanks = 64;

// counts[f][tile][dest]
__global__ void __launch_bounds__(kPackThreads)
pack_count_kernel(const int64_t* __restrict__ ids, int64_t B, int G, int tiles, int* __restrict__ counts) {
  __shared__ int s_cnt[kMaxRanks];
  const int f = blockIdx.y, tile = blockIdx.x;
  if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t beg = (int64_t)tile * kPackTile, end = min(B, beg + kPackTile);
  for (int64_t b = beg + threadIdx.x; b < end; b += kPackThreads) {
    const int64_t id = ids[(int64_t)f * B + b];
    if (id >= 0) atomicAdd(&s_cnt[(int)(id % G)], 1);
  }
  __syncthreads();
  if (threadIdx.x < G) counts[((int64_t)f * tiles + tile) * G + threadIdx.x] = s_cnt[threadIdx.x];
}

// exclusive scan over tiles for each (f, dest); one CTA per field, one thread per destination
__global__ void pack_scan_kernel(int* __restrict__ counts, int G, int tiles, int C, int32_t* __restrict__ overflow) {
  const int f = blockIdx.x, d = threadIdx.x;
  if (d >= G) return;
  int run = 0;
  for (int t = 0; t < tiles; ++t) {
    int* p = counts + ((int64_t)f * tiles + t) * G + d;
    const int c = *p;
    *p = run;
    run += c;
  }
  if (run > C && overflow != nullptr) atomicMax(overflow, run);
}

// push mode: rows that no owner will write (negative id, overflowed list) are zeroed here, in the local output
struct ZeroFill {
  float* out[4];
  int64_t stride[4];
  int32_t dim[4];
  int32_t n;
};

__global__ void __launch_bounds__(kPackThreads)
pack_scatter_kernel(const int64_t* __restrict__ ids, int64_t B, int F, int G, int tiles, int C,
                    const int* __restrict__ bases, int64_t* __restrict__ send_ids, int32_t* __restrict__ ret_pos,
                    int64_t* const* __restrict__ peer_ids, int32_t* const* __restrict__ peer_b, ZeroFill zf,
                    int32_t* __restrict__ slot_b, int my_rank) {
  constexpr int NW = kPackThreads / 32;
  __shared__ int s_cnt[NW][kMaxRanks];
  __shared__ int s_base[kMaxRanks];
  __shared__ int s_total[kMaxRanks];
  __shared__ int64_t* s_peer[kMaxRanks];
  if (peer_ids != nullptr && threadIdx.x < G) s_peer[threadIdx.x] = peer_ids[threadIdx.x];
  const int f = blockIdx.y, tile = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NW * kMaxRanks; i += kPackThreads) (&s_cnt[0][0])[i] = 0;
  if (threadIdx.x < G) s_base[threadIdx.x] = bases[((int64_t)f * tiles + tile) * G + threadIdx.x];
  __syncthreads();
  const int64_t beg = (int64_t)tile * kPackTile, end = min(B, beg + kPackTile);
  int64_t id[kPackItems];
  int rank[kPackItems], dest[kPackItems];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kPackItems; ++i) {
    const int64_t b = beg + warp * (32 * kPackItems) + i * 32 + lane;
    const bool in = b < end;
    id[i] = in ? ids[(int64_t)f * B + b] : -1;
    const bool ok = id[i] >= 0;
    dest[i] = ok ? (int)(id[i] % G) : kMaxRanks;  // kMaxRanks = "nothing to send"
    const unsigned m = __match_any_sync(0xffffffffu, dest[i]);
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (ok) base = s_cnt[warp][dest[i]];
    __syncwarp();
    if (ok && lane == leader) s_cnt[warp][dest[i]] = base + __popc(m);
    __syncwarp();
    rank[i] = base + __popc(m & lt);
  }
  __syncthreads();
  if (threadIdx.x < G) {
    int run = s_base[threadIdx.x];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int c = s_cnt[w][threadIdx.x];
      s_cnt[w][threadIdx.x] = run;
      run += c;
    }
    s_total[threadIdx.x] = run;  // in the field's LAST tile: the final length of the list for this destination
  }
  __syncthreads();
  if (peer_ids != nullptr && tile == tiles - 1) {
    // peer mode: the owner's list is rewritten in full every step — the slots behind the last lookup read "no
    // lookup" (-1) — so the owner never has to reset it and a forward without a backward leaves nothing stale
    for (int d = 0; d < G; ++d) {
      int64_t* list = s_peer[d] + ((int64_t)f * G + my_rank) * C;
      for (int slot = min(s_total[d], C) + (int)threadIdx.x; slot < C; slot += kPackThreads) list[slot] = -1;
    }
  }
  if (slot_b != nullptr && tile == tiles - 1) {  // the local slot -> sample map is written in full as well
    for (int d = 0; d < G; ++d) {
      int32_t* list = slot_b + ((int64_t)d * F + f) * C;
      for (int slot = min(s_total[d], C) + (int)threadIdx.x; slot < C; slot += kPackThreads) list[slot] = -1;
    }
  }
#pragma unroll
  for (int i = 0; i < kPackItems; ++i) {
    const int64_t b = beg + warp * (32 * kPackItems) + i * 32 + lane;
    if (b < end) {
      int32_t pos = -1;
      if (id[i] >= 0) {
        const int slot = s_cnt[warp][dest[i]] + rank[i];
        if (slot < C) {
          pos = (int32_t)(((int64_t)dest[i] * F + f) * C + slot);
          if (slot_b != nullptr) slot_b[pos] = (int32_t)b;  // inverse of ret_pos: the sample behind slot `pos`
          if (peer_ids != nullptr) {
            const int64_t at = ((int64_t)f * G + my_rank) * C + slot;
            s_peer[dest[i]][at] = id[i] / G;
            if (peer_b != nullptr) peer_b[dest[i]][at] = (int32_t)b;  // push mode: where the owner must deliver the row
          } else {
            send_ids[pos] = id[i] / G;
          }
        }
      }
      if (pos < 0) {
        for (int k = 0; k < zf.n; ++k) {
          float* o = zf.out[k] + b * zf.stride[k] + (int64_t)f * zf.dim[k];
          for (int e = 0; e < zf.dim[k]; ++e) o[e] = 0.f;
        }
      }
      ret_pos[(int64_t)f * B + b] = pos;
    }
  }
}

// dst[pos[f,b], :] = scale * src[b, f, :]   (one sub-warp per row)
template <int VEC, int LPR>
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const float* __restrict__ src, int64_t src_row_stride, const int32_t* __restrict__ pos,
                    int64_t B, int F, int D, float scale, float* __restrict__ dst, int64_t dst_row_stride) {
  const int64_t g = ((int64_t)blockIdx.x * 256 + threadIdx.x) / LPR;
  const int lane = threadIdx.x % LPR;
  if (g >= B * F || lane * VEC >= D) return;
  const int64_t b = g / F;
  const int f = (int)(g - b * F);
  const int32_t p = pos[(int64_t)f * B + b];
  if (p < 0) return;
  RowVec<VEC> r = load_row_stream<VEC>(src + b * src_row_stride + (int64_t)f * D + lane * VEC);
  r.scale(scale);
  store_row<VEC>(dst + (int64_t)p * dst_row_stride + lane * VEC, r);
}

// peer_dst[owner][(my_rank*F*C + (pos mod F*C)) * stride + col : +D] = scale * src[b, f, :]
template <int VEC, int LPR>
__global__ void __launch_bounds__(256)
scatter_rows_peer_kernel(const float* __restrict__ src, int64_t src_row_stride, const int32_t* __restrict__ pos,
                         int64_t B, int F, int D, float scale, float* const* __restrict__ peer_dst,
                         int64_t dst_row_stride, int64_t dst_col, int FC, int my_rank) {
  const int64_t g = ((int64_t)blockIdx.x * 256 + threadIdx.x) / LPR;
  const int lane = threadIdx.x % LPR;
  if (g >= B * F || lane * VEC >= D) return;
  const int64_t b = g / F;
  const int f = (int)(g - b * F);
  const int32_t p = pos[(int64_t)f * B + b];
  if (p < 0) return;
  const int owner = p / FC;
  const int64_t row = (int64_t)my_rank * FC + (p - owner * FC);
  RowVec<VEC> r = load_row_stream<VEC>(src + b * src_row_stride + (int64_t)f * D + lane * VEC);
  r.scale(scale);
  store_row<VEC>(peer_dst[owner] + row * dst_row_stride + dst_col + lane * VEC, r);
}

// every width of a slot in ONE launch: the slot's lanes cover the float4 chunks of [width 0 | width 1 | ...], so the
// stores of one slot to its owner are adjacent (e.g. DeepFM: 64 B embedding gradient + 4 B first-order gradient)
constexpr int kMaxWidths = 4;
struct PeerWidths {
  const float* src[kMaxWidths];
  int64_t stride[kMaxWidths];
  int32_t dim[kMaxWidths];
  int32_t col[kMaxWidths];
  int32_t chunk0[kMaxWidths + 1];  // first float4 chunk of each width
  int32_t n;
};

template <int LPR>
__global__ void __launch_bounds__(256)
scatter_rows_peer_multi_kernel(PeerWidths w, const int32_t* __restrict__ pos, int64_t B, int F, float scale,
                               float* const* __restrict__ peer_dst, int64_t dst_row_stride, int FC, int my_rank) {
  const int64_t g = ((int64_t)blockIdx.x * 256 + threadIdx.x) / LPR;
  const int lane = threadIdx.x % LPR;
  if (g >= B * F || lane >= w.chunk0[w.n]) return;
  const int64_t b = g / F;
  const int f = (int)(g - b * F);
  const int32_t p = pos[(int64_t)f * B + b];
  if (p < 0) return;
  int k = 0;
  while (k + 1 < w.n && lane >= w.chunk0[k + 1]) ++k;
  const int c = (lane - w.chunk0[k]) * 4;           // first element of this lane's chunk inside width k
  const int D = w.dim[k];
  const int owner = p / FC;
  const int64_t row = (int64_t)my_rank * FC + (p - owner * FC);
  const float* s = w.src[k] + b * w.stride[k] + (int64_t)f * D + c;
  float* d = peer_dst[owner] + row * dst_row_stride + w.col[k] + c;
  if (c + 3 < D && (D & 3) == 0) {
    float4 v = *reinterpret_cast<const float4*>(s);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    *reinterpret_cast<float4*>(d) = v;
  } else {
    for (int i = 0; i < 4 && c + i < D; ++i) d[i] = s[i] * scale;
  }
}

// Same dispatch in DESTINATION order: one sub-warp per (owner, field, slot) of this rank's lists, `slot_b` (written by
// the pack kernels) naming the sample behind the slot.  Neighbouring sub-warps store neighbouring slots of the same
// owner, so the NVLink stores are long contiguous runs (whole slots, padding included) instead of isolated 64-byte
// rows — scattered 16-byte pieces ran at ~250 GB/s (profiles/r2_kernels_step_n2_push_first.txt).
template <int LPR>
__global__ void __launch_bounds__(256)
scatter_rows_peer_ordered_kernel(PeerWidths w, const int32_t* __restrict__ slot_b, int F, int C, int G, float scale,
                                 float* const* __restrict__ peer_dst, int64_t dst_row_stride, int my_rank) {
  const int lane = threadIdx.x % LPR;
  const int64_t groups = (int64_t)gridDim.x * (256 / LPR);
  const int64_t n_slots = (int64_t)G * F * C;
  const int n_chunks = (int)(dst_row_stride / 4);  // the whole slot is written: the pad lanes store zeros
  if (lane >= n_chunks) return;
  int k = 0;
  while (k + 1 < w.n && lane >= w.chunk0[k + 1]) ++k;
  const bool pad = lane >= w.chunk0[w.n];
  const int c = (lane - w.chunk0[k]) * 4;
  const int D = w.dim[k];
  for (int64_t s0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) / LPR; s0 < n_slots; s0 += groups * 4) {
    float4 v[4];
    float* dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t s = s0 + u * groups;
      dst[u] = nullptr;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s < n_slots) {
        const int b = slot_b[s];
        if (b >= 0) {
          const int owner = (int)(s / ((int64_t)F * C));
          const int64_t rel = s - (int64_t)owner * F * C;  // f * C + slot
          const int f = (int)(rel / C);
          dst[u] = peer_dst[owner] + ((int64_t)my_rank * F * C + rel) * dst_row_stride + lane * 4;
          if (!pad) {
            const float* src = w.src[k] + (int64_t)b * w.stride[k] + (int64_t)f * D + c;
            if (c + 3 < D && (D & 3) == 0) {
              v[u] = *reinterpret_cast<const float4*>(src);
            } else {
              v[u].x = src[0];
              if (c + 1 < D) v[u].y = src[1];
              if (c + 2 < D) v[u].z = src[2];
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (dst[u] == nullptr) continue;
      v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
      *reinterpret_cast<float4*>(dst[u]) = v[u];
    }
  }
}

// ---- push mode forward: the OWNER gathers the rows of the lists it received and stores each one straight into the
// requester's output row over NVLink / NVSwitch — the fused form of  owner-side gather -> all-to-all(rows) -> gather by
// slot.  Random accesses stay local (the owner's own shard); remote traffic is contiguous row stores into a buffer of
// batch x F x D floats.  Every width of a slot in one launch (lanes cover the float4 chunks of [width 0 | width 1 | ...]),
// kPushUnroll slots in flight per sub-warp.
constexpr int kPushUnroll = 4;
struct PushArgs {
  const void* const* table_ptrs[kMaxWidths];  // [T] shard bases of width k (device array)
  float* const* peer_out[kMaxWidths];         // [G] the requesters' output buffers of width k (device array)
  int64_t row_stride[kMaxWidths];
  int64_t out_stride[kMaxWidths];
  int32_t dim[kMaxWidths];
  int32_t chunk0[kMaxWidths + 1];
  int32_t n;
};

template <int LPR>
__global__ void __launch_bounds__(256)
gather_push_kernel(PushArgs a, const int64_t* __restrict__ recv_ids, const int32_t* __restrict__ recv_b,
                   const int64_t* __restrict__ shard_rows, int F, int G, int C, int32_t* err_flag) {
  const int lane = threadIdx.x % LPR;
  const int64_t groups = (int64_t)gridDim.x * (256 / LPR);
  const int64_t g0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) / LPR;
  const int64_t n_slots = (int64_t)F * G * C;
  if (lane >= a.chunk0[a.n]) return;
  int k = 0;
  while (k + 1 < a.n && lane >= a.chunk0[k + 1]) ++k;
  const int c = (lane - a.chunk0[k]) * 4;  // first element of this lane's chunk inside width k
  const int D = a.dim[k];
  const bool vec = (c + 3 < D) && ((D & 3) == 0);
  for (int64_t s0 = g0; s0 < n_slots; s0 += groups * kPushUnroll) {
    float4 v[kPushUnroll];
    float* dst[kPushUnroll];
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u) {
      const int64_t s = s0 + u * groups;
      dst[u] = nullptr;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s < n_slots) {
        const int64_t id = recv_ids[s];
        if (id >= 0) {
          const int f = (int)(s / ((int64_t)G * C));
          const int src = (int)((s - (int64_t)f * G * C) / C);
          if (id < shard_rows[f]) {
            const float* row = reinterpret_cast<const float*>(a.table_ptrs[k][f]) + id * a.row_stride[k] + c;
            if (vec) {
              v[u] = ldg_stream_f4(row);
            } else {
              v[u].x = row[0];
              if (c + 1 < D) v[u].y = row[1];
              if (c + 2 < D) v[u].z = row[2];
            }
            dst[u] = a.peer_out[k][src] + (int64_t)recv_b[s] * a.out_stride[k] + (int64_t)f * D + c;
          } else if (err_flag != nullptr && lane == 0) {
            *err_flag = 1;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u) {
      if (dst[u] == nullptr) continue;
      if (vec) {
        *reinterpret_cast<float4*>(dst[u]) = v[u];
      } else {
        dst[u][0] = v[u].x;
        if (c + 1 < D) dst[u][1] = v[u].y;
        if (c + 2 < D) dst[u][2] = v[u].z;
      }
    }
  }
}

}  // namespace ptrec

using namespace ptrec;

extern "C" int ptrec_gather_push(const void* const* const* table_ptrs, float* const* const* peer_out,
                                 const int64_t* row_strides, const int64_t* out_row_strides, const int32_t* dims,
                                 int32_t n_widths, const int64_t* recv_ids, const int32_t* recv_b,
                                 const int64_t* shard_rows, int32_t F, int32_t G, int32_t C, int32_t* err_flag,
                                 void* stream) {
  PTREC_CHECK_ARG(table_ptrs && peer_out && row_strides && out_row_strides && dims && recv_ids && recv_b && shard_rows,
                  PTREC_EINVAL, "gather_push: null pointer");
  PTREC_CHECK_ARG(n_widths >= 1 && n_widths <= kMaxWidths && F >= 1 && G >= 1 && G <= kMaxRanks && C >= 1, PTREC_EINVAL,
                  "gather_push: bad sizes widths=%d F=%d G=%d C=%d", n_widths, F, G, C);
  PushArgs a;
  int chunks = 0;
  for (int k = 0; k < kMaxWidths; ++k) {
    const bool on = k < n_widths;
    a.table_ptrs[k] = on ? table_ptrs[k] : nullptr;
    a.peer_out[k] = on ? peer_out[k] : nullptr;
    a.row_stride[k] = on ? row_strides[k] : 0;
    a.out_stride[k] = on ? out_row_strides[k] : 0;
    a.dim[k] = on ? dims[k] : 0;
    a.chunk0[k] = chunks;
    if (on) {
      PTREC_CHECK_ARG(dims[k] >= 1 && dims[k] <= 128, PTREC_EINVAL, "gather_push: width %d", k);
      PTREC_CHECK_ARG((dims[k] & 3) != 0 || (row_strides[k] % 4 == 0 && out_row_strides[k] % 4 == 0), PTREC_EALIGN,
                      "gather_push: width %d strides must be multiples of 4 floats", k);
      chunks += (dims[k] + 3) / 4;
    }
  }
  a.chunk0[kMaxWidths] = chunks;
  for (int k = n_widths; k <= kMaxWidths; ++k) a.chunk0[k] = chunks;
  a.n = n_widths;
  PTREC_CHECK_ARG(chunks <= 32, PTREC_EUNSUPPORTED, "gather_push: slot wider than 128 floats");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_slots = (int64_t)F * G * C;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define PTREC_PUSH(P)                                                                                          \
  {                                                                                                            \
    const int64_t want = ceil_div(n_slots * P, 256 * kPushUnroll);                                             \
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));            \
    gather_push_kernel<P><<<grid, 256, 0, st>>>(a, recv_ids, recv_b, shard_rows, F, G, C, err_flag);            \
    PTREC_LAUNCH_CHECK("gather_push_kernel");                                                                  \
    return PTREC_OK;                                                                                           \
  }
  if (chunks <= 1) PTREC_PUSH(1)
  if (chunks <= 2) PTREC_PUSH(2)
  if (chunks <= 4) PTREC_PUSH(4)
  if (chunks <= 8) PTREC_PUSH(8)
  if (chunks <= 16) PTREC_PUSH(16)
  PTREC_PUSH(32)
#undef PTREC_PUSH
}

extern "C" int ptrec_a2a_scatter_rows_peer_ordered(const float* const* srcs, const int64_t* src_row_strides,
                                                   const int32_t* dims, const int64_t* dst_cols, int32_t n_widths,
                                                   const int32_t* slot_b, int32_t F, float scale,
                                                   float* const* peer_dst, int64_t dst_row_stride, int32_t C, int32_t G,
                                                   int32_t my_rank, void* stream) {
  PTREC_CHECK_ARG(srcs && src_row_strides && dims && dst_cols && slot_b && peer_dst, PTREC_EINVAL,
                  "a2a_scatter_rows_peer_ordered: null pointer");
  PTREC_CHECK_ARG(n_widths >= 1 && n_widths <= kMaxWidths, PTREC_EINVAL, "a2a_scatter_rows_peer_ordered: 1..%d widths",
                  kMaxWidths);
  PTREC_CHECK_ARG(G >= 1 && G <= kMaxRanks && my_rank >= 0 && my_rank < G && C >= 1 && F >= 1, PTREC_EINVAL,
                  "a2a_scatter_rows_peer_ordered: bad G=%d rank=%d C=%d", G, my_rank, C);
  PTREC_CHECK_ARG(dst_row_stride % 4 == 0 && dst_row_stride <= 128, PTREC_EALIGN,
                  "a2a_scatter_rows_peer_ordered: slot width must be a multiple of 4 floats, at most 128");
  PeerWidths w;
  w.n = n_widths;
  int chunks = 0;
  for (int k = 0; k < n_widths; ++k) {
    PTREC_CHECK_ARG(srcs[k] && dims[k] >= 1 && dims[k] <= 128, PTREC_EINVAL, "a2a_scatter_rows_peer_ordered: width %d", k);
    // the lanes of a slot cover its columns 4 by 4 in order: width k must start where width k-1's chunks end
    PTREC_CHECK_ARG(dst_cols[k] == (int64_t)chunks * 4, PTREC_EINVAL,
                    "a2a_scatter_rows_peer_ordered: widths must be packed at 16-byte boundaries in slot order");
    const bool vec = (dims[k] & 3) == 0;
    PTREC_CHECK_ARG(!vec || (((uintptr_t)srcs[k] & 15) == 0 && src_row_strides[k] % 4 == 0), PTREC_EALIGN,
                    "a2a_scatter_rows_peer_ordered: width %d misaligned", k);
    w.src[k] = srcs[k];
    w.stride[k] = src_row_strides[k];
    w.dim[k] = dims[k];
    w.col[k] = (int32_t)dst_cols[k];
    w.chunk0[k] = chunks;
    chunks += (dims[k] + 3) / 4;
  }
  w.chunk0[n_widths] = chunks;
  for (int k = n_widths; k < kMaxWidths; ++k) { w.src[k] = nullptr; w.stride[k] = 0; w.dim[k] = 0; w.col[k] = 0; }
  for (int k = n_widths + 1; k <= kMaxWidths; ++k) w.chunk0[k] = chunks;
  PTREC_CHECK_ARG((int64_t)chunks * 4 <= dst_row_stride, PTREC_EINVAL, "a2a_scatter_rows_peer_ordered: slot too narrow");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_slots = (int64_t)G * F * C;
  const int lanes = (int)(dst_row_stride / 4);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define PTREC_SCO(P)                                                                                          \
  {                                                                                                           \
    const int64_t want = ceil_div(n_slots * P, 256 * 4);                                                      \
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));           \
    scatter_rows_peer_ordered_kernel<P><<<grid, 256, 0, st>>>(w, slot_b, F, C, G, scale, peer_dst,            \
                                                              dst_row_stride, my_rank);                       \
    PTREC_LAUNCH_CHECK("scatter_rows_peer_ordered_kernel");                                                   \
    return PTREC_OK;                                                                                          \
  }
  if (lanes <= 1) PTREC_SCO(1)
  if (lanes <= 2) PTREC_SCO(2)
  if (lanes <= 4) PTREC_SCO(4)
  if (lanes <= 8) PTREC_SCO(8)
  if (lanes <= 16) PTREC_SCO(16)
  PTREC_SCO(32)
#undef PTREC_SCO
}

extern "C" size_t ptrec_a2a_pack_workspace_bytes(int64_t B, int32_t F, int32_t G) {
  const int64_t tiles = ceil_div(B, kPackTile);
  return align_up((size_t)(F * tiles * G) * sizeof(int) + 16, 256);
}

static int pack_impl(const int64_t* ids, int64_t B, int32_t F, int32_t G, int32_t C, int64_t* send_ids,
                     int64_t* const* peer_ids, int32_t* const* peer_b, const ZeroFill& zf, int32_t* slot_b,
                     int32_t my_rank, int32_t* ret_pos, int32_t* overflow, void* workspace, size_t workspace_bytes,
                     void* stream) {
  PTREC_CHECK_ARG(ids && (send_ids || peer_ids) && ret_pos && workspace, PTREC_EINVAL, "a2a_pack: null pointer");
  PTREC_CHECK_ARG(my_rank >= 0 && my_rank < G, PTREC_EINVAL, "a2a_pack: rank %d outside [0, %d)", my_rank, G);
  PTREC_CHECK_ARG(B >= 0 && F >= 1 && F <= 65535 && G >= 1 && G <= kMaxRanks && C >= 1, PTREC_EINVAL,
                  "a2a_pack: bad sizes B=%lld F=%d G=%d C=%d", (long long)B, F, G, C);
  PTREC_CHECK_ARG((int64_t)G * F * C < (int64_t)0x7fffffff, PTREC_EUNSUPPORTED, "a2a_pack: G*F*C must be < 2^31");
  PTREC_CHECK_ARG(workspace_bytes >= ptrec_a2a_pack_workspace_bytes(B, F, G), PTREC_EWORKSPACE, "a2a_pack: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (send_ids != nullptr)
    PTREC_CUDA(cudaMemsetAsync(send_ids, 0xFF, (size_t)G * F * C * sizeof(int64_t), st));  // -1 = empty slot
  if (B == 0) return PTREC_OK;
  const int tiles = (int)ceil_div(B, kPackTile);
  int* counts = reinterpret_cast<int*>(workspace);
  dim3 grid(tiles, F);
  pack_count_kernel<<<grid, kPackThreads, 0, st>>>(ids, B, G, tiles, counts);
  PTREC_LAUNCH_CHECK("pack_count_kernel");
  pack_scan_kernel<<<F, kMaxRanks, 0, st>>>(counts, G, tiles, C, overflow);
  PTREC_LAUNCH_CHECK("pack_scan_kernel");
  pack_scatter_kernel<<<grid, kPackThreads, 0, st>>>(ids, B, F, G, tiles, C, counts, send_ids, ret_pos, peer_ids,
                                                     peer_b, zf, slot_b, my_rank);
  PTREC_LAUNCH_CHECK("pack_scatter_kernel");
  return PTREC_OK;
}

extern "C" int ptrec_a2a_pack_by_owner(const int64_t* ids, int64_t B, int32_t F, int32_t G, int32_t C,
                                       int64_t* send_ids, int32_t* ret_pos, int32_t* overflow, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  PTREC_CHECK_ARG(send_ids, PTREC_EINVAL, "a2a_pack: null send_ids");
  ZeroFill zf{};
  return pack_impl(ids, B, F, G, C, send_ids, nullptr, nullptr, zf, nullptr, 0, ret_pos, overflow, workspace,
                   workspace_bytes, stream);
}

extern "C" int ptrec_a2a_pack_by_owner_peer(const int64_t* ids, int64_t B, int32_t F, int32_t G, int32_t C,
                                            int32_t my_rank, int64_t* const* peer_ids, int32_t* ret_pos,
                                            int32_t* slot_b, int32_t* overflow, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  PTREC_CHECK_ARG(peer_ids, PTREC_EINVAL, "a2a_pack_peer: null peer pointer array");
  ZeroFill zf{};
  return pack_impl(ids, B, F, G, C, nullptr, peer_ids, nullptr, zf, slot_b, my_rank, ret_pos, overflow, workspace,
                   workspace_bytes, stream);
}

extern "C" int ptrec_a2a_pack_by_owner_push(const int64_t* ids, int64_t B, int32_t F, int32_t G, int32_t C,
                                            int32_t my_rank, int64_t* const* peer_ids, int32_t* const* peer_b,
                                            float* const* local_out, const int64_t* out_row_strides,
                                            const int32_t* dims, int32_t n_widths, int32_t* ret_pos,
                                            int32_t* slot_b, int32_t* overflow, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  PTREC_CHECK_ARG(peer_ids && peer_b && local_out && out_row_strides && dims, PTREC_EINVAL,
                  "a2a_pack_push: null pointer");
  PTREC_CHECK_ARG(n_widths >= 1 && n_widths <= 4, PTREC_EINVAL, "a2a_pack_push: 1..4 widths");
  ZeroFill zf{};
  zf.n = n_widths;
  for (int k = 0; k < n_widths; ++k) {
    zf.out[k] = local_out[k];
    zf.stride[k] = out_row_strides[k];
    zf.dim[k] = dims[k];
  }
  return pack_impl(ids, B, F, G, C, nullptr, peer_ids, peer_b, zf, slot_b, my_rank, ret_pos, overflow, workspace,
                   workspace_bytes, stream);
}

extern "C" int ptrec_a2a_scatter_rows(const float* src, int64_t src_row_stride, const int32_t* ret_pos, int64_t B,
                                      int32_t F, int32_t D, float scale, float* dst, int64_t dst_row_stride,
                                      void* stream) {
  PTREC_CHECK_ARG(src && ret_pos && dst, PTREC_EINVAL, "a2a_scatter_rows: null pointer");
  const bool d_ok = D == 1 || D == 2 || (D >= 4 && D <= 128 && D % 4 == 0);
  PTREC_CHECK_ARG(d_ok, PTREC_EUNSUPPORTED, "a2a_scatter_rows: D=%d unsupported", D);
  const int vec = D >= 4 ? 4 : D;
  PTREC_CHECK_ARG(((uintptr_t)src % (vec * 4)) == 0 && ((uintptr_t)dst % (vec * 4)) == 0 && src_row_stride % vec == 0 &&
                      dst_row_stride >= D && dst_row_stride % vec == 0,
                  PTREC_EALIGN, "a2a_scatter_rows: misaligned");
  if (B == 0) return PTREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define PTREC_SC(V, P)                                                                                   \
  {                                                                                                      \
    const unsigned grid = (unsigned)ceil_div(B * F * P, 256);                                            \
    scatter_rows_kernel<V, P><<<grid, 256, 0, st>>>(src, src_row_stride, ret_pos, B, F, D, scale, dst,   \
                                                    dst_row_stride);                                     \
    PTREC_LAUNCH_CHECK("scatter_rows_kernel");                                                           \
    return PTREC_OK;                                                                                     \
  }
  if (D == 1) PTREC_SC(1, 1)
  if (D == 2) PTREC_SC(2, 1)
  const int lanes = D / 4;
  if (lanes <= 1) PTREC_SC(4, 1)
  if (lanes <= 2) PTREC_SC(4, 2)
  if (lanes <= 4) PTREC_SC(4, 4)
  if (lanes <= 8) PTREC_SC(4, 8)
  if (lanes <= 16) PTREC_SC(4, 16)
  PTREC_SC(4, 32)
#undef PTREC_SC
}

extern "C" int ptrec_a2a_scatter_rows_peer(const float* src, int64_t src_row_stride, const int32_t* ret_pos,
                                           int64_t B, int32_t F, int32_t D, float scale, float* const* peer_dst,
                                           int64_t dst_row_stride, int64_t dst_col, int32_t C, int32_t G,
                                           int32_t my_rank, void* stream) {
  PTREC_CHECK_ARG(src && ret_pos && peer_dst, PTREC_EINVAL, "a2a_scatter_rows_peer: null pointer");
  PTREC_CHECK_ARG(G >= 1 && G <= kMaxRanks && my_rank >= 0 && my_rank < G && C >= 1, PTREC_EINVAL,
                  "a2a_scatter_rows_peer: bad G=%d rank=%d C=%d", G, my_rank, C);
  const bool d_ok = D == 1 || D == 2 || (D >= 4 && D <= 128 && D % 4 == 0);
  PTREC_CHECK_ARG(d_ok, PTREC_EUNSUPPORTED, "a2a_scatter_rows_peer: D=%d unsupported", D);
  const int vec = D >= 4 ? 4 : D;
  PTREC_CHECK_ARG(((uintptr_t)src % (vec * 4)) == 0 && src_row_stride % vec == 0 && dst_row_stride >= D &&
                      dst_row_stride % vec == 0 && dst_col % vec == 0,
                  PTREC_EALIGN, "a2a_scatter_rows_peer: misaligned");
  if (B == 0) return PTREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int FC = F * C;
#define PTREC_SC(V, P)                                                                                         \
  {                                                                                                            \
    const unsigned grid = (unsigned)ceil_div(B * F * P, 256);                                                  \
    scatter_rows_peer_kernel<V, P><<<grid, 256, 0, st>>>(src, src_row_stride, ret_pos, B, F, D, scale,         \
                                                         peer_dst, dst_row_stride, dst_col, FC, my_rank);      \
    PTREC_LAUNCH_CHECK("scatter_rows_peer_kernel");                                                            \
    return PTREC_OK;                                                                                           \
  }
  if (D == 1) PTREC_SC(1, 1)
  if (D == 2) PTREC_SC(2, 1)
  const int lanes = D / 4;
  if (lanes <= 1) PTREC_SC(4, 1)
  if (lanes <= 2) PTREC_SC(4, 2)
  if (lanes <= 4) PTREC_SC(4, 4)
  if (lanes <= 8) PTREC_SC(4, 8)
  if (lanes <= 16) PTREC_SC(4, 16)
  PTREC_SC(4, 32)
#undef PTREC_SC
}

extern "C" int ptrec_a2a_scatter_rows_peer_multi(const float* const* srcs, const int64_t* src_row_strides,
                                                 const int32_t* dims, const int64_t* dst_cols, int32_t n_widths,
                                                 const int32_t* ret_pos, int64_t B, int32_t F, float scale,
                                                 float* const* peer_dst, int64_t dst_row_stride, int32_t C, int32_t G,
                                                 int32_t my_rank, void* stream) {
  PTREC_CHECK_ARG(srcs && src_row_strides && dims && dst_cols && ret_pos && peer_dst, PTREC_EINVAL,
                  "a2a_scatter_rows_peer_multi: null pointer");
  PTREC_CHECK_ARG(n_widths >= 1 && n_widths <= kMaxWidths, PTREC_EINVAL, "a2a_scatter_rows_peer_multi: 1..%d widths",
                  kMaxWidths);
  PTREC_CHECK_ARG(G >= 1 && G <= kMaxRanks && my_rank >= 0 && my_rank < G && C >= 1, PTREC_EINVAL,
                  "a2a_scatter_rows_peer_multi: bad G=%d rank=%d C=%d", G, my_rank, C);
  PeerWidths w;
  w.n = n_widths;
  int chunks = 0;
  for (int k = 0; k < n_widths; ++k) {
    PTREC_CHECK_ARG(srcs[k] && dims[k] >= 1 && dims[k] <= 128, PTREC_EINVAL, "a2a_scatter_rows_peer_multi: width %d", k);
    const bool vec = (dims[k] & 3) == 0;
    PTREC_CHECK_ARG(!vec || (((uintptr_t)srcs[k] & 15) == 0 && src_row_strides[k] % 4 == 0 && dst_cols[k] % 4 == 0 &&
                             dst_row_stride % 4 == 0),
                    PTREC_EALIGN, "a2a_scatter_rows_peer_multi: width %d misaligned", k);
    w.src[k] = srcs[k];
    w.stride[k] = src_row_strides[k];
    w.dim[k] = dims[k];
    w.col[k] = (int32_t)dst_cols[k];
    w.chunk0[k] = chunks;
    chunks += (dims[k] + 3) / 4;
  }
  w.chunk0[n_widths] = chunks;
  for (int k = n_widths; k < kMaxWidths; ++k) { w.src[k] = nullptr; w.stride[k] = 0; w.dim[k] = 0; w.col[k] = 0; }
  for (int k = n_widths + 1; k <= kMaxWidths; ++k) w.chunk0[k] = chunks;
  PTREC_CHECK_ARG(chunks <= 32, PTREC_EUNSUPPORTED, "a2a_scatter_rows_peer_multi: slot wider than 128 floats");
  if (B == 0) return PTREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int FC = F * C;
#define PTREC_SCM(P)                                                                                            \
  {                                                                                                             \
    const unsigned grid = (unsigned)ceil_div(B * F * P, 256);                                                   \
    scatter_rows_peer_multi_kernel<P><<<grid, 256, 0, st>>>(w, ret_pos, B, F, scale, peer_dst, dst_row_stride,  \
                                                            FC, my_rank);                                      \
    PTREC_LAUNCH_CHECK("scatter_rows_peer_multi_kernel");                                                       \
    return PTREC_OK;                                                                                            \
  }
  if (chunks <= 1) PTREC_SCM(1)
  if (chunks <= 2) PTREC_SCM(2)
  if (chunks <= 4) PTREC_SCM(4)
  if (chunks <= 8) PTREC_SCM(8)
  if (chunks <= 16) PTREC_SCM(16)
  PTREC_SCM(32)
#undef PTREC_SCM
}
