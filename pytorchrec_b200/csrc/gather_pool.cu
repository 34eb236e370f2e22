// K1: multi-table embedding gather + sum/mean/sqrtn pooling (forward), and index_prep.
//
// Bound: HBM.  Algorithmic bytes per launch = lookups*8 (ids) + lookups*D*4 (rows) + bags*D*4 (out).
// Layout: tables are independent [rows, D] fp32 (or bf16, dtype PTREC_BF16: same kernels, rows read as bf16 and
// widened exactly; output and pooling arithmetic stay fp32) row-major arrays reached through a device pointer
// array; ids are feature-major padded matrices; output rows are sample-major so that the DNN input
// is a view of them.
//
//  * one-hot fields (bag_len 1, no mask): a sub-warp of D/4 lanes owns one (sample, field) lookup and
//    issues ONE 128-bit `ld.global.nc.L1::no_allocate` per lane, 4 lookups in flight per lane.  A CTA
//    owns a tile of 16 samples x all one-hot fields; the tile's index lists are staged into shared
//    memory with 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) completing on an mbarrier, so that the
//    stride-B id reads become 128-byte contiguous transactions.
//  * pooled bags: warp per bag, sub-warps stride over the bag, deterministic xor-tree reduction;
//    4 consecutive bags' index lists are staged with one bulk copy.
#include "common.cuh"
#include "scan.cuh"

namespace ptrec {

struct FeatSel {
  int32_t n;
  int16_t idx[kMaxFeatures];
};

constexpr int kOneHotTileB = 16;     // samples per CTA tile
constexpr int kOneHotThreads = 128;  // 4 warps
constexpr int kUnroll = 4;

// ------------------------------------------------------------------------------------ one-hot
// SHARDED: the tables are row-wise shards owned by `shards` GPUs of one NVLink domain (owner = id mod shards,
// local row = id div shards); table_ptrs is [T][shards] and may hold PEER pointers, so the same 128-bit row loads
// travel over NVLink / NVSwitch and the lookup needs no collective.  table_rows stays the GLOBAL row count.
template <int VEC, int LPR, bool SHARDED, bool WB>
__global__ void __launch_bounds__(kOneHotThreads)
gather_onehot_kernel(const void* const* __restrict__ table_ptrs, const int64_t* __restrict__ table_rows,
                     int D, int64_t row_stride, const ptrec_feature_desc* __restrict__ feats, FeatSel sel,
                     const int64_t* __restrict__ ids, int64_t B, float* __restrict__ out,
                     int64_t out_row_stride, int32_t* err_flag, int use_bulk, int shards) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: ids tile [n_sel][kOneHotTileB] int64 | per-selected-feature {table ptr x shards, rows, out_col, flags}
  int64_t* s_ids = reinterpret_cast<int64_t*>(smem_raw);
  const int nsel = sel.n;
  const int nsh = SHARDED ? shards : 1;
  const void** s_tab = reinterpret_cast<const void**>(s_ids + (size_t)nsel * kOneHotTileB);
  int64_t* s_rows = reinterpret_cast<int64_t*>(s_tab + (size_t)nsel * nsh);
  int64_t* s_col = s_rows + nsel;
  int64_t* s_flag = s_col + nsel;
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * kOneHotTileB;
  const int nb = (int)min((int64_t)kOneHotTileB, B - b0);

  if (use_bulk) {
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
  }
  for (int s = tid; s < nsel; s += blockDim.x) {
    const ptrec_feature_desc fd = feats[sel.idx[s]];
    for (int g = 0; g < nsh; ++g)
      s_tab[(size_t)s * nsh + g] = table_ptrs[(size_t)fd.table * nsh + g];
    s_rows[s] = table_rows[fd.table];
    s_col[s] = fd.out_col;
    s_flag[s] = fd.flags;
    const int64_t* src = ids + fd.id_base * B + b0;
    if (use_bulk) {
      bulk_g2s(s_ids + (size_t)s * kOneHotTileB, src, (uint32_t)nb * 8u, &s_bar);
    } else {
      for (int j = 0; j < nb; ++j) s_ids[(size_t)s * kOneHotTileB + j] = src[j];
    }
  }
  if (use_bulk) {
    if (tid == 0) mbar_arrive_expect_tx(&s_bar, (uint32_t)nsel * (uint32_t)nb * 8u);
    __syncthreads();  // publishes s_tab/s_rows/s_col
    mbar_wait(&s_bar, 0);
  } else {
    __syncthreads();
  }

  constexpr int NSG = kOneHotThreads / LPR;  // sub-warps per CTA
  const int sg = tid / LPR;
  const int lane = tid % LPR;
  const bool lane_on = lane * VEC < D;
  const int total = nb * nsel;

  for (int i0 = sg; i0 < total; i0 += NSG * kUnroll) {
    RowVec<VEC> r[kUnroll];
    float* dst[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = i0 + u * NSG;
      dst[u] = nullptr;
      r[u].zero();
      if (i < total) {
        const int bl = i / nsel, s = i - bl * nsel;
        const int64_t id = s_ids[(size_t)s * kOneHotTileB + bl];
        dst[u] = out + (b0 + bl) * out_row_stride + s_col[s] + lane * VEC;
        if ((uint64_t)id < (uint64_t)s_rows[s]) {
          const void* tab;
          int64_t row;
          if (SHARDED) {
            int own;
            if ((uint64_t)id >> 32) {
              row = id / nsh;
              own = (int)(id - row * nsh);
            } else {  // 32-bit divide: the common case
              const uint32_t q = (uint32_t)id / (uint32_t)nsh;
              own = (int)((uint32_t)id - q * (uint32_t)nsh);
              row = q;
            }
            tab = s_tab[(size_t)s * nsh + own];
          } else {
            tab = s_tab[s];
            row = id;
          }
          if (lane_on) r[u] = load_table_row<VEC, WB, true>(tab, row * row_stride + lane * VEC);
        } else if (err_flag != nullptr && lane == 0 && !(id < 0 && (s_flag[s] & PTREC_FEAT_NEG_IS_PAD))) {
          *err_flag = 1;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (dst[u] != nullptr && lane_on) store_row<VEC>(dst[u], r[u]);
    }
  }
}

// ------------------------------------------------------------------------------------ pooled bags
constexpr int kBagThreads = 128;  // 4 warps = 4 bags per CTA
constexpr int kBagsPerCta = kBagThreads / 32;
constexpr int kBagStageMaxL = 2048;  // 4 * 2048 * 8 B = 64 KB of shared memory at most

template <int VEC, int LPR, bool WB>
__global__ void __launch_bounds__(kBagThreads)
gather_bag_kernel(const void* const* __restrict__ table_ptrs, const int64_t* __restrict__ table_rows,
                  int D, int64_t row_stride, const ptrec_feature_desc* __restrict__ feats, FeatSel sel,
                  const int64_t* __restrict__ ids, const int32_t* __restrict__ lens, int64_t B,
                  float* __restrict__ out, int64_t out_row_stride, float* __restrict__ bag_scale,
                  int32_t* err_flag, int ids_aligned16, int stage_cap_ids) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int64_t* s_ids = reinterpret_cast<int64_t*>(smem_raw);
  __shared__ __align__(8) uint64_t s_bar;

  const int fidx = sel.idx[blockIdx.y];
  const ptrec_feature_desc fd = feats[fidx];
  const int L = fd.bag_len;
  const int64_t b0 = (int64_t)blockIdx.x * kBagsPerCta;
  if (b0 >= B) return;
  const int nb = (int)min((int64_t)kBagsPerCta, B - b0);
  const int64_t first_slot = fd.id_base * B + b0 * L;
  const int64_t n_slots = (int64_t)nb * L;
  const int64_t* gsrc = ids + first_slot;
  const bool staged = n_slots <= stage_cap_ids;
  const int64_t* my_ids;  // index list of this warp's bag

  const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
  if (staged) {
    const bool bulk = ids_aligned16 && ((first_slot & 1) == 0) && ((n_slots & 1) == 0);
    if (bulk) {
      if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&s_bar, (uint32_t)n_slots * 8u);
        bulk_g2s(s_ids, gsrc, (uint32_t)n_slots * 8u, &s_bar);
      }
      __syncthreads();
      mbar_wait(&s_bar, 0);
    } else {
      for (int64_t j = threadIdx.x; j < n_slots; j += blockDim.x) s_ids[j] = gsrc[j];
      __syncthreads();
    }
    my_ids = s_ids + (int64_t)warp * L;
  } else {
    my_ids = gsrc + (int64_t)warp * L;
  }
  if (warp >= nb) return;

  const int64_t b = b0 + warp;
  const void* tab = table_ptrs[fd.table];
  const int64_t rows = table_rows[fd.table];
  constexpr int RPW = 32 / LPR;  // rows in flight per warp-load
  const int sg = lane32 / LPR, lane = lane32 % LPR;
  const bool lane_on = lane * VEC < D;
  int len = L;
  if (fd.mask_mode == PTREC_MASK_LENS) len = lens[(int64_t)fd.lens_col * B + b];

  RowVec<VEC> acc;
  acc.zero();
  int count = 0;
  for (int l0 = sg; l0 < L; l0 += RPW * kUnroll) {
    RowVec<VEC> r[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int l = l0 + u * RPW;
      r[u].zero();
      if (l < L) {
        const int64_t id = my_ids[l];
        bool valid;
        switch (fd.mask_mode) {
          case PTREC_MASK_PAD: valid = id != 0; break;
          case PTREC_MASK_PAD_KEEP_FIRST: valid = (id != 0) || (l == 0); break;
          case PTREC_MASK_LENS: valid = l < len; break;
          default: valid = true;
        }
        if (id < 0 && (fd.flags & PTREC_FEAT_NEG_IS_PAD)) valid = false;
        if (valid) {
          ++count;
          if ((uint64_t)id < (uint64_t)rows) {
            if (lane_on) r[u] = load_table_row<VEC, WB, true>(tab, id * row_stride + lane * VEC);
          } else if (err_flag != nullptr && lane == 0) {
            *err_flag = 1;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) acc.add(r[u]);
  }
  // deterministic xor tree across the RPW sub-warps
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc.v[k] += __shfl_xor_sync(0xffffffffu, acc.v[k], o);
    count += __shfl_xor_sync(0xffffffffu, count, o);
  }
  const float sc = pool_scale(fd.pooling, count);
  if (sg == 0) {
    acc.scale(sc);
    if (lane_on) store_row<VEC>(out + b * out_row_stride + fd.out_col + lane * VEC, acc);
    if (lane == 0 && bag_scale != nullptr) bag_scale[(int64_t)fidx * B + b] = sc;
  }
}

// ------------------------------------------------------------------------------------ dispatch
template <int VEC, int LPR, bool WB>
static int launch_gather(const void* const* table_ptrs, const int64_t* table_rows, int D, int64_t row_stride,
                         const ptrec_feature_desc* feats, const FeatSel& onehot, const FeatSel& bags,
                         int max_bag_len, const int64_t* ids, const int32_t* lens, int64_t B,
                         float* out, int64_t out_row_stride, float* bag_scale, int32_t* err_flag,
                         int shards, cudaStream_t st) {
  const int ids_al = aligned16(ids) ? 1 : 0;
  if (onehot.n > 0) {
    const int use_bulk = (ids_al && (B % 2 == 0)) ? 1 : 0;
    const size_t smem = (size_t)onehot.n * (kOneHotTileB * 8 + 8 * (shards > 1 ? shards : 1) + 8 + 8 + 8);
    const unsigned grid = (unsigned)ceil_div(B, kOneHotTileB);
    if (shards > 1) {
      if (smem > 48 * 1024)
        PTREC_CUDA(cudaFuncSetAttribute(gather_onehot_kernel<VEC, LPR, true, WB>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      gather_onehot_kernel<VEC, LPR, true, WB><<<grid, kOneHotThreads, smem, st>>>(
          table_ptrs, table_rows, D, row_stride, feats, onehot, ids, B, out, out_row_stride, err_flag, use_bulk,
          shards);
    } else {
      gather_onehot_kernel<VEC, LPR, false, WB><<<grid, kOneHotThreads, smem, st>>>(
          table_ptrs, table_rows, D, row_stride, feats, onehot, ids, B, out, out_row_stride, err_flag, use_bulk, 1);
    }
    PTREC_LAUNCH_CHECK("gather_onehot_kernel");
  }
  if (bags.n > 0) {
    const int stage_L = max_bag_len <= kBagStageMaxL ? max_bag_len : 0;
    const size_t smem = (size_t)stage_L * kBagsPerCta * 8;
    if (smem > 48 * 1024) {
      PTREC_CUDA(cudaFuncSetAttribute(gather_bag_kernel<VEC, LPR, WB>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid((unsigned)ceil_div(B, kBagsPerCta), (unsigned)bags.n);
    gather_bag_kernel<VEC, LPR, WB><<<grid, kBagThreads, smem, st>>>(
        table_ptrs, table_rows, D, row_stride, feats, bags, ids, lens, B, out, out_row_stride, bag_scale,
        err_flag, ids_al, stage_L * kBagsPerCta);
    PTREC_LAUNCH_CHECK("gather_bag_kernel");
  }
  return PTREC_OK;
}

// ------------------------------------------------------------------------------------ index_prep
struct ValidCountIn {
  const int64_t* ids;
  const int32_t* lens;
  int64_t L;
  int mask_mode;
  __device__ int operator()(int64_t b) const {
    if (mask_mode == PTREC_MASK_NONE) return (int)L;
    if (mask_mode == PTREC_MASK_LENS) return min((int)L, max(0, lens[b]));
    int c = 0;
    const int64_t* row = ids + b * L;
    for (int64_t l = 0; l < L; ++l) c += (row[l] != 0 || (mask_mode == PTREC_MASK_PAD_KEEP_FIRST && l == 0));
    return c;
  }
};
struct OffsetsOut {
  int64_t* offsets;
  int64_t n;
  __device__ void operator()(int64_t b, int prefix, int v) const {
    offsets[b] = prefix;
    if (b == n - 1) offsets[n] = (int64_t)prefix + v;
  }
};

// warp per bag: ordered compaction of the valid slots
__global__ void __launch_bounds__(256)
index_compact_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ lens, int64_t B,
                     int64_t L, int mask_mode, const int64_t* __restrict__ offsets,
                     int64_t* __restrict__ out_ids) {
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const int64_t* row = ids + b * L;
  int64_t w = offsets[b];
  const int len = (mask_mode == PTREC_MASK_LENS) ? lens[b] : 0;
  for (int64_t l0 = 0; l0 < L; l0 += 32) {
    const int64_t l = l0 + lane;
    int64_t id = 0;
    bool valid = false;
    if (l < L) {
      id = row[l];
      valid = (mask_mode == PTREC_MASK_LENS) ? (l < len) : slot_valid(mask_mode, id, (int)l, nullptr, 0, 0, 0);
    }
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (valid) out_ids[w + __popc(m & ((1u << lane) - 1u))] = id;
    w += __popc(m);
  }
}

}  // namespace ptrec

using namespace ptrec;

extern "C" size_t ptrec_index_prep_workspace_bytes(int64_t B) {
  return align_up((size_t)scan_num_tiles(B) * sizeof(int) + 16, 256);
}

extern "C" int ptrec_index_prep(const int64_t* ids_padded, const int32_t* lens, int64_t B, int64_t L,
                                int32_t mask_mode, int64_t* out_ids, int64_t* out_offsets,
                                void* workspace, size_t workspace_bytes, void* stream) {
  PTREC_CHECK_ARG(B >= 0 && L >= 1, PTREC_EINVAL, "index_prep: bad B=%lld L=%lld", (long long)B, (long long)L);
  PTREC_CHECK_ARG(B * L < (int64_t)0x7fffffff, PTREC_EUNSUPPORTED, "index_prep: B*L must be < 2^31");
  PTREC_CHECK_ARG(mask_mode >= 0 && mask_mode <= 3, PTREC_EINVAL, "index_prep: bad mask_mode %d", mask_mode);
  PTREC_CHECK_ARG(mask_mode != PTREC_MASK_LENS || lens != nullptr, PTREC_EINVAL, "index_prep: lens required");
  PTREC_CHECK_ARG(out_ids && out_offsets && (ids_padded || B == 0), PTREC_EINVAL, "index_prep: null pointer");
  PTREC_CHECK_ARG(workspace_bytes >= ptrec_index_prep_workspace_bytes(B) && workspace, PTREC_EWORKSPACE,
                  "index_prep: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    PTREC_CUDA(cudaMemsetAsync(out_offsets, 0, sizeof(int64_t), st));
    return PTREC_OK;
  }
  int* tile_sums = reinterpret_cast<int*>(workspace);
  const int tiles = scan_num_tiles(B);
  ValidCountIn in{ids_padded, lens, L, mask_mode};
  OffsetsOut outf{out_offsets, B};
  scan_tile_sums_kernel<<<tiles, kScanThreads, 0, st>>>(in, B, tile_sums);
  PTREC_LAUNCH_CHECK("scan_tile_sums_kernel");
  scan_top_kernel<<<1, 1024, 0, st>>>(tile_sums, tiles, nullptr);
  PTREC_LAUNCH_CHECK("scan_top_kernel");
  scan_apply_kernel<<<tiles, kScanThreads, 0, st>>>(in, outf, B, tile_sums);
  PTREC_LAUNCH_CHECK("scan_apply_kernel");
  index_compact_kernel<<<(unsigned)ceil_div(B, 8), 256, 0, st>>>(ids_padded, lens, B, L, mask_mode,
                                                                  out_offsets, out_ids);
  PTREC_LAUNCH_CHECK("index_compact_kernel");
  return PTREC_OK;
}

static int gather_pool_fwd_impl(const void* const* table_ptrs, const int64_t* table_rows, int32_t T, int32_t shards,
                                int32_t D, int64_t row_stride, int32_t dtype, const ptrec_feature_desc* feats,
                                const ptrec_feature_desc* feats_host, int32_t F, const int64_t* ids,
                                const int32_t* lens, int64_t B, float* out, int64_t out_row_stride,
                                float* bag_scale, int32_t* err_flag, void* stream) {
  PTREC_CHECK_ARG(dtype == PTREC_F32 || dtype == PTREC_BF16, PTREC_EUNSUPPORTED, "gather: unknown table dtype %d", dtype);
  PTREC_CHECK_ARG(T >= 1 && T <= kMaxTables && F >= 1 && F <= kMaxFeatures, PTREC_EINVAL,
                  "gather: T=%d F=%d out of range (max %d)", T, F, kMaxFeatures);
  PTREC_CHECK_ARG(table_ptrs && table_rows && feats && feats_host && ids && out, PTREC_EINVAL, "gather: null pointer");
  PTREC_CHECK_ARG(B >= 0, PTREC_EINVAL, "gather: B < 0");
  const bool d_ok = D == 1 || D == 2 || (D >= 4 && D <= 128 && D % 4 == 0);
  PTREC_CHECK_ARG(d_ok, PTREC_EUNSUPPORTED, "gather: D=%d unsupported (1, 2, or a multiple of 4 up to 128)", D);
  const int vec = D >= 4 ? 4 : D;
  PTREC_CHECK_ARG(row_stride >= D && row_stride % vec == 0, PTREC_EALIGN, "gather: table row_stride %lld invalid for D=%d",
                  (long long)row_stride, D);
  PTREC_CHECK_ARG(((uintptr_t)out % (vec * 4)) == 0 && (out_row_stride % vec) == 0, PTREC_EALIGN,
                  "gather: out / out_row_stride not aligned to %d bytes", vec * 4);
  if (B == 0) return PTREC_OK;

  FeatSel onehot, bags;
  onehot.n = bags.n = 0;
  int max_bag_len = 0;
  bool need_lens = false;
  int64_t total_L = 0;
  for (int f = 0; f < F; ++f) {
    const ptrec_feature_desc& fd = feats_host[f];
    PTREC_CHECK_ARG(fd.table >= 0 && fd.table < T && fd.bag_len >= 1, PTREC_EINVAL, "gather: bad feature %d", f);
    PTREC_CHECK_ARG(f == 0 || fd.table >= feats_host[f - 1].table, PTREC_EINVAL, "gather: features must be ordered by table");
    PTREC_CHECK_ARG(fd.id_base == total_L, PTREC_EINVAL, "gather: feature %d id_base must be the running sum of bag_len", f);
    PTREC_CHECK_ARG(fd.out_col % vec == 0, PTREC_EALIGN, "gather: feature %d out_col not a multiple of %d", f, vec);
    total_L += fd.bag_len;
    if (fd.bag_len == 1 && fd.mask_mode == PTREC_MASK_NONE) {
      onehot.idx[onehot.n++] = (int16_t)f;
    } else {
      bags.idx[bags.n++] = (int16_t)f;
      if (fd.bag_len > max_bag_len) max_bag_len = fd.bag_len;
      if (fd.mask_mode == PTREC_MASK_LENS) need_lens = true;
    }
  }
  PTREC_CHECK_ARG(!need_lens || lens, PTREC_EINVAL, "gather: lens required by a PTREC_MASK_LENS feature");
  PTREC_CHECK_ARG(shards == 1 || bags.n == 0, PTREC_EUNSUPPORTED,
                  "gather: row-wise sharded tables serve one-hot fields only (pooled bags stay unsharded)");
  cudaStream_t st = (cudaStream_t)stream;

#define PTREC_GATHER(V, P)                                                                                           \
  return dtype == PTREC_BF16                                                                                         \
             ? launch_gather<V, P, true>(table_ptrs, table_rows, D, row_stride, feats, onehot, bags, max_bag_len, ids, \
                                         lens, B, out, out_row_stride, bag_scale, err_flag, shards, st)              \
             : launch_gather<V, P, false>(table_ptrs, table_rows, D, row_stride, feats, onehot, bags, max_bag_len, ids, \
                                          lens, B, out, out_row_stride, bag_scale, err_flag, shards, st)
  if (D == 1) PTREC_GATHER(1, 1);
  if (D == 2) PTREC_GATHER(2, 1);
  const int lanes = D / 4;
  if (lanes <= 1) PTREC_GATHER(4, 1);
  if (lanes <= 2) PTREC_GATHER(4, 2);
  if (lanes <= 4) PTREC_GATHER(4, 4);
  if (lanes <= 8) PTREC_GATHER(4, 8);
  if (lanes <= 16) PTREC_GATHER(4, 16);
  PTREC_GATHER(4, 32);
#undef PTREC_GATHER
}

extern "C" int ptrec_embedding_gather_pool_fwd(const void* const* table_ptrs, const int64_t* table_rows,
                                               int32_t T, int32_t D, int64_t row_stride, int32_t dtype,
                                               const ptrec_feature_desc* feats,
                                               const ptrec_feature_desc* feats_host, int32_t F,
                                               const int64_t* ids, const int32_t* lens, int64_t B,
                                               float* out, int64_t out_row_stride, float* bag_scale,
                                               int32_t* err_flag, void* stream) {
  return gather_pool_fwd_impl(table_ptrs, table_rows, T, 1, D, row_stride, dtype, feats, feats_host, F, ids, lens, B,
                              out, out_row_stride, bag_scale, err_flag, stream);
}

extern "C" int ptrec_embedding_gather_pool_fwd_sharded(const void* const* shard_ptrs, const int64_t* table_rows,
                                                       int32_t T, int32_t G, int32_t D, int64_t row_stride,
                                                       int32_t dtype, const ptrec_feature_desc* feats,
                                                       const ptrec_feature_desc* feats_host, int32_t F,
                                                       const int64_t* ids, int64_t B, float* out,
                                                       int64_t out_row_stride, int32_t* err_flag, void* stream) {
  PTREC_CHECK_ARG(G >= 1 && G <= 64, PTREC_EINVAL, "gather_sharded: G=%d out of range (1..64)", G);
  return gather_pool_fwd_impl(shard_ptrs, table_rows, T, G, D, row_stride, dtype, feats, feats_host, F, ids, nullptr,
                              B, out, out_row_stride, nullptr, err_flag, stream);
}
