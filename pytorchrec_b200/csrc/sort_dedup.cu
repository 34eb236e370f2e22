// K2a: per-table stable LSD radix sort of the batch's lookup slots + dedup into segments.
//
// Integer-only, bit-exact against torch.sort(stable=True) / torch.unique(sorted=True) per table.
// Bound: launch latency at the configs' sizes (N = 4e5..2e6 keys move ~16 B each per pass);
// the design goal is therefore few, wide, spin-free launches:
//   per pass (8- or 10-bit digits): histogram (tile x digit counts) -> per-table scan -> stable scatter
//   (warp match-any ranking, no atomics on the output path); then head flags -> scan -> segment list.
// Tables are sorted independently (a tile never straddles two tables), so only ceil(log2(rows)/RB)
// passes are needed instead of covering a global (table, id) key: 2 passes of 10 bits for 1e5..1e6-row
// tables (3 with 8-bit digits), 3 for 5e7 rows.
#include "common.cuh"
#include "scan.cuh"

namespace ptrec {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048
constexpr int kMaxRadix = 1024;

struct TableLayout {
  int32_t T;
  int32_t total_tiles;
  int64_t Lstart[kMaxTables + 1];      // slots of table t: [Lstart[t]*B, Lstart[t+1]*B)
  int32_t tile_prefix[kMaxTables + 1]; // first sort tile of table t
};

__device__ __forceinline__ int table_of_tile(const TableLayout& lay, int tile) {
  int lo = 0, hi = lay.T - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (lay.tile_prefix[mid] <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__device__ __forceinline__ int table_of_slot(const TableLayout& lay, int64_t B, int64_t j) {
  int lo = 0, hi = lay.T - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (lay.Lstart[mid] * B <= j) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// key of slot p straight from the id matrices (first pass only).  A sort tile lies inside one table,
// whose features are a contiguous descriptor range [f0, f0+nf): no search for single-feature tables,
// no division for one-hot features, 32-bit arithmetic throughout (N < 2^31).
struct KeyGen {
  const ptrec_feature_desc* feats;  // shared-memory copy
  int f0, nf;
  const int64_t* ids;
  const int32_t* lens;
  int64_t rows;  // rows of this tile's table
  int64_t B;
  __device__ __forceinline__ uint32_t operator()(int64_t p) const {
    int f = f0;
    for (int e = f0 + nf - 1; f < e && feats[f + 1].id_base * B <= p; ++f) {}
    const ptrec_feature_desc& fd = feats[f];
    const int64_t id = ids[p];
    bool valid = (uint64_t)id < (uint64_t)rows;
    if (fd.mask_mode != PTREC_MASK_NONE) {
      const uint32_t rel = (uint32_t)(p - fd.id_base * B);
      const uint32_t b = rel / (uint32_t)fd.bag_len;
      const int l = (int)(rel - b * (uint32_t)fd.bag_len);
      valid = valid && slot_valid(fd.mask_mode, id, l, lens, fd.lens_col, B, b);
    }
    return valid ? (uint32_t)id : kMaskedKey;
  }
};

// feature range of table t (thread 0 scans the <= 128 descriptors; result broadcast through smem)
__device__ __forceinline__ void table_feature_range(const ptrec_feature_desc* s_feats, int F, int t, int* s_rng) {
  if (threadIdx.x == 0) {
    int f0 = 0, nf = 0;
    for (int f = 0; f < F; ++f) {
      if (s_feats[f].table == t) {
        if (nf == 0) f0 = f;
        ++nf;
      }
    }
    s_rng[0] = f0;
    s_rng[1] = nf;
  }
}

template <bool FIRST, int RB>
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(TableLayout lay, int64_t B, int shift, const uint32_t* __restrict__ keys_in,
                  const ptrec_feature_desc* __restrict__ feats, int F, const int64_t* __restrict__ ids,
                  const int32_t* __restrict__ lens, const int64_t* __restrict__ table_rows,
                  int* __restrict__ hist) {
  constexpr int kRadix = 1 << RB;
  __shared__ int s_hist[kRadix];
  __shared__ ptrec_feature_desc s_feats[FIRST ? kMaxFeatures : 1];
  const int tile = blockIdx.x;
  const int t = table_of_tile(lay, tile);
  const int k = tile - lay.tile_prefix[t];
  const int tiles_t = lay.tile_prefix[t + 1] - lay.tile_prefix[t];
  const int64_t beg = lay.Lstart[t] * B + (int64_t)k * kSortTile;
  const int64_t end = min(lay.Lstart[t + 1] * B, beg + kSortTile);
  __shared__ int s_rng[2];
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) s_hist[d] = 0;
  if (FIRST) {
    load_feats(s_feats, feats, F);
    __syncthreads();
    table_feature_range(s_feats, F, t, s_rng);
  }
  __syncthreads();
  KeyGen gen{s_feats, FIRST ? s_rng[0] : 0, FIRST ? s_rng[1] : 0, ids, lens, FIRST ? table_rows[t] : 0, B};
  for (int64_t j = beg + threadIdx.x; j < end; j += kSortThreads) {
    const uint32_t key = FIRST ? gen(j) : keys_in[j];
    atomicAdd(&s_hist[(key >> shift) & (kRadix - 1)], 1);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads)
    hist[(int64_t)lay.tile_prefix[t] * kRadix + (int64_t)d * tiles_t + k] = s_hist[d];
}

// one CTA per table: exclusive scan of its [256][tiles_t] counts, seeded with the table's first slot
__global__ void __launch_bounds__(1024) radix_scan_kernel(TableLayout lay, int64_t B, int kRadix, int* __restrict__ hist) {
  __shared__ int s_warp[33];
  __shared__ int s_carry;
  const int t = blockIdx.x;
  const int tiles_t = lay.tile_prefix[t + 1] - lay.tile_prefix[t];
  const int M = tiles_t * kRadix;
  int* h = hist + (int64_t)lay.tile_prefix[t] * kRadix;
  if (threadIdx.x == 0) s_carry = (int)(lay.Lstart[t] * B);
  __syncthreads();
  for (int base = 0; base < M; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < M ? h[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, s_warp, &total);
    const int carry = s_carry;
    if (i < M) h[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
}

template <bool FIRST, int RB>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(TableLayout lay, int64_t B, int shift, const uint32_t* __restrict__ keys_in,
                     const int32_t* __restrict__ perm_in, const ptrec_feature_desc* __restrict__ feats,
                     int F, const int64_t* __restrict__ ids, const int32_t* __restrict__ lens,
                     const int64_t* __restrict__ table_rows, const int* __restrict__ offsets,
                     uint32_t* __restrict__ keys_out, int32_t* __restrict__ perm_out) {
  constexpr int NW = kSortThreads / 32;
  constexpr int kRadix = 1 << RB;
  __shared__ int s_cnt[NW][kRadix];
  __shared__ int s_goff[kRadix];
  __shared__ ptrec_feature_desc s_feats[FIRST ? kMaxFeatures : 1];
  const int tile = blockIdx.x;
  const int t = table_of_tile(lay, tile);
  const int k = tile - lay.tile_prefix[t];
  const int tiles_t = lay.tile_prefix[t + 1] - lay.tile_prefix[t];
  const int64_t beg = lay.Lstart[t] * B + (int64_t)k * kSortTile;
  const int64_t end = min(lay.Lstart[t + 1] * B, beg + kSortTile);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) {
#pragma unroll
    for (int w = 0; w < NW; ++w) s_cnt[w][d] = 0;
    s_goff[d] = offsets[(int64_t)lay.tile_prefix[t] * kRadix + (int64_t)d * tiles_t + k];
  }
  __shared__ int s_rng[2];
  if (FIRST) {
    load_feats(s_feats, feats, F);
    __syncthreads();
    table_feature_range(s_feats, F, t, s_rng);
  }
  __syncthreads();
  KeyGen gen{s_feats, FIRST ? s_rng[0] : 0, FIRST ? s_rng[1] : 0, ids, lens, FIRST ? table_rows[t] : 0, B};

  uint32_t key[kSortItems];
  int32_t pay[kSortItems];
  int rank[kSortItems];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const int64_t j = beg + warp * (32 * kSortItems) + i * 32 + lane;
    const bool in = j < end;
    key[i] = 0;
    pay[i] = 0;
    if (in) {
      key[i] = FIRST ? gen(j) : keys_in[j];
      pay[i] = FIRST ? (int32_t)j : perm_in[j];
    }
    const int d = in ? (int)((key[i] >> shift) & (kRadix - 1)) : kRadix;  // kRadix = "no key" group
    const unsigned m = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (in) base = s_cnt[warp][d];
    __syncwarp();
    if (in && lane == leader) s_cnt[warp][d] = base + __popc(m);
    __syncwarp();
    rank[i] = base + __popc(m & lt);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) {  // exclusive prefix over warps for digit d
    int run = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int c = s_cnt[w][d];
      s_cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const int64_t j = beg + warp * (32 * kSortItems) + i * 32 + lane;
    if (j < end) {
      const int d = (int)((key[i] >> shift) & (kRadix - 1));
      const int pos = s_goff[d] + s_cnt[warp][d] + rank[i];
      keys_out[pos] = key[i];
      perm_out[pos] = pay[i];
    }
  }
}

// ---- dedup: head flags -> exclusive scan -> segment list --------------------------------------
struct HeadIn {
  TableLayout lay;
  int64_t B;
  const uint32_t* keys;
  __device__ int operator()(int64_t j) const {
    if (j == 0) return 1;
    if (keys[j] != keys[j - 1]) return 1;
    const int t = table_of_slot(lay, B, j);
    return (lay.Lstart[t] * B == j) ? 1 : 0;
  }
};
struct SegOut {
  TableLayout lay;
  int64_t B;
  int64_t N;
  int32_t* seg_start;
  ptrec_segment_meta* seg_meta;
  const uint32_t* keys;
  const int32_t* perm;
  __device__ void operator()(int64_t j, int prefix, int v) const {
    if (v) {
      seg_start[prefix] = (int32_t)j;
      int4 m;
      m.x = (int)keys[j];
      m.y = perm[j];
      m.z = table_of_slot(lay, B, j);
      m.w = 0;
      *reinterpret_cast<int4*>(seg_meta + prefix) = m;  // one 16-byte record per segment
    }
    if (j == N - 1) seg_start[prefix + v] = (int32_t)N;
  }
};


// ---- one-sweep path: P + 2 launches instead of 3 P + 3 ------------------------------------------------------------
// os_hist_kernel   keys generated ONCE from the id matrices (validity, masks) into a uint32 array; the digit
//                  histograms of EVERY pass are counted in the same sweep (shared-memory atomics, then one global
//                  integer atomicAdd per non-empty (table, pass, digit) — integer adds: order-independent); the
//                  look-back words of the later launches are reset by the same CTAs.
// os_pass_kernel   one launch per digit: a tile ranks its keys (warp match-any, stable), publishes its digit counts
//                  and obtains the counts of the tiles before it IN ITS TABLE by decoupled look-back (aggregate /
//                  inclusive words, 30-bit count + 2 flag bits, relaxed gpu-scope loads); digit bases come from the
//                  table's global histogram, scanned by every CTA for itself.  A tile only ever waits on tiles with a
//                  smaller block index, and every wait is bounded (trap instead of a hang).
// os_dedup_kernel  head flags + look-back scan of the head counts + the 16-byte segment records, in one launch.
// Same stable LSD order, so the outputs are bit-identical to the other two paths.
constexpr uint32_t kOsAggr = 1u << 30, kOsIncl = 2u << 30, kOsMask = (1u << 30) - 1u;
constexpr int kOsMaxPasses = 4;

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// exclusive prefix over the tiles [first, tile) of the published words at w[j * pitch]; bounded spin
__device__ __forceinline__ uint32_t os_look_back(const uint32_t* w, int64_t pitch, int first, int tile) {
  uint32_t ex = 0;
  for (int j = tile - 1; j >= first; --j) {
    uint32_t v = ld_relaxed_u32(w + (int64_t)j * pitch);
    for (uint32_t spins = 0; (v >> 30) == 0; ++spins) {
      if (spins > (1u << 22)) __trap();
      __nanosleep(20);
      v = ld_relaxed_u32(w + (int64_t)j * pitch);
    }
    ex += v & kOsMask;
    if (v & kOsIncl) break;
  }
  return ex;
}

template <int RB>
__global__ void __launch_bounds__(kSortThreads)
os_hist_kernel(TableLayout lay, int64_t B, int P, const ptrec_feature_desc* __restrict__ feats, int F,
               const int64_t* __restrict__ ids, const int32_t* __restrict__ lens,
               const int64_t* __restrict__ table_rows, uint32_t* __restrict__ keys0, int* __restrict__ ghist,
               uint32_t* __restrict__ status, uint32_t* __restrict__ status_seg) {
  constexpr int kRadix = 1 << RB;
  __shared__ int s_hist[kOsMaxPasses * kRadix];
  __shared__ ptrec_feature_desc s_feats[kMaxFeatures];
  __shared__ int s_rng[2];
  const int tile = blockIdx.x;
  const int t = table_of_tile(lay, tile);
  const int k = tile - lay.tile_prefix[t];
  const int64_t beg = lay.Lstart[t] * B + (int64_t)k * kSortTile;
  const int64_t end = min(lay.Lstart[t + 1] * B, beg + kSortTile);
  for (int d = threadIdx.x; d < P * kRadix; d += kSortThreads) s_hist[d] = 0;
  // reset the look-back words this tile owns in every later launch of this call
  for (int d = threadIdx.x; d < P * kRadix; d += kSortThreads)
    status[((int64_t)(d / kRadix) * lay.total_tiles + tile) * kRadix + (d % kRadix)] = 0u;
  if (threadIdx.x == 0) status_seg[tile] = 0u;
  load_feats(s_feats, feats, F);
  __syncthreads();
  table_feature_range(s_feats, F, t, s_rng);
  __syncthreads();
  KeyGen gen{s_feats, s_rng[0], s_rng[1], ids, lens, table_rows[t], B};
  for (int64_t j = beg + threadIdx.x; j < end; j += kSortThreads) {
    const uint32_t key = gen(j);
    keys0[j] = key;
    for (int p = 0; p < P; ++p) atomicAdd(&s_hist[p * kRadix + ((key >> (RB * p)) & (kRadix - 1))], 1);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < P * kRadix; d += kSortThreads) {
    const int c = s_hist[d];
    if (c) atomicAdd(&ghist[(int64_t)t * (kOsMaxPasses * kRadix) + d], c);
  }
}

template <int RB>
__global__ void __launch_bounds__(kSortThreads)
os_pass_kernel(TableLayout lay, int64_t B, int p, const uint32_t* __restrict__ keys_in,
               const int32_t* __restrict__ perm_in, const int* __restrict__ ghist, uint32_t* __restrict__ status,
               uint32_t* __restrict__ keys_out, int32_t* __restrict__ perm_out) {
  constexpr int NW = kSortThreads / 32;
  constexpr int kRadix = 1 << RB;
  constexpr int DPT = kRadix / kSortThreads;  // consecutive digits per thread (1 or 4)
  __shared__ int s_cnt[NW][kRadix];
  __shared__ int s_goff[kRadix];
  __shared__ int s_warp[33];
  const int tile = blockIdx.x;
  const int t = table_of_tile(lay, tile);
  const int first = lay.tile_prefix[t];
  const int64_t beg = lay.Lstart[t] * B + (int64_t)(tile - first) * kSortTile;
  const int64_t end = min(lay.Lstart[t + 1] * B, beg + kSortTile);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int shift = RB * p;
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) {
#pragma unroll
    for (int w = 0; w < NW; ++w) s_cnt[w][d] = 0;
  }
  __syncthreads();
  uint32_t key[kSortItems];
  int32_t pay[kSortItems];
  int rank[kSortItems];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const int64_t j = beg + warp * (32 * kSortItems) + i * 32 + lane;
    const bool in = j < end;
    key[i] = 0;
    pay[i] = 0;
    if (in) {
      key[i] = keys_in[j];
      pay[i] = perm_in != nullptr ? perm_in[j] : (int32_t)j;
    }
    const int d = in ? (int)((key[i] >> shift) & (kRadix - 1)) : kRadix;  // kRadix = "no key" group
    const unsigned m = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (in) base = s_cnt[warp][d];
    __syncwarp();
    if (in && lane == leader) s_cnt[warp][d] = base + __popc(m);
    __syncwarp();
    rank[i] = base + __popc(m & lt);
  }
  __syncthreads();
  // this thread's digits: exclusive prefix over the warps, tile count, publish, look back, digit base
  uint32_t* st = status + ((int64_t)p * lay.total_tiles) * kRadix;
  int cnt[DPT], tot[DPT];
  int mine = 0;
#pragma unroll
  for (int i = 0; i < DPT; ++i) {
    const int d = threadIdx.x * DPT + i;
    int run = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int c = s_cnt[w][d];
      s_cnt[w][d] = run;
      run += c;
    }
    cnt[i] = run;
    st_relaxed_u32(st + (int64_t)tile * kRadix + d, (tile == first ? kOsIncl : kOsAggr) | (uint32_t)run);
    tot[i] = ghist[(int64_t)t * (kOsMaxPasses * kRadix) + p * kRadix + d];
    mine += tot[i];
  }
  int total;
  int dbase = block_exclusive_scan(mine, s_warp, &total) + (int)(lay.Lstart[t] * B);
#pragma unroll
  for (int i = 0; i < DPT; ++i) {
    const int d = threadIdx.x * DPT + i;
    uint32_t ex = 0;
    if (tile != first) {
      ex = os_look_back(st + d, kRadix, first, tile);
      st_relaxed_u32(st + (int64_t)tile * kRadix + d, kOsIncl | (ex + (uint32_t)cnt[i]));
    }
    s_goff[d] = dbase + (int)ex;
    dbase += tot[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const int64_t j = beg + warp * (32 * kSortItems) + i * 32 + lane;
    if (j < end) {
      const int d = (int)((key[i] >> shift) & (kRadix - 1));
      const int pos = s_goff[d] + s_cnt[warp][d] + rank[i];
      keys_out[pos] = key[i];
      perm_out[pos] = pay[i];
    }
  }
}

__global__ void __launch_bounds__(kScanThreads)
os_dedup_kernel(HeadIn in, SegOut out, int64_t n, uint32_t* __restrict__ status_seg, int32_t* __restrict__ n_seg) {
  __shared__ int s_warp[33];
  __shared__ int s_prefix;
  const int tile = blockIdx.x;
  const int64_t base = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k;
    v[k] = (i < n) ? in(i) : 0;
    s += v[k];
  }
  int total;
  int prefix = block_exclusive_scan(s, s_warp, &total);
  if (threadIdx.x == 0) {
    uint32_t ex = 0;
    if (tile == 0) {
      st_relaxed_u32(status_seg, kOsIncl | (uint32_t)total);
    } else {
      st_relaxed_u32(status_seg + tile, kOsAggr | (uint32_t)total);
      ex = os_look_back(status_seg, 1, 0, tile);
      st_relaxed_u32(status_seg + tile, kOsIncl | (ex + (uint32_t)total));
    }
    s_prefix = (int)ex;
    if (tile == (int)gridDim.x - 1) *n_seg = (int)ex + total;
  }
  __syncthreads();
  prefix += s_prefix;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k;
    if (i < n) out(i, prefix, v[k]);
    prefix += v[k];
  }
}


// ---- fast path: one CTA sorts one table's batch entirely in shared memory ---------------------------------------
// When every table's slots fit (<= kSmemSortMax, e.g. the Criteo-shaped configs: 16384 lookups per table, ~21 k on
// the owner side of the sharded path) the whole sort + dedup is 2 launches instead of 3 per radix pass + 3:
//   smem_sort_kernel    keys generated once into shared memory; 16-bit local indices are LSD-radix sorted (8-bit
//                       digits, per-warp contiguous chunks + match-any ranking = stable); sorted keys / perm written
//                       coalesced; segments of this table counted
//   smem_segments_kernel  prefix over the (<= 128) table counts, head flags -> block scan -> segment records
// Same outputs, bit for bit, as the general path.
constexpr int kSmemSortThreads = 1024;
constexpr int kSmemSortMax = 22528;      // slots per table that FIT: 8 B each (key + 2 x uint16 index) = 176 KB
// ...but one CTA per table keeps only T SMs busy: measured at cfg2 (26 tables x 16384 slots) 118 us vs 80 us for the
// 9-launch global path, so the fast path is taken only where launch latency dominates (small batches)
constexpr int kSmemSortAuto = 4096;
constexpr int kSmemSortMaxFeat = 64;     // features per table copied to shared memory

__global__ void __launch_bounds__(kSmemSortThreads)
smem_sort_kernel(TableLayout lay, int64_t B, int passes, const ptrec_feature_desc* __restrict__ feats, int F,
                 const int64_t* __restrict__ ids, const int32_t* __restrict__ lens,
                 const int64_t* __restrict__ table_rows, int n_cap, uint32_t* __restrict__ keys_out,
                 int32_t* __restrict__ perm_out, int* __restrict__ seg_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_key = reinterpret_cast<uint32_t*>(smem_raw);              // [n_cap]
  uint16_t* s_idx0 = reinterpret_cast<uint16_t*>(s_key + n_cap);        // [n_cap]
  uint16_t* s_idx1 = s_idx0 + n_cap;                                    // [n_cap]
  uint16_t* s_cnt = s_idx1 + n_cap;                                     // [32][256]
  __shared__ ptrec_feature_desc s_feats[kSmemSortMaxFeat];
  __shared__ int s_dbase[256];
  __shared__ int s_wt[8];
  __shared__ int s_red[32];
  const int t = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t g0 = lay.Lstart[t] * B;
  const int n = (int)(lay.Lstart[t + 1] * B - g0);
  // features of this table: a contiguous descriptor range (features are ordered by table)
  int f0 = 0, nf = 0;
  for (int f = 0; f < F; ++f) {
    const int ft = feats[f].table;
    if (ft == t) {
      if (nf == 0) f0 = f;
      ++nf;
    }
  }
  for (int i = threadIdx.x; i < nf; i += kSmemSortThreads) s_feats[i] = feats[f0 + i];
  __syncthreads();
  KeyGen gen{s_feats, 0, nf, ids, lens, table_rows[t], B};
  for (int i = threadIdx.x; i < n; i += kSmemSortThreads) {
    s_key[i] = gen(g0 + i);
    s_idx0[i] = (uint16_t)i;
  }
  __syncthreads();
  const int chunk = ((n + 31) / 32 + 31) / 32 * 32;  // contiguous slots per warp, multiple of 32
  const int beg = min(n, warp * chunk), end = min(n, beg + chunk);
  const unsigned lt = (1u << lane) - 1u;
  uint16_t* in = s_idx0;
  uint16_t* out = s_idx1;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    for (int d = lane; d < 256; d += 32) s_cnt[warp * 256 + d] = 0;
    __syncwarp();
    for (int i0 = beg; i0 < end; i0 += 32) {  // phase 1: this warp's digit counts
      const int i = i0 + lane;
      const bool ok = i < end;
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int d = (int)((s_key[in[i]] >> shift) & 255u);
        const unsigned m = __match_any_sync(act, d);
        if (lane == __ffs(m) - 1) s_cnt[warp * 256 + d] += (uint16_t)__popc(m);
      }
      __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 256) {  // phase 2: exclusive prefix over warps per digit, then over digits
      const int d = threadIdx.x;
      int run = 0;
      for (int w = 0; w < 32; ++w) {
        const int c = s_cnt[w * 256 + d];
        s_cnt[w * 256 + d] = (uint16_t)run;
        run += c;
      }
      int v = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      if (lane == 31) s_wt[warp] = v;
      s_dbase[d] = v - run;  // exclusive inside this warp of digits
    }
    __syncthreads();
    if (threadIdx.x < 256) {
      int pre = 0;
      for (int w = 0; w < warp; ++w) pre += s_wt[w];
      s_dbase[threadIdx.x] += pre;
    }
    __syncthreads();
    for (int i0 = beg; i0 < end; i0 += 32) {  // phase 3: stable scatter
      const int i = i0 + lane;
      const bool ok = i < end;
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const uint16_t idx = in[i];
        const int d = (int)((s_key[idx] >> shift) & 255u);
        const unsigned m = __match_any_sync(act, d);
        const int base = s_cnt[warp * 256 + d];
        __syncwarp(act);
        if (lane == __ffs(m) - 1) s_cnt[warp * 256 + d] = (uint16_t)(base + __popc(m));
        out[s_dbase[d] + base + __popc(m & lt)] = idx;
      }
      __syncwarp();
    }
    __syncthreads();
    uint16_t* tmp = in;
    in = out;
    out = tmp;
  }
  // sorted order is in `in`: write keys / perm, count the heads
  int heads = 0;
  for (int j = threadIdx.x; j < n; j += kSmemSortThreads) {
    const uint16_t idx = in[j];
    const uint32_t k = s_key[idx];
    keys_out[g0 + j] = k;
    perm_out[g0 + j] = (int32_t)(g0 + idx);
    heads += (j == 0 || s_key[in[j - 1]] != k) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) heads += __shfl_xor_sync(0xffffffffu, heads, o);
  if (lane == 0) s_red[warp] = heads;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 32; ++w) tot += s_red[w];
    seg_count[t] = tot;
  }
}

__global__ void __launch_bounds__(kSmemSortThreads)
smem_segments_kernel(TableLayout lay, int64_t B, int64_t N, const uint32_t* __restrict__ keys,
                     const int32_t* __restrict__ perm, const int* __restrict__ seg_count,
                     int32_t* __restrict__ seg_start, ptrec_segment_meta* __restrict__ seg_meta,
                     int32_t* __restrict__ n_seg) {
  __shared__ int s_warp[33];
  __shared__ int s_base;
  const int t = blockIdx.x;
  const int64_t g0 = lay.Lstart[t] * B;
  const int n = (int)(lay.Lstart[t + 1] * B - g0);
  if (threadIdx.x == 0) {
    int pre = 0;
    for (int u = 0; u < t; ++u) pre += seg_count[u];
    s_base = pre;
    if (t == lay.T - 1) {
      const int total = pre + seg_count[t];
      *n_seg = total;
      seg_start[total] = (int32_t)N;
    }
  }
  __syncthreads();
  int carry = s_base;
  for (int b0 = 0; b0 < n; b0 += kSmemSortThreads) {
    const int j = b0 + threadIdx.x;
    int v = 0;
    uint32_t k = 0;
    if (j < n) {
      k = keys[g0 + j];
      v = (j == 0 || keys[g0 + j - 1] != k) ? 1 : 0;
    }
    int total;
    const int ex = block_exclusive_scan(v, s_warp, &total);
    if (v) {
      const int pos = carry + ex;
      seg_start[pos] = (int32_t)(g0 + j);
      int4 m;
      m.x = (int)k;
      m.y = perm[g0 + j];
      m.z = t;
      m.w = 0;
      *reinterpret_cast<int4*>(seg_meta + pos) = m;
    }
    carry += total;
    __syncthreads();
  }
}

}  // namespace ptrec

using namespace ptrec;

static int g_one_sweep = 1;  // large batches: 1 = one-sweep (P + 2 launches; default), 0 = 3 launches per radix pass + 3
extern "C" void ptrec_set_one_sweep_sort(int32_t on) { g_one_sweep = on ? 1 : 0; }
extern "C" int32_t ptrec_one_sweep_sort_enabled(void) { return g_one_sweep; }
static int g_smem_sort = 1;  // 0 = never, 1 = automatic (small batches), 2 = whenever it fits
extern "C" void ptrec_set_smem_sort(int32_t mode) { g_smem_sort = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }
extern "C" int32_t ptrec_smem_sort_enabled(void) { return g_smem_sort; }

static int build_layout(const ptrec_feature_desc* feats_host, int F, int T, int64_t B, TableLayout* lay,
                        int64_t* N_out) {
  PTREC_CHECK_ARG(T >= 1 && T <= kMaxTables && F >= 1 && F <= kMaxFeatures, PTREC_EINVAL,
                  "sort_dedup: T=%d F=%d out of range", T, F);
  lay->T = T;
  int64_t L = 0;
  int f = 0;
  for (int t = 0; t < T; ++t) {
    lay->Lstart[t] = L;
    while (f < F && feats_host[f].table == t) {
      PTREC_CHECK_ARG(feats_host[f].id_base == L && feats_host[f].bag_len >= 1, PTREC_EINVAL,
                      "sort_dedup: feature %d id_base/bag_len inconsistent", f);
      L += feats_host[f].bag_len;
      ++f;
    }
  }
  PTREC_CHECK_ARG(f == F, PTREC_EINVAL, "sort_dedup: features must be ordered by table (0..T-1)");
  lay->Lstart[T] = L;
  for (int t = T + 1; t <= kMaxTables; ++t) lay->Lstart[t] = L;
  int tiles = 0;
  for (int t = 0; t < T; ++t) {
    lay->tile_prefix[t] = tiles;
    tiles += (int)ceil_div((lay->Lstart[t + 1] - lay->Lstart[t]) * B, kSortTile);
  }
  for (int t = T; t <= kMaxTables; ++t) lay->tile_prefix[t] = tiles;
  lay->total_tiles = tiles;
  *N_out = L * B;
  return PTREC_OK;
}

extern "C" size_t ptrec_sort_dedup_workspace_bytes(int64_t N, int32_t T) {
  const size_t tiles = (size_t)ceil_div(N, kSortTile) + (size_t)T + 1;
  size_t bytes = 0;
  bytes += align_up((size_t)N * 4, 256);                       // keys_tmp
  bytes += align_up((size_t)N * 4, 256);                       // perm_tmp
  bytes += align_up(tiles * kMaxRadix * 4, 256);               // hist
  bytes += align_up(((size_t)scan_num_tiles(N) + 1) * 4, 256); // head scan tile sums
  bytes += align_up((size_t)(kMaxTables + 1) * 4, 256);        // per-table segment counts (shared-memory path)
  bytes += align_up((size_t)T * kOsMaxPasses * kMaxRadix * 4, 256);      // one-sweep: per-table digit histograms
  bytes += align_up(tiles * kOsMaxPasses * kMaxRadix * 4, 256);          // one-sweep: look-back words per pass
  bytes += align_up(tiles * 4, 256);                                     // one-sweep: look-back words of the dedup scan
  return bytes + 256;
}

extern "C" int ptrec_sort_dedup(const ptrec_feature_desc* feats, const ptrec_feature_desc* feats_host,
                                int32_t F, int32_t T, const int64_t* table_rows, int64_t max_rows_host,
                                const int64_t* ids, const int32_t* lens, int64_t B,
                                uint32_t* sorted_keys, int32_t* perm, int32_t* seg_start,
                                ptrec_segment_meta* seg_meta, int32_t* n_seg, void* workspace,
                                size_t workspace_bytes, void* stream) {
  PTREC_CHECK_ARG(feats && feats_host && table_rows && ids && sorted_keys && perm && seg_start && seg_meta && n_seg,
                  PTREC_EINVAL, "sort_dedup: null pointer");
  PTREC_CHECK_ARG(max_rows_host >= 1 && max_rows_host < (int64_t)0xFFFFFFFFLL, PTREC_EUNSUPPORTED,
                  "sort_dedup: tables must have < 2^32-1 rows");
  TableLayout lay;
  int64_t N = 0;
  int rc = build_layout(feats_host, F, T, B, &lay, &N);
  if (rc != PTREC_OK) return rc;
  PTREC_CHECK_ARG(N < (int64_t)0x7fffffff, PTREC_EUNSUPPORTED, "sort_dedup: N=%lld slots must be < 2^31", (long long)N);
  PTREC_CHECK_ARG(workspace && workspace_bytes >= ptrec_sort_dedup_workspace_bytes(N, T), PTREC_EWORKSPACE,
                  "sort_dedup: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    PTREC_CUDA(cudaMemsetAsync(n_seg, 0, sizeof(int32_t), st));
    PTREC_CUDA(cudaMemsetAsync(seg_start, 0, sizeof(int32_t), st));
    return PTREC_OK;
  }
  // key bits: smallest b with 2^b - 1 >= max_rows (so the masked key 0xFFFFFFFF sorts strictly last);
  // 10-bit digits when they save a pass over 8-bit digits
  int bits = 1;
  while (bits < 32 && ((1ull << bits) - 1ull) < (unsigned long long)max_rows_host) ++bits;
  const int p8 = (bits + 7) / 8, p10 = (bits + 9) / 10;
  const int RB = p10 < p8 ? 10 : 8;
  const int P = RB == 10 ? p10 : p8;

  unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
  uint32_t* keys_tmp = reinterpret_cast<uint32_t*>(w); w += align_up((size_t)N * 4, 256);
  int32_t* perm_tmp = reinterpret_cast<int32_t*>(w);   w += align_up((size_t)N * 4, 256);
  int* hist = reinterpret_cast<int*>(w);
  w += align_up(((size_t)ceil_div(N, kSortTile) + (size_t)T + 1) * kMaxRadix * 4, 256);
  int* tile_sums = reinterpret_cast<int*>(w);
  w += align_up(((size_t)scan_num_tiles(N) + 1) * 4, 256);
  int* seg_count = reinterpret_cast<int*>(w);
  w += align_up((size_t)(kMaxTables + 1) * 4, 256);
  int* ghist = reinterpret_cast<int*>(w);
  const size_t ghist_bytes = (size_t)T * kOsMaxPasses * kMaxRadix * 4;
  w += align_up(ghist_bytes, 256);
  uint32_t* os_status = reinterpret_cast<uint32_t*>(w);
  w += align_up(((size_t)ceil_div(N, kSortTile) + (size_t)T + 1) * kOsMaxPasses * kMaxRadix * 4, 256);
  uint32_t* os_status_seg = reinterpret_cast<uint32_t*>(w);

  // shared-memory fast path: every table's slots fit one CTA
  int64_t n_max = 0;
  bool fits = ptrec_smem_sort_enabled() != 0;
  {
    int f = 0;
    for (int t = 0; t < T; ++t) {
      n_max = std::max<int64_t>(n_max, (lay.Lstart[t + 1] - lay.Lstart[t]) * B);
      int nf = 0;
      while (f < F && feats_host[f].table == t) { ++nf; ++f; }
      if (nf > kSmemSortMaxFeat) fits = false;
    }
  }
  if (fits && n_max <= (ptrec_smem_sort_enabled() == 2 ? kSmemSortMax : kSmemSortAuto)) {
    const int n_cap = (int)((n_max + 7) / 8 * 8);
    const size_t smem = (size_t)n_cap * 8 + 32 * 256 * 2;
    static bool attr_set = false;
    if (!attr_set) {
      PTREC_CUDA(cudaFuncSetAttribute(smem_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((size_t)kSmemSortMax * 8 + 32 * 256 * 2)));
      attr_set = true;
    }
    smem_sort_kernel<<<T, kSmemSortThreads, smem, st>>>(lay, B, (bits + 7) / 8, feats, F, ids, lens, table_rows, n_cap,
                                                        sorted_keys, perm, seg_count);
    PTREC_LAUNCH_CHECK("smem_sort_kernel");
    smem_segments_kernel<<<T, kSmemSortThreads, 0, st>>>(lay, B, N, sorted_keys, perm, seg_count, seg_start, seg_meta,
                                                         n_seg);
    PTREC_LAUNCH_CHECK("smem_segments_kernel");
    return PTREC_OK;
  }

  if (g_one_sweep && P <= kOsMaxPasses && scan_num_tiles(N) <= lay.total_tiles) {
    // keys0 lives in whichever key buffer pass 0 does NOT write
    const bool first_to_out = ((P - 1) % 2) == 0;
    uint32_t* keys0 = first_to_out ? keys_tmp : sorted_keys;
    PTREC_CUDA(cudaMemsetAsync(ghist, 0, ghist_bytes, st));
    if (RB == 10)
      os_hist_kernel<10><<<lay.total_tiles, kSortThreads, 0, st>>>(lay, B, P, feats, F, ids, lens, table_rows, keys0, ghist,
                                                                  os_status, os_status_seg);
    else
      os_hist_kernel<8><<<lay.total_tiles, kSortThreads, 0, st>>>(lay, B, P, feats, F, ids, lens, table_rows, keys0, ghist,
                                                                 os_status, os_status_seg);
    PTREC_LAUNCH_CHECK("os_hist_kernel");
    const uint32_t* kin = keys0;
    const int32_t* pin = nullptr;
    for (int p = 0; p < P; ++p) {
      const bool to_out = ((P - 1 - p) % 2) == 0;
      uint32_t* kout = to_out ? sorted_keys : keys_tmp;
      int32_t* pout = to_out ? perm : perm_tmp;
      if (RB == 10)
        os_pass_kernel<10><<<lay.total_tiles, kSortThreads, 0, st>>>(lay, B, p, kin, pin, ghist, os_status, kout, pout);
      else
        os_pass_kernel<8><<<lay.total_tiles, kSortThreads, 0, st>>>(lay, B, p, kin, pin, ghist, os_status, kout, pout);
      PTREC_LAUNCH_CHECK("os_pass_kernel");
      kin = kout;
      pin = pout;
    }
    HeadIn hin{lay, B, sorted_keys};
    SegOut sout{lay, B, N, seg_start, seg_meta, sorted_keys, perm};
    os_dedup_kernel<<<scan_num_tiles(N), kScanThreads, 0, st>>>(hin, sout, N, os_status_seg, n_seg);
    PTREC_LAUNCH_CHECK("os_dedup_kernel");
    return PTREC_OK;
  }

  const uint32_t* kin = nullptr;
  const int32_t* pin = nullptr;
  for (int p = 0; p < P; ++p) {
    const bool to_out = ((P - 1 - p) % 2) == 0;
    uint32_t* kout = to_out ? sorted_keys : keys_tmp;
    int32_t* pout = to_out ? perm : perm_tmp;
    const int shift = RB * p;
#define PTREC_HIST(FIRST_, RB_) \
  radix_hist_kernel<FIRST_, RB_><<<lay.total_tiles, kSortThreads, 0, st>>>(lay, B, shift, kin, feats, F, ids, lens, table_rows, hist)
#define PTREC_SCAT(FIRST_, RB_) \
  radix_scatter_kernel<FIRST_, RB_><<<lay.total_tiles, kSortThreads, 0, st>>>(lay, B, shift, kin, pin, feats, F, ids, lens, table_rows, hist, kout, pout)
    if (p == 0) { if (RB == 10) PTREC_HIST(true, 10); else PTREC_HIST(true, 8); }
    else        { if (RB == 10) PTREC_HIST(false, 10); else PTREC_HIST(false, 8); }
    PTREC_LAUNCH_CHECK("radix_hist_kernel");
    radix_scan_kernel<<<T, 1024, 0, st>>>(lay, B, 1 << RB, hist);
    PTREC_LAUNCH_CHECK("radix_scan_kernel");
    if (p == 0) { if (RB == 10) PTREC_SCAT(true, 10); else PTREC_SCAT(true, 8); }
    else        { if (RB == 10) PTREC_SCAT(false, 10); else PTREC_SCAT(false, 8); }
    PTREC_LAUNCH_CHECK("radix_scatter_kernel");
#undef PTREC_HIST
#undef PTREC_SCAT
    kin = kout;
    pin = pout;
  }

  const int stiles = scan_num_tiles(N);
  HeadIn hin{lay, B, sorted_keys};
  SegOut sout{lay, B, N, seg_start, seg_meta, sorted_keys, perm};
  scan_tile_sums_kernel<<<stiles, kScanThreads, 0, st>>>(hin, N, tile_sums);
  PTREC_LAUNCH_CHECK("scan_tile_sums_kernel(heads)");
  scan_top_kernel<<<1, 1024, 0, st>>>(tile_sums, stiles, n_seg);
  PTREC_LAUNCH_CHECK("scan_top_kernel(heads)");
  scan_apply_kernel<<<stiles, kScanThreads, 0, st>>>(hin, sout, N, tile_sums);
  PTREC_LAUNCH_CHECK("scan_apply_kernel(heads)");
  return PTREC_OK;
}
