/*
 * ptrec_b200.h — C ABI of libptrec_b200.so, the sm_100a (B200) hot path of PyTorchRec.
 *
 * The reference (Troublem1/PyTorchRec) has no FFI: its hot path is the Python convention
 *   nn.Embedding(column.category_num, D)[column.get_feature_data(batch)]  -> interaction -> loss
 *   -> autograd embedding_dense_backward -> dense optimizer.step()
 * (torchrec/model/FunkSVD.py:39-51, SVDPP.py:36-66, IModel.py:116-125).  Each entry point below
 * names the reference lines whose work it replaces.  A maintainer binds these with ctypes
 * (see INTEGRATION.md); pytorchrec_b200/_lib.py is that binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - all buffers are caller-owned; the library never allocates memory that outlives a call;
 *   - scratch is supplied by the caller: ask `*_workspace_bytes()` first;
 *   - calls enqueue work on `stream` (a cudaStream_t passed as void*) and never synchronise;
 *   - return value: 0 on success, a negative PTREC_E* code otherwise; text via ptrec_last_error().
 *   - tables and optimizer state are updated IN PLACE.
 */
#ifndef PTREC_B200_H
#define PTREC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTREC_ABI_VERSION 26

/* error codes */
#define PTREC_OK 0
#define PTREC_EINVAL (-1)       /* bad argument                                   */
#define PTREC_EALIGN (-2)       /* pointer / stride not aligned as required       */
#define PTREC_EUNSUPPORTED (-3) /* unsupported D / dtype / size                   */
#define PTREC_ECUDA (-4)        /* a CUDA runtime call failed                     */
#define PTREC_EWORKSPACE (-5)   /* workspace too small                            */

/* table element types (the `dtype` argument of K1 / K2).  PTREC_BF16: rows stored as bf16 — half the HBM bytes per
 * lookup; row_stride counts ELEMENTS.  K1 widens the row exactly to fp32 (outputs, pooling and every gradient stay
 * fp32); K2b widens the row, applies the optimizer in fp32 against fp32 state tensors (same row stride in elements, not
 * interleaved with the weights) and rounds the row back to bf16 (nearest even).  The row-wise sharded entry points
 * take fp32 tables only. */
#define PTREC_F32 0
#define PTREC_BF16 1

/* pooling modes: reference idioms SVDPP.py:52-55 (sum / sqrt(count)), SASRec.py:109-110 (mean) */
#define PTREC_POOL_SUM 0
#define PTREC_POOL_MEAN 1
#define PTREC_POOL_SQRTN 2

/* which slots of a padded [B, L] bag are valid */
#define PTREC_MASK_NONE 0           /* every slot (one-hot fields, candidate lists)                  */
#define PTREC_MASK_PAD 1            /* id != 0              SVDPP.py:49                              */
#define PTREC_MASK_PAD_KEEP_FIRST 2 /* id != 0 or slot 0    model/utils.py:5-10 (get_valid_his_index) */
#define PTREC_MASK_LENS 3           /* slot < lens[b]       SASRec.py:109-110 / HistoryDataReader     */

/* feature flags */
#define PTREC_FEAT_NEG_IS_PAD 1 /* ids < 0 are empty slots (fixed-capacity all-to-all lists): skipped, no error */

/* fused optimizers */
#define PTREC_OPT_SGD 0
#define PTREC_OPT_ADAGRAD 1
#define PTREC_OPT_ROWWISE_ADAGRAD 2
#define PTREC_OPT_LAZY_ADAM 3

/*
 * One sparse feature (= one CategoricalColumn feeding one table).  A batch of B samples carries,
 * for feature f, a padded id matrix [B, bag_len] (bag_len == 1 for one-hot fields), stored at
 * ids + id_base * B.  Features MUST be ordered by `table` so that every table's lookups are one
 * contiguous range of `ids` (the segmented sort relies on it).  40 bytes, no padding.
 */
typedef struct ptrec_feature_desc {
  int32_t table;     /* index into table_ptrs                                              */
  int32_t bag_len;   /* L: padded bag length, >= 1                                         */
  int32_t pooling;   /* PTREC_POOL_*                                                        */
  int32_t mask_mode; /* PTREC_MASK_*                                                        */
  int32_t lens_col;  /* row of `lens` ([n_cols, B] int32) used by PTREC_MASK_LENS, else -1  */
  int32_t flags;     /* PTREC_FEAT_* bits                                                  */
  int64_t id_base;   /* sum of bag_len over the features before this one                    */
  int64_t out_col;   /* float offset of this feature's pooled vector inside an output row   */
} ptrec_feature_desc;

/* One segment (= one unique (table, row) touched by the batch) as emitted by ptrec_sort_dedup: everything the
 * fused update needs about the segment in one 16-byte record (one 128-bit load). */
typedef struct ptrec_segment_meta {
  uint32_t key;       /* row id (0xFFFFFFFF = the table's run of masked slots)   */
  int32_t first_slot; /* perm[seg_start[u]]: slot of the segment's first lookup   */
  int32_t table;      /* table index                                              */
  int32_t reserved;
} ptrec_segment_meta;

/* hyper-parameters of the fused row update; passed by value (host memory) */
typedef struct ptrec_optim_args {
  int32_t kind;  /* PTREC_OPT_* */
  int32_t step;  /* 1-based step count (Adam bias correction, Adagrad lr_decay)             */
  float lr;
  float eps;
  float beta1;
  float beta2;
  float weight_decay; /* L2 on touched rows only (g += wd * w); 0 for dense-parity         */
  float lr_decay;     /* Adagrad: clr = lr / (1 + (step-1) * lr_decay)                      */
} ptrec_optim_args;

int ptrec_abi_version(void);
const char* ptrec_last_error(void); /* thread-local, valid until the next failing call */
int64_t ptrec_launch_count(void);   /* kernels launched by this library since load (process-wide) */
/* Gradient-finalising reductions off the critical path: while a stream is set, the entry points that end in a reduction
 * of per-CTA partials into a bias / weight-vector gradient (ptrec_tc_gemm_split2h_fused with colsum, ptrec_rowdot_bwd_h2,
 * ptrec_fm_head_bwd) launch that reduction on it, behind an event recorded on the producer's stream.  The caller joins the
 * stream before it reads those gradients and gives every call in flight its own workspace.  NULL (default): reductions
 * follow their producer on its stream. */
void ptrec_set_reduce_stream(void* stream);
/* cudaLimitMaxL2FetchGranularity of the current device (32 / 64 / 128 bytes): random row reads narrower
 * than the limit over-fetch from HBM.  Affects the whole device context; the caller decides. */
int ptrec_set_l2_fetch_granularity(int32_t bytes);
int ptrec_get_l2_fetch_granularity(void);

/* ---------------------------------------------------------------------------------------------
 * a1/a6 index preparation.  Replaces the mask/len arithmetic of SVDPP.py:49,53 and
 * model/utils.py:5-10: turns one padded id matrix into CSR (ids of valid slots, in order, and
 * bag offsets).  Bit-exact integer work.
 *   ids_padded [B, L] int64; lens [B] int32 or NULL; mask_mode PTREC_MASK_*;
 *   out_ids [B*L] (first out_offsets[B] entries meaningful); out_offsets [B+1] int64.
 * workspace: ptrec_index_prep_workspace_bytes(B).
 */
size_t ptrec_index_prep_workspace_bytes(int64_t B);
int ptrec_index_prep(const int64_t* ids_padded, const int32_t* lens, int64_t B, int64_t L,
                     int32_t mask_mode, int64_t* out_ids, int64_t* out_offsets, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1 multi-table embedding gather + sum/mean/sqrtn pooling, forward.
 * Replaces T x `nn.Embedding.forward` (FunkSVD.py:47-48, SVDPP.py:50,57-61, NCF.py:62-65) and the
 * masked pooling chains SVDPP.py:49-55 / SASRec.py:109-110 in ONE launch per bag class.
 *   table_ptrs [T] device array of table base pointers ([rows_t, D] row-major, 16-byte aligned)
 *   row_stride floats between consecutive rows of every table (== D for plain tables; 2*D / 4*D when the
 *              optimizer state is interleaved with the weights so that one 128-byte DRAM line holds both)
 *   table_rows [T] int64 (bounds check -> *err_flag = 1 on an out-of-range id, lookup skipped)
 *   ids        [sum_f B*bag_len_f] int64, feature-major, each feature [B, bag_len]
 *   lens       [n_cols, B] int32 or NULL
 *   out        [B, out_row_stride] float32: out[b*out_row_stride + out_col_f + d]
 *   bag_scale  [F, B] float32 or NULL: written with the factor the bag sum was multiplied by
 *              (1, 1/count, 1/sqrt(count); count clamped to >= 1) — the backward needs it.
 *   err_flag   int32 device word or NULL.
 *   feats      [F] device copy of the descriptors; feats_host the same array in host memory
 *              (launch geometry is derived from it without touching the device).
 */
int ptrec_embedding_gather_pool_fwd(const void* const* table_ptrs, const int64_t* table_rows,
                                    int32_t T, int32_t D, int64_t row_stride, int32_t dtype,
                                    const ptrec_feature_desc* feats,
                                    const ptrec_feature_desc* feats_host, int32_t F,
                                    const int64_t* ids, const int32_t* lens, int64_t B, float* out,
                                    int64_t out_row_stride, float* bag_scale, int32_t* err_flag,
                                    void* stream);

/* K1 over ROW-WISE SHARDS in peer memory (C1, forward): table t is split over G GPUs of one NVLink domain,
 * owner(id) = id mod G, local row = id div G; shard_ptrs [T][G] holds the base pointer of every shard as seen from
 * THIS device (its own shard plus the peers' through symmetric / IPC mappings), so the row loads travel over
 * NVLink / NVSwitch inside the gather kernel and the forward needs no collective.  table_rows [T] = GLOBAL row
 * counts.  One-hot fields only (bag_len 1, PTREC_MASK_NONE).  The caller orders the owners' updates before these
 * reads (one collective per step, see pytorchrec_b200/distributed/sharded.py).  No reference counterpart
 * (single-device reference, torchrec/task/Task.py:187-190). */
int ptrec_embedding_gather_pool_fwd_sharded(const void* const* shard_ptrs, const int64_t* table_rows,
                                            int32_t T, int32_t G, int32_t D, int64_t row_stride, int32_t dtype,
                                            const ptrec_feature_desc* feats,
                                            const ptrec_feature_desc* feats_host, int32_t F,
                                            const int64_t* ids, int64_t B, float* out, int64_t out_row_stride,
                                            int32_t* err_flag, void* stream);

/* K2a has two code paths with identical outputs: one CTA per table entirely in shared memory (2 launches), and
 * the multi-launch global radix sort.  mode 0 = global only, 1 = automatic (shared memory for small batches, where
 * launch latency dominates; default), 2 = shared memory whenever a table's batch fits (<= 22528 slots). */
void ptrec_set_smem_sort(int32_t mode);
int32_t ptrec_smem_sort_enabled(void);
/* Large batches (beyond the shared-memory path): on = the one-sweep radix sort — keys and the digit histograms of
 * every pass in one launch, one decoupled-look-back launch per digit, one look-back dedup launch (P + 2 launches
 * + a memset, instead of 3 P + 3); off = the histogram / scan / scatter launches per pass.  Identical outputs. */
void ptrec_set_one_sweep_sort(int32_t on);
int32_t ptrec_one_sweep_sort_enabled(void);

/* ---------------------------------------------------------------------------------------------
 * K2a segmented sort + dedup of the lookups of one batch (the integer half of the backward).
 * Replaces the scatter order of aten::embedding_dense_backward (implicit at IModel.py:123).
 * Positions p in [0, N), N = B * sum_f bag_len_f, are the slots of `ids`.  Per table, slots are
 * stably sorted by id (masked slots get key 0xFFFFFFFF and sort last); equal-id runs are the
 * segments.  Bit-exact against torch.sort(stable=True) / torch.unique(sorted=True) per table.
 *   sorted_keys [N] uint32   id of the slot at sorted position j (0xFFFFFFFF = masked)
 *   perm        [N] int32    slot p at sorted position j
 *   seg_start   [N+1] int32  sorted position where segment u starts; seg_start[n_seg] = N
 *   seg_meta    [N] ptrec_segment_meta   key / first slot / table of segment u
 *   n_seg       [1] int32    number of segments (masked runs included, one per table at most)
 * workspace: ptrec_sort_dedup_workspace_bytes(N, T).  max_rows_host = max_t rows_t (bounds key bits);
 * table_rows [T] device int64: ids outside [0, rows_t) are treated as masked.
 */
size_t ptrec_sort_dedup_workspace_bytes(int64_t N, int32_t T);
int ptrec_sort_dedup(const ptrec_feature_desc* feats, const ptrec_feature_desc* feats_host, int32_t F,
                     int32_t T, const int64_t* table_rows, int64_t max_rows_host,
                     const int64_t* ids, const int32_t* lens, int64_t B, uint32_t* sorted_keys,
                     int32_t* perm, int32_t* seg_start, ptrec_segment_meta* seg_meta, int32_t* n_seg,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2b segment-reduce of the pooled-output gradient + fused row update.  Replaces
 * aten::embedding_dense_backward + zero_grad + the dense optimizer sweep over every row
 * (IModel.py:122-124, optim/optimizers.py:7-11, torch.optim.SGD/Adagrad/SparseAdam arithmetic).
 * One pass per unique row: sum its gradient slots in sorted (= batch) order, read-modify-write
 * the weight row and its state.  No [rows, D] gradient tensor is ever materialised.
 *   row_stride: floats between consecutive rows of the tables AND of the element-wise state arrays
 *     (interleaved layout: state1_ptrs[t] = table_ptrs[t] + D, row_stride = 2*D).
 *   state1_ptrs / state2_ptrs [T] device arrays (NULL when the optimizer has no such state):
 *     ADAGRAD: state1 = sum of squares [rows, D];  ROWWISE_ADAGRAD: state1 = [rows];
 *     LAZY_ADAM: state1 = exp_avg, state2 = exp_avg_sq, both [rows, D].
 *   grad_out [B, grad_row_stride] float32, same column layout as the forward `out`.
 *   bag_scale as written by the forward (NULL = all ones).
 * workspace: ptrec_embedding_bwd_workspace_bytes(N, D).
 */
size_t ptrec_embedding_bwd_workspace_bytes(int64_t N, int32_t D);
int ptrec_embedding_bwd_fused(void* const* table_ptrs, void* const* state1_ptrs,
                              void* const* state2_ptrs, int32_t T, int32_t D, int64_t row_stride, int32_t dtype,
                              const ptrec_feature_desc* feats,
                              const ptrec_feature_desc* feats_host, int32_t F, int64_t B,
                              const uint32_t* sorted_keys, const int32_t* perm,
                              const int32_t* seg_start, const ptrec_segment_meta* seg_meta,
                              const int32_t* n_seg, const float* grad_out,
                              int64_t grad_row_stride, const float* bag_scale,
                              const ptrec_optim_args* opt_host, void* workspace,
                              size_t workspace_bytes, void* stream);
/* the four named entry points of SURVEY.md §8b: same arguments as ptrec_embedding_bwd_fused; each checks
 * opt_host->kind and forwards */
#define PTREC_BWD_FUSED_ARGS                                                                                   \
  void* const* table_ptrs, void* const* state1_ptrs, void* const* state2_ptrs, int32_t T, int32_t D,           \
      int64_t row_stride, int32_t dtype, const ptrec_feature_desc* feats, const ptrec_feature_desc* feats_host, \
      int32_t F, int64_t B, const uint32_t* sorted_keys, const int32_t* perm, const int32_t* seg_start,        \
      const ptrec_segment_meta* seg_meta, const int32_t* n_seg, const float* grad_out, int64_t grad_row_stride, \
      const float* bag_scale, const ptrec_optim_args* opt_host, void* workspace, size_t workspace_bytes,       \
      void* stream
int ptrec_embedding_bwd_fused_sgd(PTREC_BWD_FUSED_ARGS);
int ptrec_embedding_bwd_fused_adagrad(PTREC_BWD_FUSED_ARGS);
int ptrec_embedding_bwd_fused_rowwise_adagrad(PTREC_BWD_FUSED_ARGS);
int ptrec_embedding_bwd_fused_lazy_adam(PTREC_BWD_FUSED_ARGS);

/* Segment-reduce only (no update): writes the per-unique-row gradient sums so that a stock sparse
 * optimizer (torch.optim.SparseAdam / SGD) can consume them.  row_grad [N, D] float32 (first
 * n_seg rows meaningful; masked segments are written as zeros). */
int ptrec_embedding_bwd_segment_sum(int32_t T, int32_t D, const ptrec_feature_desc* feats,
                                    const ptrec_feature_desc* feats_host, int32_t F, int64_t B,
                                    const uint32_t* sorted_keys, const int32_t* perm,
                                    const int32_t* seg_start, const ptrec_segment_meta* seg_meta,
                                    const int32_t* n_seg, const float* grad_out,
                                    int64_t grad_row_stride, const float* bag_scale,
                                    float* row_grad, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3 FM second-order interaction  y[b] = 0.5 * sum_k ((sum_f v[b,f,k])^2 - sum_f v[b,f,k]^2).
 * Generalises the two-field dot of FunkSVD.py:51,62 / SVDPP.py:65 to F fields.
 *   v [B, F, D] float32 with row stride v_row_stride (floats);  y [B].
 * backward: grad_v[b,f,k] = gy[b] * (S[b,k] - v[b,f,k]) (+ grad_in[b,f,k] when grad_in != NULL,
 * which fuses autograd's accumulation of the DNN-branch gradient).
 */
int ptrec_fm2_fwd(const float* v, int64_t v_row_stride, int64_t B, int32_t F, int32_t D, float* y,
                  void* stream);
int ptrec_fm2_bwd(const float* v, int64_t v_row_stride, const float* gy, const float* grad_in,
                  int64_t grad_in_row_stride, int64_t B, int32_t F, int32_t D, float* grad_v,
                  int64_t grad_v_row_stride, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4 DIN attention pooling: activation unit + masked weighted sum, fused (nothing of shape [B, L, 4*DQ] or
 * [B, L, H1] reaches HBM).  Not in the reference; the masking / history conventions are the reference's
 * (length column clipped to >= 1, right-padded histories: HistoryDataReader.py:55-69, SASRec.py:83,109-110).
 *   a_l = W3 relu(W2 relu(W1 [q, k_l, q-k_l, q*k_l] + b1) + b2) + b3 ;  pooled[b] = sum_{l < lens[b]} a_l k_l
 *   q [B, DQ] (row stride q_stride), keys [B, L, DQ] (strides in floats), lens [B] int32 or NULL (= L)
 *   W1 [H1, 4*DQ], b1 [H1], W2 [H2, H1], b2 [H2], W3 [1, H2], b3 [1]  (nn.Linear layouts), fp32
 *   out [B, DQ];  scores [B, L] or NULL (a_l, 0 beyond the length)
 * backward: g_q [B, DQ], g_keys [B, L, DQ] (zeros beyond the length), grad_params = flat
 *   [gW1 | gb1 | gW2 | gb2 | gW3 | gb3] of ptrec_din_attn_pool_grad_floats() floats.
 * Built for DQ in {16, 32} and (H1, H2) in {(80, 40), (64, 32)}.
 */
int ptrec_din_attn_pool_fwd(const float* q, int64_t q_stride, const float* keys, int64_t k_stride_b,
                            int64_t k_stride_l, const int32_t* lens, int64_t B, int32_t L, int32_t DQ, int32_t H1,
                            int32_t H2, const float* W1, const float* b1, const float* W2, const float* b2,
                            const float* W3, const float* b3, float* out, float* scores, void* stream);
int32_t ptrec_din_attn_pool_grad_floats(int32_t DQ, int32_t H1, int32_t H2);
size_t ptrec_din_attn_pool_bwd_workspace_bytes(int64_t B, int32_t DQ, int32_t H1, int32_t H2);
/* The forward has two builds with the same contract: fp32 CUDA cores (din_attn.cu) and tcgen05 tensor cores with
 * fp32-faithful fp16 x 2 operand planes (din_attn_tc.cu: both hidden layers as M = 128 MMAs over tiles of 128
 * positions); the backward likewise (recompute + four gradient contractions as MMAs, operand tiles read in both
 * orientations).  mode bit 0: tensor-core forward, bit 1: tensor-core backward (default 3: both); where the shape has no
 * tensor-core build the fp32 kernel runs. */
void ptrec_set_din_tc(int32_t mode);
int32_t ptrec_din_tc_enabled(void);
int ptrec_din_attn_pool_bwd(const float* q, int64_t q_stride, const float* keys, int64_t k_stride_b,
                            int64_t k_stride_l, const int32_t* lens, int64_t B, int32_t L, int32_t DQ, int32_t H1,
                            int32_t H2, const float* W1, const float* b1, const float* W2, const float* b2,
                            const float* W3, const float* b3, const float* g_pooled, float* g_q, float* g_keys,
                            int64_t gk_stride_b, int64_t gk_stride_l, float* grad_params, void* workspace,
                            size_t workspace_bytes, void* stream);

/* K4 with the key gather fused in (round 2; tensor-core builds only: PTREC_EUNSUPPORTED otherwise, and the caller gathers).
 * The keys of the attention unit are rows of two fp32 tables (DIN: item and category), DQ / 2 columns each:
 *   key (b, l) = [ table0[ids0[b * ids_stride_b + ids_offset + l]] | table1[ids1[b * ids_stride_b + ids_offset + l]] ]
 * row_stride = the tables' row pitch in floats (a table interleaved with its optimizer state has pitch 2D or 4D); an id
 * outside [0, rows) raises *err_flag (may be NULL) and reads as a zero row.  q, the outputs and the workspace are those of
 * ptrec_din_attn_pool_fwd / _bwd; g_keys still leaves as dense rows (the fused table update consumes them). */
int ptrec_din_attn_pool_fwd_ids(const float* q, int64_t q_stride, const float* table0, const float* table1,
                                int64_t row_stride0, int64_t row_stride1, int64_t rows0, int64_t rows1, const int64_t* ids0,
                                const int64_t* ids1, int64_t ids_stride_b, int64_t ids_offset, int32_t* err_flag,
                                const int32_t* lens, int64_t B, int32_t L, int32_t DQ, int32_t H1, int32_t H2,
                                const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                                const float* b3, float* out, float* scores, void* stream);
int ptrec_din_attn_pool_bwd_ids(const float* q, int64_t q_stride, const float* table0, const float* table1,
                                int64_t row_stride0, int64_t row_stride1, int64_t rows0, int64_t rows1, const int64_t* ids0,
                                const int64_t* ids1, int64_t ids_stride_b, int64_t ids_offset, const int32_t* lens, int64_t B,
                                int32_t L, int32_t DQ, int32_t H1, int32_t H2, const float* W1, const float* b1,
                                const float* W2, const float* b2, const float* W3, const float* b3, const float* g_pooled,
                                float* g_q, float* g_keys, int64_t gk_stride_b, int64_t gk_stride_l, float* grad_params,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K5 DCN-v2 cross layer on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM).
 * Not in the reference; the layer is  x_{l+1} = x0 (.) (x_l W^T + b) + x_l  with W stored
 * [d_out, d_in] row-major like nn.Linear.  All activations are bf16 [B, ld] row-major, d % 8 == 0 and
 * ld % 8 == 0 (pad the concatenated input once); pointers 16-byte aligned.
 *   fwd    out = x0 (.) u + x_l,  u = x_l W^T + bias;  u_out (may be NULL) receives u for the backward
 *   dgrad  g_x = g_u W + g_out with g_u = g_out (.) x0 supplied by the caller; weight_t = W^T [d_in, d_out];
 *          g_u_prev (may be NULL) receives g_x (.) x0, i.e. the g_u of the layer below
 *   wgrad  grad_w [d_out, d_in] fp32 = g_u^T x_l   (workspace: ptrec_dcn_cross_wgrad_workspace_bytes)
 */
/* Round 2: the three entry points run on the CTA-pair GEMM kernel of K6 with one bf16 operand plane (256 x 256 pair tiles,
 * TMA-store epilogue with the cross-layer arithmetic fused, MN-major operands for the weight gradient); the 128 x 128
 * single-CTA kernel of round 1 stays selectable: */
void ptrec_set_dcn_2sm(int32_t enabled);
int32_t ptrec_dcn_2sm_enabled(void);
int ptrec_dcn_cross_fwd(const void* x_l, const void* x0, const void* weight, const float* bias, int64_t B,
                        int32_t d, int64_t ld, void* out, void* u_out, void* stream);
int ptrec_dcn_cross_dgrad(const void* g_u, const void* weight_t, const void* g_out, const void* x0, int64_t B,
                          int32_t d, int64_t ld, void* g_x, void* g_u_prev, void* stream);
size_t ptrec_dcn_cross_wgrad_workspace_bytes(int64_t B, int32_t d);
int ptrec_dcn_cross_wgrad(const void* g_u, const void* x_l, int64_t B, int32_t d, int64_t ld, float* grad_w,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Element-wise companions of the cross GEMMs (csrc/dcn_glue.cu), one launch each; dp = d rounded up to 8; every bf16
 * [B, dp] operand is dense (row pitch dp) and 16-byte aligned.
 *   prep_weight   W fp32 [d, d], b fp32 [d] -> w16 bf16 [dp, dp], w16t = its transpose, bias_pad fp32 [dp] (zero padded)
 *   pack_input    x fp32 [B, d] (pitch ldx) -> out bf16 [B, dp] (zero padded);   unpack: the inverse into fp32 [B, d]
 *   bwd_init      g fp32 [B, d] (pitch ldg) -> g_out = bf16(g) padded, g_u = g_out * x0
 *   bwd_layer     g_x0 fp32 [B, dp] (+)= g_out * u;  grad_bias fp32 [d] = column sums of g_u (fixed order)
 *   bwd_final     out fp32 [B, d] = g_x0 + g_out */
int ptrec_dcn_prep_weight(const float* W, const float* b, int32_t d, int32_t dp, void* w16, void* w16t, float* bias_pad,
                          void* stream);
int ptrec_dcn_pack_input(const float* x, int64_t ldx, int64_t B, int32_t d, int32_t dp, void* out, void* stream);
int ptrec_dcn_unpack(const void* x, int64_t B, int32_t d, int32_t dp, float* out, void* stream);
int ptrec_dcn_bwd_init(const float* g, int64_t ldg, const void* x0, int64_t B, int32_t d, int32_t dp, void* g_out,
                       void* g_u, void* stream);
size_t ptrec_dcn_bwd_layer_workspace_bytes(int64_t B, int32_t dp);
int ptrec_dcn_bwd_layer(const void* g_out, const void* u, const void* g_u, int64_t B, int32_t d, int32_t dp, float* g_x0,
                        int32_t accumulate, float* grad_bias, void* workspace, size_t workspace_bytes, void* stream);
int ptrec_dcn_bwd_final(const float* g_x0, const void* g_out, int64_t B, int32_t d, int32_t dp, float* out, void* stream);
/* The cross half of DCN's closing Linear(concat(cross, deep)) -> 1, read from / written to the bf16 tensors of the cross
 * layers (no fp32 [B, d] copy of x_L, no fp32 [B, d] gradient):
 *   head_fwd  y[b] = sum_{c < d} x_L[b, c] w[c]
 *   head_bwd  g_out = bf16(g_y (x) w) zero padded, g_u = g_out * x0, grad_w[c] = sum_b g_y[b] x_L[b, c]
 *             (workspace: ptrec_dcn_bwd_layer_workspace_bytes(B, dp); fixed-order sums) */
int ptrec_dcn_head_fwd(const void* x_l, int64_t B, int32_t d, int32_t dp, const float* w, float* y, void* stream);
int ptrec_dcn_head_bwd(const float* g_y, const float* w, const void* x_l, const void* x0, int64_t B, int32_t d, int32_t dp,
                       void* g_out, void* g_u, float* grad_w, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K8 the FM head: first-order sum + second-order interaction + dense-feature linear term + bias in one pass, and
 * (DeepFM) the tower input [v | x] written by the same pass; the Linear(H, 1) that closes the tower as a row dot.
 * Written in the reference's idiom these are  biases + dot (SVDPP.py:65-66),  concat -> MLP -> Linear(., 1)
 * (NCF.py:68-74);  here they replace ~20 latency-bound torch launches per step.
 *   v [B, F*D] (row stride given), w1 [B, F] first-order values or NULL, x [B, nd] dense features or NULL,
 *   wd [nd] or NULL, bias [1] or NULL;  logit [B];  deep_in [B, >= F*D+nd] or NULL;  deep_in_planes (or NULL): the
 *   tower input again as the exact bf16 planes [3][B][ld] K6 consumes (ld a multiple of 8 >= F*D+nd, pad = 0).
 * Backward: g [B] = d loss / d logit, g_deep_in = gradient of the tower input or NULL (added on the fly);
 *   grad_v [B, F*D], grad_w1 [B, F] or NULL, grad_x [B, nd] or NULL, grad_wd [nd] / grad_bias [1] or NULL
 *   (two-level fixed-order sums; workspace ptrec_fm_head_bwd_workspace_bytes).
 * Needs D a power of two in [4, 128], F*D <= 2048, nd <= 128 (ptrec_fm_head_supported); callers fall back to K3 +
 * library ops otherwise.
 */
int ptrec_fm_head_supported(int32_t F, int32_t D, int32_t nd);
int ptrec_fm_head_fwd(const float* v, int64_t v_row_stride, const float* w1, int64_t w1_row_stride, const float* x,
                      int64_t x_row_stride, const float* wd, const float* bias, int64_t B, int32_t F, int32_t D,
                      int32_t nd, float* logit, float* deep_in, int64_t deep_in_row_stride, void* deep_in_planes,
                      int64_t deep_in_planes_ld, void* stream);
size_t ptrec_fm_head_bwd_workspace_bytes(int32_t nd);
int ptrec_fm_head_bwd(const float* v, int64_t v_row_stride, const float* x, int64_t x_row_stride, const float* wd,
                      const float* g, const float* g_deep_in, int64_t g_deep_in_row_stride, int64_t B, int32_t F,
                      int32_t D, int32_t nd, float* grad_v, int64_t grad_v_row_stride, float* grad_w1, float* grad_x,
                      int64_t grad_x_row_stride, float* grad_wd, float* grad_bias, void* workspace,
                      size_t workspace_bytes, void* stream);
/* y[b] = h[b, :] . w  (h [B, H], w [H]);  backward: grad_h = g (x) w (or NULL), grad_w = h^T g (or NULL) */
/* K8 writing / reading the K6 fused tower's operand planes directly (fp16 x 2, carried scales: see "the fused tower"):
 * ptrec_fm_head_fwd_h2: ptrec_fm_head_fwd whose tower input leaves as two fp16 planes [2][B][planes_ld] split with
 *   *scale (deep_in may be NULL: the tower then reads the planes only); *max_out is raised to max |tower input|.
 * ptrec_rowdot_bwd_h2: ptrec_rowdot_bwd where h is the ReLU output of the tower's last layer: instead of fp32 g_h it
 *   writes the planes of (g (x) w) * (h > 0) split with *scale — the gradient of that layer's pre-activation — its
 *   column sums (colsum [H] or NULL: the layer's bias gradient, fixed order) and raises *max_out to its |max|.
 *   workspace: 2 x ptrec_rowdot_bwd_workspace_bytes(H). */
int ptrec_fm_head_fwd_h2(const float* v, int64_t v_row_stride, const float* w1, int64_t w1_row_stride, const float* x,
                         int64_t x_row_stride, const float* wd, const float* bias, int64_t B, int32_t F, int32_t D,
                         int32_t nd, float* logit, float* deep_in, int64_t deep_in_row_stride, void* planes,
                         int64_t planes_ld, const float* scale, float* max_out, void* stream);
int ptrec_rowdot_bwd_h2(const float* h, int64_t h_row_stride, const float* w, const float* g, int64_t B, int32_t H,
                        void* planes, int64_t planes_ld, const float* scale, float* max_out, float* colsum,
                        float* grad_w, void* workspace, size_t workspace_bytes, void* stream);
int ptrec_rowdot_supported(int32_t H);
int ptrec_rowdot_fwd(const float* h, int64_t h_row_stride, const float* w, int64_t B, int32_t H, float* y,
                     void* stream);
size_t ptrec_rowdot_bwd_workspace_bytes(int32_t H);
int ptrec_rowdot_bwd(const float* h, int64_t h_row_stride, const float* w, const float* g, int64_t B, int32_t H,
                     float* grad_h, int64_t grad_h_row_stride, float* grad_w, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * K7 dense-parameter update in one launch (the dense half of optimizer.step(), IModel.py:124 ->
 * torchrec/optim/optimizers.py:7-11 -> torch.optim.SGD / Adagrad / Adam arithmetic, same operation order).
 *   tensors     [n_tensors] device array of {param, grad, state1, state2, numel} (contiguous fp32 tensors;
 *               state1 = Adagrad sum | Adam exp_avg, state2 = Adam exp_avg_sq, NULL where unused)
 *   chunk_start [n_tensors + 1] int32 device array: first CTA of each tensor, ptrec_dense_optim_chunk() elements per
 *               CTA; n_chunks = chunk_start[n_tensors]
 *   args        host; kind PTREC_OPT_SGD | PTREC_OPT_ADAGRAD | PTREC_OPT_LAZY_ADAM (= dense Adam here); step >= 1
 */
typedef struct ptrec_dense_tensor {
  void* param;
  const void* grad;
  void* state1;
  void* state2;
  int64_t numel;
} ptrec_dense_tensor;
int32_t ptrec_dense_optim_chunk(void);
int ptrec_dense_optim_step(const ptrec_dense_tensor* tensors, const int32_t* chunk_start, int32_t n_tensors,
                           int32_t n_chunks, const ptrec_optim_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * C2 cross-GPU ordering and the dense-gradient all-reduce over NVLink peer memory (csrc/peer_sync.cu).  No reference
 * counterpart (single-device reference, torchrec/task/Task.py:187-190); replaces the NCCL launches of the row-wise
 * sharded step (a 1-element all_reduce used as a fence; cat + all_reduce + div + copy-back of the dense gradients).
 *   ptrec_peer_barrier  every rank's work enqueued before it on its stream is complete and visible (system scope)
 *     before any rank's work enqueued after it starts.  peer_flags [G] device array: rank p's flag array
 *     (uint32 [ptrec_peer_sync_slots()][ptrec_peer_sync_max_ranks()], symmetric memory, zero-initialised) as seen
 *     from this GPU; local_epoch uint32 [slots] device (zero-initialised, private).  Every rank must issue the same
 *     sequence of calls per slot.  Graph-capturable (the epoch lives in device memory).
 *   ptrec_dense_pack    gradients named by a K7 descriptor table -> contiguous `stage` (tensor t at element
 *     chunk_start[t] * ptrec_dense_optim_chunk(), zero padded): n_chunks * chunk floats.
 *   ptrec_dense_optim_step_reduce  K7 with gradient = grad_scale * sum_r peer_stage[r][.] (rank order): the
 *     all-reduce fused into the optimizer step.  peer_stage [G] device array of the ranks' stages.
 */
int32_t ptrec_peer_sync_max_ranks(void);
int32_t ptrec_peer_sync_slots(void);
int ptrec_peer_barrier(uint32_t* const* peer_flags, uint32_t* local_epoch, int32_t slot, int32_t G, int32_t my_rank,
                       void* stream);
int ptrec_dense_pack(const ptrec_dense_tensor* tensors, const int32_t* chunk_start, int32_t n_tensors,
                     int32_t n_chunks, float* stage, void* stream);
int ptrec_dense_optim_step_reduce(const ptrec_dense_tensor* tensors, const int32_t* chunk_start, int32_t n_tensors,
                                  int32_t n_chunks, const ptrec_optim_args* args, const float* const* peer_stage,
                                  int32_t G, float grad_scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K6 fp32-faithful Linear layers of the DNN tower on tcgen05 (replaces the fp32 cuBLAS sgemm behind nn.Linear in
 * torchrec/model/layer/Dense.py:9-17 / MLP.py:13-21 and its backward).  Operands are fp32 matrices split exactly
 * into three bf16 planes (x = x0 + x1 + x2); products use the six plane pairs of weight >= 2^-16 with fp32
 * accumulation, i.e. fp32-level error (the north star's 1e-5) at tensor-core speed.
 *
 * ptrec_tc_split3: src fp32 [R, C] (pitch ld) ->
 *   planes    [3][R][pl_ld] bf16 (or NULL)   row-major planes: the A operand of  y = x W^T  and of  dx = g W
 *   planes_t  [3][C][pt_ld] bf16 (or NULL)   transposed planes: both operands of dW = g^T x (K = batch)
 *   relu_ref  fp32 [R, C] (pitch ld_ref) or NULL: src is zeroed where relu_ref <= 0 (ReLU backward fused)
 *   colsum    fp32 [C] or NULL: column sums of the (masked) src = bias gradient, fixed summation order;
 *             needs workspace ptrec_tc_split3_workspace_bytes(R, C)
 * Plane pitches are multiples of 8 elements; pad columns are written as zero.
 *
 * ptrec_tc_gemm_split3: out[M, N] (pitch ldo, multiple of 4, >= N rounded up to 4) =
 *   A[M, K] B[N, K]^T (+ bias[N]) (ReLU)   with A, B given as planes [3][M][lda], [3][N][ldb].
 *   out_planes (or NULL): the result written a second time as bf16 planes [3][M][out_planes_ld] — the next
 *   layer's A operand, so that layer needs no split pass (out_planes_ld a multiple of 8 >= N; splits == 1 only).
 *   splits > 1 cuts K into ranges accumulated through fp32 partials in the workspace
 *   (ptrec_tc_gemm_split3_workspace_bytes) and summed in a fixed order; no bias / ReLU then.
 *   y  = x W^T + b : A = planes(x) [B, K],      B = planes(W) [N, K]
 *   dx = g W       : A = planes(g) [B, N],      B = planes_t(W) [K, N]
 *   dW = g^T x     : A = planes_t(g) [N, B],    B = planes_t(x) [K, B],  splits = ..._default_splits()
 *
 * ptrec_tc_gemm_split3_tn: out[M, N] = A^T B with A given as planes [3][K][lda] (M contiguous) and B as planes
 *   [3][K][ldb] (N contiguous): the MN-major operand form of tcgen05, i.e. the reduction runs over the ROWS of the
 *   stored matrices.   dW = g^T x : A = planes(g) [B, N], B = planes(x) [B, K] — the row-major planes the forward
 *   and the input-gradient GEMMs already use, so the weight gradient needs no transposed copies.
 */
/* The GEMM has a CTA-pair (cta_group::2, 256 x 256 tiles; default) and a single-CTA (128 x 256) kernel with
 * identical results; tests run both. */
void ptrec_tc_set_2sm(int32_t enabled);
void ptrec_tc_set_bk(int32_t bk); /* CTA-pair kernel: K elements per stage, 32 (4 stages) or 64 (2 stages) */
int32_t ptrec_tc_2sm_enabled(void);
size_t ptrec_tc_split3_workspace_bytes(int64_t R, int64_t C);
int ptrec_tc_split3(const float* src, int64_t ld, int64_t R, int64_t C, const float* relu_ref, int64_t ld_ref,
                    void* planes, int64_t pl_ld, void* planes_t, int64_t pt_ld, float* colsum, void* workspace,
                    size_t workspace_bytes, void* stream);
size_t ptrec_tc_gemm_split3_workspace_bytes(int64_t M, int64_t ldo, int32_t splits);
int32_t ptrec_tc_gemm_split3_default_splits(int64_t M, int64_t N, int64_t K);
int ptrec_tc_gemm_split3(const void* a_planes, int64_t M, int64_t lda, const void* b_planes, int64_t N, int64_t ldb,
                         int64_t K, const float* bias, int32_t relu, float* out, int64_t ldo, void* out_planes,
                         int64_t out_planes_ld, int32_t splits, void* workspace, size_t workspace_bytes, void* stream);
int ptrec_tc_gemm_split3_tn(const void* a_planes, int64_t M, int64_t lda, const void* b_planes, int64_t N, int64_t ldb,
                            int64_t K, float* out, int64_t ldo, int32_t splits, void* workspace,
                            size_t workspace_bytes, void* stream);

/* K6, fp16 x 2 operand format (same GEMM kernels, 3 MMAs per product instead of 6, two operand planes instead of three).
 * fp16 has 5 exponent bits, so every tensor is first multiplied by a power of two s that puts its largest magnitude in
 * [2^13, 2^14) (exact); then  x s = h0 + h1 / 2^11,  h0 = fp16(x s),  h1 = fp16((x s - h0) 2^11)  — 2 x 11 = 22 mantissa
 * bits, and h1 is a normal fp16 number whenever h0 is.  The product uses A0 B0 (main accumulator) and A0 B1 + A1 B0
 * (correction accumulator, carrying 2^11); the fp32 epilogue forms (main + 2^-11 corr) / (s_a s_b).  Dropped: A1 B1,
 * 2^-22 relative.
 *
 * ptrec_tc_split2h: as ptrec_tc_split3 with fp16 planes [2][R][pl_ld] / [2][C][pt_ld]; scale_out receives s (one fp32
 *   device word, read by the GEMMs on the device: no host synchronisation).  s comes from the absolute maximum of src
 *   (before the relu_ref mask), found by a first kernel of the same call.  workspace: ptrec_tc_split2h_workspace_bytes
 *   (always needed).
 *   absmax_in (or NULL): one device word holding max |src| already (written by the GEMM that produced src, below);
 *   the call then skips its own pass over src.
 * ptrec_tc_gemm_split2h / _tn: as ptrec_tc_gemm_split3 / _tn on fp16 planes; scale_a / scale_b are the device words the
 *   two split calls wrote.  No output planes.  absmax_out (or NULL; splits == 1 only): one fp32 device word, ZERO on
 *   entry, that the epilogue raises to max |out| (atomicMax on the bit pattern) — pass it as absmax_in when out is
 *   split for the next GEMM. */
/* fp16 x 2 GEMMs on CTA pairs: pair tiles of 256 x 256 (one accumulator pair in TMEM) or 256 x 128 (two accumulator
 * pairs: the MMAs of the next tile overlap the epilogue of the current one).  Identical results up to summation order
 * of nothing — both accumulate a tile's K range in the same order — so tests require bit equality. */
void ptrec_tc_set_bn(int32_t bn); /* 128 or 256 */
int32_t ptrec_tc_get_bn(void);
size_t ptrec_tc_split2h_workspace_bytes(int64_t R, int64_t C);
int ptrec_tc_split2h(const float* src, int64_t ld, int64_t R, int64_t C, const float* relu_ref, int64_t ld_ref,
                     void* planes, int64_t pl_ld, void* planes_t, int64_t pt_ld, float* colsum, float* scale_out,
                     const float* absmax_in, void* workspace, size_t workspace_bytes, void* stream);
int ptrec_tc_gemm_split2h(const void* a_planes, const float* scale_a, int64_t M, int64_t lda, const void* b_planes,
                          const float* scale_b, int64_t N, int64_t ldb, int64_t K, const float* bias, int32_t relu,
                          float* out, int64_t ldo, float* absmax_out, int32_t splits, void* workspace,
                          size_t workspace_bytes, void* stream);
int ptrec_tc_gemm_split2h_tn(const void* a_planes, const float* scale_a, int64_t M, int64_t lda, const void* b_planes,
                             const float* scale_b, int64_t N, int64_t ldb, int64_t K, float* out, int64_t ldo,
                             int32_t splits, void* workspace, size_t workspace_bytes, void* stream);

/* K6, the fused tower (fp16 x 2, CTA-pair kernel).  Between two GEMMs of an MLP (Dense.py:9-17 stacked by MLP.py:13-21)
 * the reference materialises an fp32 activation / gradient; the entry points above then read it back twice (maximum,
 * split).  Here the producing GEMM's epilogue writes its result directly as the consumer's fp16 planes, so no pass runs
 * between two GEMMs.  That needs the power-of-two scale BEFORE the tensor exists, so scales are CARRIED from step to step:
 *
 * slots  fp32 [n_slots][2] = {scale, max}: one slot per tensor of the tower (input, every weight, every hidden
 *   activation, every gradient).  `max` is raised (atomicMax on the bit pattern) by every kernel that writes the
 *   tensor's planes; the caller seeds it with a measured maximum before the first roll.
 * ptrec_tc_scale_roll: once per forward.  scale <- the power of two that puts `max` in [2^5, 2^6) (a tensor may grow
 *   256-fold between two steps before fp16 overflows; elements below 2^-19 of the maximum start losing mantissa bits,
 *   at an absolute error of 2^-41 of the maximum), max <- 0, call_scales[i] <- scale (the array one forward / backward
 *   reads); *err |= 1 if a maximum left the fp16 range under the scale it was split with (the caller polls it).
 * ptrec_tc_split2h_prescaled: ptrec_tc_split2h with the scale read from *scale_in; one kernel, no maximum pass;
 *   *max_out (or NULL) is raised to max |masked src|.
 * ptrec_tc_gemm_split2h_fused: ptrec_tc_gemm_split2h (splits == 1) whose epilogue can also
 *   - write the result as fp16 planes [2][M][out_planes_ld] split with *out_scale (out may then be NULL),
 *   - multiply it by a bit mask first (mask_in: [M][mask_ld] words, bit c%32 of word c/32 = keep; ReLU backward),
 *   - write the mask of its own positive entries (mask_out, same layout; mask_ld a multiple of 4 words >= ceil(N/32)),
 *   - reduce its column sums (colsum [N], fixed order: the bias gradient; workspace ptrec_tc_gemm_fused_workspace_bytes),
 *   - raise *absmax_out to max |masked result| (NOT zeroed by the caller here: it is a slot's `max`).
 *   Stores are staged through shared memory: whole 128-byte row segments per instruction. */
int ptrec_tc_scale_roll(float* slots, int32_t n_slots, float* call_scales, int32_t* err, void* stream);
int ptrec_tc_split2h_prescaled(const float* src, int64_t ld, int64_t R, int64_t C, const float* relu_ref, int64_t ld_ref,
                               void* planes, int64_t pl_ld, void* planes_t, int64_t pt_ld, float* colsum,
                               const float* scale_in, float* max_out, void* workspace, size_t workspace_bytes,
                               void* stream);
size_t ptrec_tc_gemm_fused_workspace_bytes(int64_t M, int64_t N);
int ptrec_tc_gemm_split2h_fused(const void* a_planes, const float* scale_a, int64_t M, int64_t lda, const void* b_planes,
                                const float* scale_b, int64_t N, int64_t ldb, int64_t K, const float* bias, int32_t relu,
                                float* out, int64_t ldo, void* out_planes, int64_t out_planes_ld, const float* out_scale,
                                const uint32_t* mask_in, uint32_t* mask_out, int64_t mask_ld, float* colsum,
                                float* absmax_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K9 the loss closing train_step (torchrec/model/IModel.py:121 with BCEWithLogitsLoss(reduction='mean')): forward and
 * gradient in one launch.  loss[0] = mean_i (max(x,0) - x t + log1p(exp(-|x|))); grad[i] = (sigmoid(x_i) - t_i) / n
 * (grad may be NULL).  Deterministic (fixed summation order).  workspace: ptrec_bce_logits_workspace_bytes(), zero
 * before the first call and owned by this entry point afterwards.
 */
size_t ptrec_bce_logits_workspace_bytes(void);
int ptrec_bce_logits_mean(const float* logits, const float* target, int64_t n, float* loss, float* grad,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * C1 row-wise sharding: pack / unpack either side of the all-to-all (NCCL, issued by the host through
 * torch.distributed).  owner(id) = id mod G, local_row = id div G.  No reference counterpart: the
 * reference is single-device (torchrec/task/Task.py:187-190).
 *   ids      [F, B] int64 one-hot lookups of this rank's batch (id < 0 = no lookup)
 *   send_ids [G, F, C] int64: local rows for each (owner, field), in batch order, unused slots -1
 *   ret_pos  [F, B] int32: (owner*F + f)*C + slot of each lookup, -1 if its list overflowed
 *   overflow int32 word or NULL: max list length seen when some list exceeded C (lookups dropped)
 * workspace: ptrec_a2a_pack_workspace_bytes(B, F, G).
 */
size_t ptrec_a2a_pack_workspace_bytes(int64_t B, int32_t F, int32_t G);
int ptrec_a2a_pack_by_owner(const int64_t* ids, int64_t B, int32_t F, int32_t G, int32_t C,
                            int64_t* send_ids, int32_t* ret_pos, int32_t* overflow, void* workspace,
                            size_t workspace_bytes, void* stream);
/* dst[ret_pos[f,b] * dst_row_stride + 0..D) = scale * src[b, f, :]  (gradient rows into the all-to-all send
 * layout; src [B, F*D] with row stride src_row_stride; dst rows may be wider than D so that several embedding
 * widths of the same fields travel in one collective) */
int ptrec_a2a_scatter_rows(const float* src, int64_t src_row_stride, const int32_t* ret_pos, int64_t B,
                           int32_t F, int32_t D, float scale, float* dst, int64_t dst_row_stride, void* stream);


/* Peer-memory dispatch (no all-to-all): the same lists and gradient rows are STORED straight into the owners'
 * receive buffers over NVLink, in the layout the owner-side ptrec_sort_dedup / ptrec_embedding_bwd_fused read:
 *   peer_ids [G] device array; peer_ids[o] = rank o's id buffer [F, G_src, C] int64; this rank writes its lists
 *            peer_ids[o][(f*G + my_rank)*C + slot] IN FULL every call (lookups, then -1 up to C), so the owner
 *            never resets anything
 *   peer_dst [G] device array; peer_dst[o] = rank o's gradient buffer [G_src*F*C, dst_row_stride] float32; this rank
 *            writes row my_rank*F*C + (ret_pos mod F*C), columns [dst_col, dst_col + D)
 * ret_pos keeps the meaning above; slot_b (optional, [G*F*C] int32, local) receives its inverse: the sample b behind
 * slot (o*F + f)*C + slot, -1 behind the last lookup of a list — what ptrec_a2a_scatter_rows_peer_ordered walks.  The
 * caller puts a barrier (ptrec_peer_barrier) between these stores and the owners' reads, and between the owners' reads
 * and the next step's stores. */
int ptrec_a2a_pack_by_owner_peer(const int64_t* ids, int64_t B, int32_t F, int32_t G, int32_t C, int32_t my_rank,
                                 int64_t* const* peer_ids, int32_t* ret_pos, int32_t* slot_b, int32_t* overflow,
                                 void* workspace, size_t workspace_bytes, void* stream);
int ptrec_a2a_scatter_rows_peer(const float* src, int64_t src_row_stride, const int32_t* ret_pos, int64_t B,
                                int32_t F, int32_t D, float scale, float* const* peer_dst, int64_t dst_row_stride,
                                int64_t dst_col, int32_t C, int32_t G, int32_t my_rank, void* stream);

/* Push mode of the forward exchange (the fused form of owner-side gather -> all-to-all(rows) -> gather by slot; no
 * reference counterpart, torchrec/task/Task.py:187-190):
 *   ptrec_a2a_pack_by_owner_push  as ..._peer, and also stores each lookup's sample index b into the owner's
 *     peer_b[o][(f*G + my_rank)*C + slot] (int32) — where the owner must deliver the row; output rows that no owner
 *     will write (negative id, overflowed list) are zeroed in local_out[k][b*out_row_strides[k] + f*dims[k] ...]
 *     (local_out / out_row_strides / dims: HOST arrays of n_widths <= 4 entries).
 *   ptrec_gather_push  run by the OWNER after a barrier: for every received slot (f, src, c) with id >= 0, read row
 *     `id` of its shard of table f (every width k: table_ptrs[k] is a DEVICE array [F] of shard bases) and store it
 *     into the requester's output  peer_out[k][src] + recv_b*out_row_strides[k] + f*dims[k]  over NVLink
 *     (peer_out[k]: DEVICE array [G] of the ranks' output buffers of width k).  table_ptrs / peer_out / row_strides /
 *     out_row_strides / dims: HOST arrays of n_widths entries.  shard_rows [F] device int64 (ids beyond -> err_flag).
 * The caller puts a barrier (ptrec_peer_barrier) between pack and gather_push, and between gather_push and the
 * requester's reads of its output. */
int ptrec_a2a_pack_by_owner_push(const int64_t* ids, int64_t B, int32_t F, int32_t G, int32_t C, int32_t my_rank,
                                 int64_t* const* peer_ids, int32_t* const* peer_b, float* const* local_out,
                                 const int64_t* out_row_strides, const int32_t* dims, int32_t n_widths,
                                 int32_t* ret_pos, int32_t* slot_b, int32_t* overflow, void* workspace,
                                 size_t workspace_bytes, void* stream);
int ptrec_gather_push(const void* const* const* table_ptrs, float* const* const* peer_out, const int64_t* row_strides,
                      const int64_t* out_row_strides, const int32_t* dims, int32_t n_widths, const int64_t* recv_ids,
                      const int32_t* recv_b, const int64_t* shard_rows, int32_t F, int32_t G, int32_t C,
                      int32_t* err_flag, void* stream);

/* Same, every embedding width of the fields in one launch (srcs / strides / dims / dst_cols: HOST arrays of n_widths
 * entries, n_widths <= 4): the stores of one slot to its owner are adjacent. */
int ptrec_a2a_scatter_rows_peer_multi(const float* const* srcs, const int64_t* src_row_strides, const int32_t* dims,
                                      const int64_t* dst_cols, int32_t n_widths, const int32_t* ret_pos, int64_t B,
                                      int32_t F, float scale, float* const* peer_dst, int64_t dst_row_stride,
                                      int32_t C, int32_t G, int32_t my_rank, void* stream);
/* The same dispatch in DESTINATION order: one sub-warp per slot of this rank's lists (slot_b from the pack call names
 * the sample), whole slots stored (dst_row_stride floats, pad lanes zero), so that the NVLink stores to one owner are
 * long contiguous runs.  dst_cols[k] must be 4 * (float4 chunks of the widths before k). */
int ptrec_a2a_scatter_rows_peer_ordered(const float* const* srcs, const int64_t* src_row_strides, const int32_t* dims,
                                        const int64_t* dst_cols, int32_t n_widths, const int32_t* slot_b, int32_t F,
                                        float scale, float* const* peer_dst, int64_t dst_row_stride, int32_t C,
                                        int32_t G, int32_t my_rank, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PTREC_B200_H */
