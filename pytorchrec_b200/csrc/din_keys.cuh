// K4 with the key gather fused in: the keys of DIN's attention unit are rows of the item / category tables, so the
// tensor-core kernels (din_attn_tc.cu) can read them by id instead of from a gathered [B, L, DQ] tensor:
//   key (b, l) = [ T0[ids0[b * ids_sb + ids_off + l]][0 : DQ/2] | T1[ids1[b * ids_sb + ids_off + l]][0 : DQ/2] ]
// (fp32 tables, possibly interleaved with their optimizer state: `stride` is the row pitch in floats).  An id outside
// [0, rows) raises *err and reads as a zero row, like the gather kernels (gather_pool.cu).
#pragma once
#include <cstdint>

namespace ptrec {

struct DinKeyIds {
  const float* base[2];
  int64_t stride[2];
  int64_t rows[2];
  const int64_t* ids[2];   // ids[0] == nullptr: keys come from the dense tensor
  int64_t ids_sb, ids_off;
  int32_t* err;
};

}  // namespace ptrec
