"""Op-level CPU restatements (test oracle; see oracle/__init__.py)."""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

MASKED = 0xFFFFFFFF


# ------------------------------------------------------------------------------------------ masks
def valid_mask(ids: Tensor, mask: str, lens: Optional[Tensor] = None) -> Tensor:
    """Boolean validity of the slots of a padded ``[B, L]`` id matrix.
    'pad'            ids > 0 ........................ torchrec/model/SVDPP.py:49 (``implicit_i_ids.gt(0)``)
    'pad_keep_first' ids > 0 with column 0 forced ... torchrec/model/utils.py:5-10 (``get_valid_his_index``)
    'lens'           l < len[b] ..................... length columns of HistoryDataReader.py:55-69 / SASRec.py:109-110
    """
    B, L = ids.shape
    if mask == "none":
        return torch.ones(B, L, dtype=torch.bool)
    if mask == "pad":
        return ids != 0
    if mask == "pad_keep_first":
        m = ids != 0
        m[:, 0] = True
        return m
    if mask == "lens":
        return torch.arange(L).unsqueeze(0) < lens.reshape(B, 1)
    raise ValueError(mask)


def index_prep_ref(ids: Tensor, mask: str, lens: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Padded ids -> (valid ids in row-major order, offsets[B+1]); integer cumsum of the mask
    (the count the reference takes with ``valid.sum(dim=-1)``, SVDPP.py:53)."""
    m = valid_mask(ids, mask, lens)
    counts = m.sum(dim=1)
    offsets = torch.zeros(ids.shape[0] + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts, 0)
    return ids[m], offsets


# ------------------------------------------------------------------------------------------ gather + pool
def pooled_lookup_ref(weight: Tensor, ids: Tensor, pooling: str = "sum", mask: str = "none",
                      lens: Optional[Tensor] = None) -> Tensor:
    """``nn.Embedding`` gather (FunkSVD.py:47-48) and, for ``[B, L]`` ids, the reference's masked
    pooling chain: materialise ``[B, L, D]``, multiply by the float mask, sum over L, divide
    (SVDPP.py:49-55 for sqrtn; SASRec.py:109-110 for mean).  Count is clamped to >= 1 (documented
    deviation: SVDPP divides by sqrt(0) on an all-PAD bag)."""
    vec = F.embedding(ids, weight)
    if ids.dim() == 1:
        return vec
    m = valid_mask(ids, mask, lens).to(vec.dtype)
    summed = (vec * m.unsqueeze(-1)).sum(dim=1)
    cnt = m.sum(dim=-1).clamp(min=1.0)
    if pooling == "sum":
        return summed
    if pooling == "mean":
        return summed / cnt.unsqueeze(-1)
    if pooling == "sqrtn":
        return summed / cnt.sqrt().unsqueeze(-1)
    raise ValueError(pooling)


def multi_table_lookup_ref(weights: Sequence[Tensor], table_of: Sequence[int], id_list: Sequence[Tensor],
                           pooling: Sequence[str], mask: Sequence[str],
                           lens: Sequence[Optional[Tensor]]) -> Tensor:
    """Per-feature lookups stacked to ``[B, F, D]`` — the 26 separate ``nn.Embedding`` calls + ``stack``
    that the fused kernel replaces (pattern of SVDPP.py:57-61 repeated per column)."""
    outs = [pooled_lookup_ref(weights[table_of[f]], id_list[f], pooling[f], mask[f], lens[f])
            for f in range(len(id_list))]
    return torch.stack(outs, dim=1)


# ------------------------------------------------------------------------------------------ sort / dedup
def sort_dedup_ref(ids_per_table: Sequence[Tensor], valid_per_table: Sequence[Tensor], rows: Sequence[int]):
    """Integer artefacts of the backward, per table, with torch CPU as ground truth:
    ``torch.sort(stable=True)`` of the keys (masked / out-of-range slots keyed 0xFFFFFFFF sort last) and the
    run structure ``torch.unique(sorted=True, return_counts=True)``.  Returns lists (sorted_keys, perm,
    unique_keys, counts) of int64 tensors; ``perm`` indexes slots *within the table*."""
    out = []
    for ids, valid, r in zip(ids_per_table, valid_per_table, rows):
        key = ids.reshape(-1).clone().to(torch.int64)
        ok = valid.reshape(-1) & (key >= 0) & (key < r)
        key[~ok] = MASKED
        skey, perm = torch.sort(key, stable=True)
        uniq, counts = torch.unique(key, sorted=True, return_counts=True)
        out.append((skey, perm, uniq, counts))
    return out


# ------------------------------------------------------------------------------------------ FM second order
def fm2_ref(v: Tensor) -> Tensor:
    """0.5 * sum_k((sum_f v)^2 - sum_f v^2) for ``v [B, F, D]``.  With F = 2 this is the reference's
    ``(u_vectors * i_vectors).sum(dim=-1)`` (FunkSVD.py:51)."""
    s = v.sum(dim=1)
    return 0.5 * (s * s - (v * v).sum(dim=1)).sum(dim=-1)


# ------------------------------------------------------------------------------------------ sparse row updates
def dense_embedding_grad_ref(ids: Tensor, grad_rows: Tensor, rows: int) -> Tensor:
    """``aten::embedding_dense_backward``: zeros [rows, D] then index_add_ (implicit at IModel.py:123)."""
    g = torch.zeros(rows, grad_rows.shape[-1], dtype=grad_rows.dtype)
    g.index_add_(0, ids.reshape(-1), grad_rows.reshape(-1, grad_rows.shape[-1]))
    return g


def rowwise_adagrad_ref(weight: Tensor, state: Tensor, grad: Tensor, lr: float, eps: float) -> None:
    """Row-wise Adagrad on a dense gradient (rows with zero gradient are unchanged):
    state[r] += mean_k g[r,k]^2 ; w[r] -= lr * g[r] / (sqrt(state[r]) + eps).  No torch / reference
    counterpart exists (unpinned); this is the published FBGEMM/TorchRec definition."""
    touched = (grad != 0).any(dim=1)
    state[touched] += (grad[touched] ** 2).mean(dim=1)
    weight[touched] -= lr * grad[touched] / (state[touched].sqrt() + eps).unsqueeze(1)


def crossed_ids_ref(id_list: Sequence[np.ndarray], category_nums: Sequence[int]) -> np.ndarray:
    """Mixed-radix cross, CrossedColumn.py:14-27."""
    out = np.zeros_like(id_list[0], dtype=np.int64)
    for ids, n in zip(id_list, category_nums):
        out = out * n + ids.astype(np.int64)
    return out


# ---------------------------------------------------------------------------------------------------------------
# K6 fp16 x 2 operand format (no reference counterpart: the reference multiplies in fp32, torchrec/model/layer/
# Dense.py:9-17; this restates include/ptrec_b200.h's definition so the device planes can be checked bit for bit)
# ---------------------------------------------------------------------------------------------------------------
def h2_scale_ref(absmax: float) -> float:
    """Power of two that puts ``absmax`` in [2^13, 2^14); 1 for 0 / inf / nan (split3.cuh::h2_scale)."""
    import math
    if not (absmax > 0.0) or math.isinf(absmax):
        return 1.0
    e = math.frexp(absmax)[1] - 1          # floor(log2(absmax))
    e = max(e, -127)                        # fp32 subnormals report the exponent field 0
    return 2.0 ** max(-126, min(13 - e, 126))


def split2h_ref(x: torch.Tensor, mask_ref: torch.Tensor = None):
    """(h0, h1, scale): x * scale = h0 + h1 / 2048 with h0 = fp16(x * scale), h1 = fp16((x * scale - h0) * 2048);
    ``scale`` from the absolute maximum of x BEFORE the ReLU mask ``mask_ref > 0`` is applied."""
    x = x.float()
    scale = h2_scale_ref(float(x.abs().max())) if x.numel() else 1.0
    if mask_ref is not None:
        x = x * (mask_ref > 0)
    xs = x * scale
    h0 = xs.half()
    h1 = ((xs - h0.float()) * 2048.0).half()
    return h0, h1, scale


H2_HEADROOM = 8  # tc_linear.cu::kH2Headroom


def h2_carried_scale_ref(absmax: float) -> float:
    """Scale a slot takes at a roll: the maximum lands in [2^5, 2^6), eight binades below ``h2_scale_ref``'s target
    (tc_linear.cu::scale_roll_kernel)."""
    import math
    e = int(math.log2(h2_scale_ref(absmax))) - H2_HEADROOM
    return 2.0 ** max(-126, e)


def split2h_prescaled_ref(x: torch.Tensor, scale: float, mask_bits: torch.Tensor = None):
    """(h0, h1) of x * scale with the scale GIVEN; ``mask_bits`` (bool, same shape) zeroes elements first."""
    x = x.float()
    if mask_bits is not None:
        x = x * mask_bits
    xs = x * scale
    h0 = xs.half()
    h1 = ((xs - h0.float()) * 2048.0).half()
    return h0, h1


def unpack_mask_ref(words: torch.Tensor, n: int) -> torch.Tensor:
    """int32 [M, W] bit mask (bit c % 32 of word c // 32) -> bool [M, n]."""
    w = words.to(torch.int64) & 0xFFFFFFFF
    cols = torch.arange(n)
    return ((w[:, cols // 32] >> (cols % 32)) & 1).bool()


def gemm_split2h_ref(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """A B^T through the fp16 x 2 format with fp32 accumulation (CPU emulation of the K6 arithmetic: plane products
    are exact in fp32, the accumulation order differs from the tensor core's)."""
    a0, a1, sa = split2h_ref(a)
    b0, b1, sb = split2h_ref(b)
    main = a0.float() @ b0.float().t()
    corr = a0.float() @ b1.float().t() + a1.float() @ b0.float().t()
    return ((main + corr * (1.0 / 2048.0)) * (1.0 / sa)) * (1.0 / sb)
