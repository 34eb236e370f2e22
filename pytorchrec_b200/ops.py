"""Thin torch-facing wrappers over the C ABI (``include/ptrec_b200.h``).

Everything here runs on CUDA tensors and enqueues on the current torch stream; nothing
synchronises with the host.  There is no CPU path: the CPU implementation of this hot path *is*
the reference (torchrec/model/IModel.py:116-125 + nn.Embedding + torch.optim).
"""
import ctypes
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import FeatureDesc, OptimArgs


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "pytorchrec_b200 kernels run on CUDA tensors only; there is no CPU fallback "
                f"(got a tensor on {t.device})")


# ----------------------------------------------------------------------------------------------
# feature layout: the ptrec_feature_desc array, host and device copies
# ----------------------------------------------------------------------------------------------
class FeatureLayout:
    """Descriptor array for the sparse features of one embedding group (all tables share ``dim``).

    ``specs``: one dict per feature, ordered by table, with keys ``table`` (int), ``bag_len`` (int),
    ``pooling`` ('sum'|'mean'|'sqrtn'), ``mask`` ('none'|'pad'|'pad_keep_first'|'lens'),
    ``lens_col`` (int, -1 when unused).  Output columns are laid out feature after feature.
    """

    def __init__(self, specs: Sequence[Dict], dim: int, n_tables: int):
        self.dim = int(dim)
        self.n_tables = int(n_tables)
        self.n_features = len(specs)
        if self.n_features < 1 or self.n_features > 256 or self.n_tables > 128:
            raise ValueError("1..256 features and 1..128 tables per embedding group")
        arr = (FeatureDesc * self.n_features)()
        id_base = 0
        prev_table = -1
        for f, s in enumerate(specs):
            if s["table"] < prev_table:
                raise ValueError("features must be ordered by table index")
            prev_table = s["table"]
            arr[f].table = int(s["table"])
            arr[f].bag_len = int(s.get("bag_len", 1))
            arr[f].pooling = _lib.POOLING_NAMES[s.get("pooling", "sum")]
            arr[f].mask_mode = _lib.MASK_NAMES[s.get("mask", "none")]
            arr[f].lens_col = int(s.get("lens_col", -1))
            arr[f].flags = _lib.FEAT_NEG_IS_PAD if s.get("neg_is_pad") else 0
            arr[f].id_base = id_base
            arr[f].out_col = int(s.get("out_col", f * self.dim))
            id_base += arr[f].bag_len
        self.host = arr
        self.total_bag_len = id_base
        self.out_width = int(max(arr[f].out_col for f in range(self.n_features))) + self.dim
        self._dev: Dict[torch.device, torch.Tensor] = {}

    def device_array(self, device: torch.device) -> torch.Tensor:
        t = self._dev.get(device)
        if t is None:
            raw = bytes(self.host)
            t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
            self._dev[device] = t
        return t

    def slots(self, batch: int) -> int:
        return self.total_bag_len * batch


_workspaces: Dict[tuple, torch.Tensor] = {}
_retired_workspaces: List[torch.Tensor] = []  # superseded buffers: never handed back to the allocator (see below)


def _workspace(name: str, nbytes: int, device: torch.device) -> torch.Tensor:
    """Byte scratch per (name, device, stream).  Caller-owned memory, per the ABI contract.

    * Keyed by the CURRENT STREAM as well: the early sort on the side stream and a main-stream sort of another
      embedding group never share scratch, so no cross-stream ordering is needed between them.
    * Grow-only, and a superseded buffer is kept alive for the life of the process: its address may be baked into a
      captured whole-step CUDA graph, whose replays would otherwise write into a block the caching allocator has
      handed to someone else (a larger evaluate() batch, a second model in the process).  Growth is geometric so the
      retired bytes stay below the live ones."""
    key = (name, device, torch.cuda.current_stream(device).cuda_stream)
    t = _workspaces.get(key)
    if t is None or t.numel() < nbytes:
        if t is not None:
            _retired_workspaces.append(t)
            nbytes = max(nbytes, 2 * t.numel())
        t = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _workspaces[key] = t
    return t


# ----------------------------------------------------------------------------------------------
# index prep
# ----------------------------------------------------------------------------------------------
def index_prep(ids_padded: torch.Tensor, lens: Optional[torch.Tensor], mask: str):
    """Padded ``[B, L]`` int64 ids -> (compact ids, offsets[B+1]).  Bit-exact integer work."""
    lib = _lib.load()
    _require_cuda(ids_padded, lens)
    assert ids_padded.dtype == torch.int64 and ids_padded.dim() == 2 and ids_padded.is_contiguous()
    B, L = ids_padded.shape
    dev = ids_padded.device
    if lens is not None:
        lens = lens.to(torch.int32).contiguous()
    out_ids = torch.empty(B * L, dtype=torch.int64, device=dev)
    offsets = torch.empty(B + 1, dtype=torch.int64, device=dev)
    nbytes = lib.ptrec_index_prep_workspace_bytes(B)
    ws = _workspace("index_prep", nbytes, dev)
    _lib.check(lib.ptrec_index_prep(_ptr(ids_padded), _ptr(lens), B, L, _lib.MASK_NAMES[mask],
                                    _ptr(out_ids), _ptr(offsets), _ptr(ws), ws.numel(), _stream(dev)),
               "ptrec_index_prep")
    return out_ids, offsets


# ----------------------------------------------------------------------------------------------
# K1 / K2 raw calls
# ----------------------------------------------------------------------------------------------
class TableSet:
    """Device pointer arrays for a list of table tensors (+ optimizer state), rebuilt when a
    tensor moves (``.to(device)``, ``load_state_dict`` keeps pointers)."""

    def __init__(self):
        self._key = None
        self.ptrs = None
        self.rows = None
        self.max_rows = 0
        self.row_stride = 0
        self.dtype = _lib.F32   # PTREC_F32 / PTREC_BF16: element type of every table of the group

    def refresh(self, weights: Sequence[torch.Tensor]):
        key = tuple((w.data_ptr(), w.stride(0), w.dtype) for w in weights)
        if key != self._key:
            dev = weights[0].device
            # the group's row stride comes from any table with more than one row (a one-row table has no meaningful
            # stride: it may still be a plain [1, D] tensor while the others are interleaved with their state)
            multi = [w for w in weights if w.shape[0] > 1]
            stride = multi[0].stride(0) if multi else weights[0].shape[1]
            wdt = weights[0].dtype
            for w in weights:
                _require_cuda(w)
                if w.dtype not in (torch.float32, torch.bfloat16) or w.dtype != wdt or w.dim() != 2 or w.stride(1) != 1 \
                        or w.device != dev:
                    raise RuntimeError("embedding tables of one group must be fp32 (or all bf16) [rows, D] CUDA tensors "
                                       "with unit inner stride on one device")
                if w.shape[0] > 1 and w.stride(0) != stride:
                    raise RuntimeError("all tables of one embedding group must share the same row stride "
                                       "(all plain, or all interleaved with their optimizer state)")
                align = (16 if w.shape[1] >= 4 else 4 * w.shape[1]) // (2 if wdt == torch.bfloat16 else 1)
                if w.data_ptr() % align != 0 or (w.shape[0] > 1 and (w.stride(0) * w.element_size()) % align != 0):
                    raise RuntimeError("embedding table base pointer / row pitch is misaligned")
            self.dtype = _lib.BF16 if wdt == torch.bfloat16 else _lib.F32
            self.row_stride = int(stride)
            self.ptrs = torch.tensor([w.data_ptr() for w in weights], dtype=torch.int64).to(dev)
            self.rows = torch.tensor([w.shape[0] for w in weights], dtype=torch.int64).to(dev)
            self.max_rows = max(w.shape[0] for w in weights)
            self._key = key
        return self


def make_ptr_array(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    dev = tensors[0].device
    return torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64).to(dev)


def gather_pool_fwd(tables: TableSet, layout: FeatureLayout, ids: torch.Tensor,
                    lens: Optional[torch.Tensor], batch: int, out: Optional[torch.Tensor] = None,
                    want_scale: bool = False, err_flag: Optional[torch.Tensor] = None,
                    out_row_stride: Optional[int] = None):
    lib = _lib.load()
    _require_cuda(ids, lens, out)
    dev = ids.device
    assert ids.dtype == torch.int64 and ids.is_contiguous() and ids.numel() == layout.slots(batch)
    if out is None:
        out = torch.empty(batch, layout.out_width, dtype=torch.float32, device=dev)
    bag_scale = torch.empty(layout.n_features, batch, dtype=torch.float32, device=dev) if want_scale else None
    _lib.check(_gather_call(lib, tables, layout, ids, lens, batch, out, bag_scale, err_flag, dev,
                            out.stride(0) if out_row_stride is None else out_row_stride),
               "ptrec_embedding_gather_pool_fwd")
    return out, bag_scale


def _gather_call(lib, tables, layout, ids, lens, batch, out, bag_scale, err_flag, dev, row_stride):
    feats_dev = layout.device_array(dev)
    return lib.ptrec_embedding_gather_pool_fwd(
        _ptr(tables.ptrs), _ptr(tables.rows), layout.n_tables, layout.dim, tables.row_stride or layout.dim, tables.dtype,
        ctypes.cast(ctypes.c_void_p(feats_dev.data_ptr()), ctypes.POINTER(FeatureDesc)), layout.host,
        layout.n_features, _ptr(ids), _ptr(lens), batch, _ptr(out), row_stride, _ptr(bag_scale),
        _ptr(err_flag), _stream(dev))


def gather_fwd_sharded(shard_ptrs: torch.Tensor, table_rows: torch.Tensor, G: int, row_stride: int,
                       layout: FeatureLayout, ids: torch.Tensor, batch: int, out: Optional[torch.Tensor] = None,
                       err_flag: Optional[torch.Tensor] = None):
    """One-hot lookups over row-wise shards reached through (peer) pointers: ``shard_ptrs`` int64 [T, G] device
    tensor of shard base addresses as seen from this device, ``table_rows`` int64 [T] GLOBAL row counts."""
    lib = _lib.load()
    _require_cuda(ids, out, shard_ptrs, table_rows)
    dev = ids.device
    assert ids.dtype == torch.int64 and ids.is_contiguous() and ids.numel() == layout.slots(batch)
    assert shard_ptrs.dtype == torch.int64 and shard_ptrs.numel() == layout.n_tables * G
    if out is None:
        out = torch.empty(batch, layout.out_width, dtype=torch.float32, device=dev)
    feats_dev = layout.device_array(dev)
    _lib.check(lib.ptrec_embedding_gather_pool_fwd_sharded(
        _ptr(shard_ptrs), _ptr(table_rows), layout.n_tables, G, layout.dim, row_stride, _lib.F32,
        ctypes.cast(ctypes.c_void_p(feats_dev.data_ptr()), ctypes.POINTER(FeatureDesc)), layout.host,
        layout.n_features, _ptr(ids), batch, _ptr(out), out.stride(0), _ptr(err_flag), _stream(dev)),
        "ptrec_embedding_gather_pool_fwd_sharded")
    return out


class SortResult:
    __slots__ = ("sorted_keys", "perm", "seg_start", "seg_meta", "n_seg", "N")

    @property
    def seg_table(self):
        """table index of each segment (column 2 of the 16-byte ptrec_segment_meta records)"""
        return self.seg_meta[:, 2]


def sort_dedup(tables: TableSet, layout: FeatureLayout, ids: torch.Tensor,
               lens: Optional[torch.Tensor], batch: int) -> SortResult:
    lib = _lib.load()
    _require_cuda(ids, lens)
    dev = ids.device
    N = layout.slots(batch)
    r = SortResult()
    r.N = N
    # int32 view of uint32 keys (torch has no uint32 arithmetic; tests reinterpret)
    r.sorted_keys = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    r.perm = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    r.seg_start = torch.empty(N + 1, dtype=torch.int32, device=dev)
    r.seg_meta = torch.empty(max(N, 1), 4, dtype=torch.int32, device=dev)  # ptrec_segment_meta[N]
    r.n_seg = torch.empty(1, dtype=torch.int32, device=dev)
    nbytes = lib.ptrec_sort_dedup_workspace_bytes(N, layout.n_tables)
    ws = _workspace("sort_dedup", nbytes, dev)
    feats_dev = layout.device_array(dev)
    _lib.check(lib.ptrec_sort_dedup(
        ctypes.cast(ctypes.c_void_p(feats_dev.data_ptr()), ctypes.POINTER(FeatureDesc)), layout.host,
        layout.n_features, layout.n_tables, _ptr(tables.rows), tables.max_rows, _ptr(ids), _ptr(lens),
        batch, _ptr(r.sorted_keys), _ptr(r.perm), _ptr(r.seg_start), _ptr(r.seg_meta), _ptr(r.n_seg),
        _ptr(ws), ws.numel(), _stream(dev)), "ptrec_sort_dedup")
    return r


_BWD_FN = {
    _lib.OPT_SGD: "ptrec_embedding_bwd_fused_sgd",
    _lib.OPT_ADAGRAD: "ptrec_embedding_bwd_fused_adagrad",
    _lib.OPT_ROWWISE_ADAGRAD: "ptrec_embedding_bwd_fused_rowwise_adagrad",
    _lib.OPT_LAZY_ADAM: "ptrec_embedding_bwd_fused_lazy_adam",
}


def bwd_fused(tables: TableSet, state1_ptrs: Optional[torch.Tensor], state2_ptrs: Optional[torch.Tensor],
              layout: FeatureLayout, batch: int, srt: SortResult, grad_out: torch.Tensor,
              bag_scale: Optional[torch.Tensor], opt: OptimArgs, grad_row_stride: Optional[int] = None) -> None:
    lib = _lib.load()
    _require_cuda(grad_out, bag_scale)
    dev = grad_out.device
    assert grad_out.dtype == torch.float32 and grad_out.stride(-1) == 1
    if grad_row_stride is None:
        assert grad_out.dim() == 2
        grad_row_stride = grad_out.stride(0)
    nbytes = lib.ptrec_embedding_bwd_workspace_bytes(srt.N, layout.dim)
    ws = _workspace("bwd_fused", nbytes, dev)
    feats_dev = layout.device_array(dev)
    fn = getattr(lib, _BWD_FN[opt.kind])
    _lib.check(fn(_ptr(tables.ptrs), _ptr(state1_ptrs), _ptr(state2_ptrs), layout.n_tables, layout.dim,
                  tables.row_stride or layout.dim, tables.dtype, ctypes.cast(ctypes.c_void_p(feats_dev.data_ptr()), ctypes.POINTER(FeatureDesc)),
                  layout.host, layout.n_features, batch, _ptr(srt.sorted_keys), _ptr(srt.perm),
                  _ptr(srt.seg_start), _ptr(srt.seg_meta), _ptr(srt.n_seg), _ptr(grad_out),
                  grad_row_stride, _ptr(bag_scale), ctypes.byref(opt), _ptr(ws), ws.numel(),
                  _stream(dev)), _BWD_FN[opt.kind])


def segment_sum(layout: FeatureLayout, batch: int, srt: SortResult, grad_out: torch.Tensor,
                bag_scale: Optional[torch.Tensor]) -> torch.Tensor:
    """Per-segment gradient sums ``[N, D]`` (first ``n_seg`` rows meaningful)."""
    lib = _lib.load()
    _require_cuda(grad_out, bag_scale)
    dev = grad_out.device
    row_grad = torch.empty(max(srt.N, 1), layout.dim, dtype=torch.float32, device=dev)
    feats_dev = layout.device_array(dev)
    _lib.check(lib.ptrec_embedding_bwd_segment_sum(
        layout.n_tables, layout.dim,
        ctypes.cast(ctypes.c_void_p(feats_dev.data_ptr()), ctypes.POINTER(FeatureDesc)), layout.host,
        layout.n_features, batch, _ptr(srt.sorted_keys), _ptr(srt.perm), _ptr(srt.seg_start),
        _ptr(srt.seg_meta), _ptr(srt.n_seg), _ptr(grad_out), grad_out.stride(0), _ptr(bag_scale),
        _ptr(row_grad), _stream(dev)), "ptrec_embedding_bwd_segment_sum")
    return row_grad


# ----------------------------------------------------------------------------------------------
# K3 FM second-order interaction
# ----------------------------------------------------------------------------------------------
def fm2_fwd(v: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _require_cuda(v)
    assert v.dtype == torch.float32 and v.dim() == 3 and v.stride(2) == 1 and v.stride(1) == v.shape[2]
    B, F, D = v.shape
    y = torch.empty(B, dtype=torch.float32, device=v.device)
    _lib.check(lib.ptrec_fm2_fwd(_ptr(v), v.stride(0), B, F, D, _ptr(y), _stream(v.device)), "ptrec_fm2_fwd")
    return y


def fm2_bwd(v: torch.Tensor, gy: torch.Tensor, grad_in: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    _require_cuda(v, gy, grad_in)
    B, F, D = v.shape
    gy = gy.contiguous()
    gv = torch.empty(B, F, D, dtype=torch.float32, device=v.device)
    gi_stride = 0
    if grad_in is not None:
        assert grad_in.dtype == torch.float32 and grad_in.stride(-1) == 1
        gi_stride = grad_in.stride(0)
    _lib.check(lib.ptrec_fm2_bwd(_ptr(v), v.stride(0), _ptr(gy), _ptr(grad_in), gi_stride, B, F, D, _ptr(gv),
                                 gv.stride(0), _stream(v.device)), "ptrec_fm2_bwd")
    return gv


class _FM2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v):
        ctx.save_for_backward(v)
        return fm2_fwd(v)

    @staticmethod
    def backward(ctx, gy):
        (v,) = ctx.saved_tensors
        return fm2_bwd(v, gy)


def fm2(v: torch.Tensor) -> torch.Tensor:
    """FM second-order term of ``v [B, F, D]`` -> ``[B]`` (fused CUDA forward and backward)."""
    if not (v.stride(2) == 1 and v.stride(1) == v.shape[2]):
        v = v.contiguous()
    return _FM2.apply(v)


# ----------------------------------------------------------------------------------------------
# C1 all-to-all pack / unpack
# ----------------------------------------------------------------------------------------------
# ----------------------------------------------------------------------------------------------
# K8 FM head + row dot
# ----------------------------------------------------------------------------------------------
def fm_head_supported(F: int, D: int, nd: int) -> bool:
    return bool(_lib.load().ptrec_fm_head_supported(F, D, nd))


def fm_head_fwd(v: torch.Tensor, w1: Optional[torch.Tensor], x: Optional[torch.Tensor], wd: Optional[torch.Tensor],
                bias: Optional[torch.Tensor], F: int, D: int, want_deep_in: bool, want_planes: bool = False):
    """v [B, F*D] (row stride % 4 == 0), w1 [B, F], x [B, nd], wd [nd], bias [1] -> (logit [B], deep_in or None);
    deep_in is a [B, F*D+nd] view of a buffer whose pitch is rounded up to 4 floats."""
    lib = _lib.load()
    _require_cuda(v, w1, x, wd, bias)
    B = v.shape[0]
    nd = 0 if x is None else x.shape[1]
    dev = v.device
    logit = torch.empty(B, dtype=torch.float32, device=dev)
    deep_buf = None
    if want_deep_in:
        pitch = (F * D + nd + 3) // 4 * 4
        deep_buf = torch.empty(B, pitch, dtype=torch.float32, device=dev)
    planes = torch.empty(3, B, _pad8(F * D + nd), dtype=torch.bfloat16, device=dev) if want_planes else None
    _lib.check(lib.ptrec_fm_head_fwd(_ptr(v), v.stride(0), _ptr(w1), w1.stride(0) if w1 is not None else 0, _ptr(x),
                                     x.stride(0) if x is not None else 0, _ptr(wd), _ptr(bias), B, F, D, nd,
                                     _ptr(logit), _ptr(deep_buf), deep_buf.stride(0) if deep_buf is not None else 0,
                                     _ptr(planes), _pad8(F * D + nd), _stream(dev)), "ptrec_fm_head_fwd")
    deep_in = deep_buf[:, :F * D + nd] if deep_buf is not None else None
    return (logit, deep_in, planes) if want_planes else (logit, deep_in)


def fm_head_bwd(v: torch.Tensor, x: Optional[torch.Tensor], wd: Optional[torch.Tensor], g: torch.Tensor,
                g_deep_in: Optional[torch.Tensor], F: int, D: int, want_w1: bool, want_wd: bool, want_bias: bool):
    """-> (grad_v [B, F*D], grad_w1 [B, F] or None, grad_wd [nd] or None, grad_bias [1] or None)"""
    lib = _lib.load()
    _require_cuda(v, x, wd, g, g_deep_in)
    B = v.shape[0]
    nd = 0 if x is None else x.shape[1]
    dev = v.device
    gv = torch.empty(B, F * D, dtype=torch.float32, device=dev)
    gw1 = torch.empty(B, F, dtype=torch.float32, device=dev) if want_w1 else None
    gwd = torch.empty(nd, dtype=torch.float32, device=dev) if (want_wd and nd) else None
    gb = torch.empty(1, dtype=torch.float32, device=dev) if want_bias else None
    ws = _workspace("fm_head", lib.ptrec_fm_head_bwd_workspace_bytes(nd), dev)
    _lib.check(lib.ptrec_fm_head_bwd(_ptr(v), v.stride(0), _ptr(x), x.stride(0) if x is not None else 0, _ptr(wd),
                                     _ptr(g), _ptr(g_deep_in), g_deep_in.stride(0) if g_deep_in is not None else 0, B,
                                     F, D, nd, _ptr(gv), gv.stride(0), _ptr(gw1), None, 0, _ptr(gwd), _ptr(gb),
                                     _ptr(ws), ws.numel(), _stream(dev)), "ptrec_fm_head_bwd")
    return gv, gw1, gwd, gb


def fm_head_fwd_h2(v: torch.Tensor, w1: Optional[torch.Tensor], x: Optional[torch.Tensor], wd: Optional[torch.Tensor],
                   bias: Optional[torch.Tensor], F: int, D: int, scale: torch.Tensor, max_out: torch.Tensor):
    """``fm_head_fwd`` whose tower input leaves as the two fp16 planes of the K6 fused tower, split with ``scale`` (a
    one-element view of the tower's carried scales); no fp32 tower input is written.
    Returns (logit [B], planes [2, B, pad8(F*D+nd)] fp16)."""
    lib = _lib.load()
    _require_cuda(v, w1, x, wd, bias, scale, max_out)
    B = v.shape[0]
    nd = 0 if x is None else x.shape[1]
    dev = v.device
    logit = torch.empty(B, dtype=torch.float32, device=dev)
    ld = _pad8(F * D + nd)
    planes = torch.empty(2, B, ld, dtype=torch.float16, device=dev)
    _lib.check(lib.ptrec_fm_head_fwd_h2(_ptr(v), v.stride(0), _ptr(w1), w1.stride(0) if w1 is not None else 0, _ptr(x),
                                        x.stride(0) if x is not None else 0, _ptr(wd), _ptr(bias), B, F, D, nd,
                                        _ptr(logit), None, 0, _ptr(planes), ld, _ptr(scale), _ptr(max_out),
                                        _stream(dev)), "ptrec_fm_head_fwd_h2")
    return logit, planes


def rowdot_bwd_h2(h: torch.Tensor, w: torch.Tensor, g: torch.Tensor, scale: torch.Tensor, max_out: torch.Tensor,
                  want_colsum: bool, want_w: bool):
    """Backward of ``rowdot`` behind a fused tower: (planes [2, B, pad8(H)] fp16 of (g (x) w) * (h > 0) split with
    ``scale``, colsum [H] or None, grad_w [H] or None).  No fp32 gradient of h is written."""
    lib = _lib.load()
    _require_cuda(h, w, g, scale, max_out)
    B, H = h.shape
    dev = h.device
    ld = _pad8(H)
    planes = torch.empty(2, B, ld, dtype=torch.float16, device=dev)
    colsum = torch.empty(H, dtype=torch.float32, device=dev) if want_colsum else None
    gw = torch.empty(H, dtype=torch.float32, device=dev) if want_w else None
    ws = _workspace("rowdot_h2", 2 * lib.ptrec_rowdot_bwd_workspace_bytes(H), dev)
    _lib.check(lib.ptrec_rowdot_bwd_h2(_ptr(h), h.stride(0), _ptr(w), _ptr(g), B, H, _ptr(planes), ld, _ptr(scale),
                                       _ptr(max_out), _ptr(colsum), _ptr(gw), _ptr(ws), ws.numel(), _stream(dev)),
               "ptrec_rowdot_bwd_h2")
    return planes, colsum, gw


def rowdot_supported(H: int) -> bool:
    return bool(_lib.load().ptrec_rowdot_supported(H))


def rowdot_fwd(h: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _require_cuda(h, w)
    B, H = h.shape
    y = torch.empty(B, dtype=torch.float32, device=h.device)
    _lib.check(lib.ptrec_rowdot_fwd(_ptr(h), h.stride(0), _ptr(w), B, H, _ptr(y), _stream(h.device)), "ptrec_rowdot_fwd")
    return y


def rowdot_bwd(h: torch.Tensor, w: torch.Tensor, g: torch.Tensor, want_h: bool, want_w: bool):
    lib = _lib.load()
    _require_cuda(h, w, g)
    B, H = h.shape
    dev = h.device
    gh = torch.empty(B, H, dtype=torch.float32, device=dev) if want_h else None
    gw = torch.empty(H, dtype=torch.float32, device=dev) if want_w else None
    ws = _workspace("rowdot", lib.ptrec_rowdot_bwd_workspace_bytes(H), dev)
    _lib.check(lib.ptrec_rowdot_bwd(_ptr(h), h.stride(0), _ptr(w), _ptr(g), B, H, _ptr(gh), gh.stride(0) if want_h else 0,
                                    _ptr(gw), _ptr(ws), ws.numel(), _stream(dev)), "ptrec_rowdot_bwd")
    return gh, gw


def a2a_pack_by_owner(ids: torch.Tensor, F: int, B: int, G: int, C: int, overflow: torch.Tensor):
    """ids [F, B] int64 -> (send_ids [G, F, C] int64 with -1 padding, ret_pos [F, B] int32)."""
    lib = _lib.load()
    _require_cuda(ids, overflow)
    dev = ids.device
    assert ids.dtype == torch.int64 and ids.is_contiguous() and ids.numel() == F * B
    send_ids = torch.empty(G, F, C, dtype=torch.int64, device=dev)
    ret_pos = torch.empty(F, B, dtype=torch.int32, device=dev)
    ws = _workspace("a2a_pack", lib.ptrec_a2a_pack_workspace_bytes(B, F, G), dev)
    _lib.check(lib.ptrec_a2a_pack_by_owner(_ptr(ids), B, F, G, C, _ptr(send_ids), _ptr(ret_pos), _ptr(overflow),
                                           _ptr(ws), ws.numel(), _stream(dev)), "ptrec_a2a_pack_by_owner")
    return send_ids, ret_pos


def a2a_scatter_rows(src: torch.Tensor, ret_pos: torch.Tensor, B: int, F: int, D: int, scale: float,
                     dst: torch.Tensor) -> None:
    """dst is a [n_slots, D] view (possibly a column slice of a wider buffer: its row stride is honoured)."""
    lib = _lib.load()
    _require_cuda(src, ret_pos, dst)
    assert src.dtype == torch.float32 and src.stride(-1) == 1 and dst.dim() == 2 and dst.stride(1) == 1
    _lib.check(lib.ptrec_a2a_scatter_rows(_ptr(src), src.stride(0), _ptr(ret_pos), B, F, D, float(scale), _ptr(dst),
                                          dst.stride(0), _stream(src.device)), "ptrec_a2a_scatter_rows")


def a2a_pack_by_owner_peer(ids: torch.Tensor, F: int, B: int, G: int, C: int, rank: int, peer_ids: torch.Tensor,
                           overflow: torch.Tensor, slot_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ids [F, B] -> ret_pos [F, B]; the lists are stored into the owners' id buffers (``peer_ids`` int64 [G] device
    tensor of their addresses, each [F, G, C] int64).  ``slot_b`` (int32 [G*F*C], optional) receives the inverse of
    ret_pos (the sample behind each slot, -1 for empty slots)."""
    lib = _lib.load()
    _require_cuda(ids, overflow, peer_ids)
    dev = ids.device
    assert ids.dtype == torch.int64 and ids.is_contiguous() and ids.numel() == F * B and peer_ids.numel() == G
    ret_pos = torch.empty(F, B, dtype=torch.int32, device=dev)
    ws = _workspace("a2a_pack", lib.ptrec_a2a_pack_workspace_bytes(B, F, G), dev)
    _lib.check(lib.ptrec_a2a_pack_by_owner_peer(_ptr(ids), B, F, G, C, rank, _ptr(peer_ids), _ptr(ret_pos), _ptr(slot_b),
                                                _ptr(overflow), _ptr(ws), ws.numel(), _stream(dev)),
               "ptrec_a2a_pack_by_owner_peer")
    return ret_pos


def a2a_pack_by_owner_push(ids: torch.Tensor, F: int, B: int, G: int, C: int, rank: int, peer_ids: torch.Tensor,
                           peer_b: torch.Tensor, outs, dims, overflow: torch.Tensor,
                           slot_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Push-mode pack: ids [F, B] -> ret_pos [F, B]; local rows go to the owners' id lists (``peer_ids`` [G] addresses
    of int64 [F, G, C]) and the sample index of each lookup to their ``peer_b`` lists (int32, same indexing); rows
    of the local outputs ``outs[k]`` [B, F*dims[k]] that no owner will write are zeroed."""
    lib = _lib.load()
    _require_cuda(ids, overflow, peer_ids, peer_b, *outs)
    dev = ids.device
    assert ids.dtype == torch.int64 and ids.is_contiguous() and ids.numel() == F * B and peer_ids.numel() == G == peer_b.numel()
    n = len(outs)
    assert n == len(dims) and all(o.dtype == torch.float32 and o.stride(-1) == 1 for o in outs)
    ret_pos = torch.empty(F, B, dtype=torch.int32, device=dev)
    ws = _workspace("a2a_pack", lib.ptrec_a2a_pack_workspace_bytes(B, F, G), dev)
    a_out = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
    a_str = (ctypes.c_int64 * n)(*[o.stride(0) for o in outs])
    a_dim = (ctypes.c_int32 * n)(*dims)
    _lib.check(lib.ptrec_a2a_pack_by_owner_push(_ptr(ids), B, F, G, C, rank, _ptr(peer_ids), _ptr(peer_b), a_out, a_str,
                                                a_dim, n, _ptr(ret_pos), _ptr(slot_b), _ptr(overflow), _ptr(ws), ws.numel(),
                                                _stream(dev)), "ptrec_a2a_pack_by_owner_push")
    return ret_pos


def gather_push(table_ptrs, peer_outs, row_strides, out_strides, dims, recv_ids: torch.Tensor, recv_b: torch.Tensor,
                shard_rows: torch.Tensor, F: int, G: int, C: int, err_flag: Optional[torch.Tensor] = None) -> None:
    """Owner side of the push-mode forward: ``table_ptrs[k]`` int64 [F] device tensor of this rank's shard bases for
    width k, ``peer_outs[k]`` int64 [G] device tensor of the ranks' output buffers of width k."""
    lib = _lib.load()
    _require_cuda(recv_ids, recv_b, shard_rows, err_flag, *table_ptrs, *peer_outs)
    n = len(dims)
    assert recv_ids.dtype == torch.int64 and recv_b.dtype == torch.int32 and recv_ids.numel() == F * G * C == recv_b.numel()
    a_tab = (ctypes.c_void_p * n)(*[t.data_ptr() for t in table_ptrs])
    a_out = (ctypes.c_void_p * n)(*[t.data_ptr() for t in peer_outs])
    a_rs = (ctypes.c_int64 * n)(*row_strides)
    a_os = (ctypes.c_int64 * n)(*out_strides)
    a_dim = (ctypes.c_int32 * n)(*dims)
    _lib.check(lib.ptrec_gather_push(a_tab, a_out, a_rs, a_os, a_dim, n, _ptr(recv_ids), _ptr(recv_b), _ptr(shard_rows),
                                     F, G, C, _ptr(err_flag), _stream(recv_ids.device)), "ptrec_gather_push")


def a2a_scatter_rows_peer(src: torch.Tensor, ret_pos: torch.Tensor, B: int, F: int, D: int, scale: float,
                          peer_dst: torch.Tensor, dst_row_stride: int, dst_col: int, C: int, G: int,
                          rank: int) -> None:
    lib = _lib.load()
    _require_cuda(src, ret_pos, peer_dst)
    assert src.dtype == torch.float32 and src.stride(-1) == 1 and peer_dst.numel() == G
    _lib.check(lib.ptrec_a2a_scatter_rows_peer(_ptr(src), src.stride(0), _ptr(ret_pos), B, F, D, float(scale),
                                               _ptr(peer_dst), dst_row_stride, dst_col, C, G, rank,
                                               _stream(src.device)), "ptrec_a2a_scatter_rows_peer")


def a2a_scatter_rows_peer_multi(srcs, dims, cols, ret_pos: torch.Tensor, B: int, F: int, scale: float,
                                peer_dst: torch.Tensor, dst_row_stride: int, C: int, G: int, rank: int) -> None:
    """Every width in one launch: srcs[k] [B, F*dims[k]] fp32 -> columns [cols[k], cols[k]+dims[k]) of the owners' slots."""
    lib = _lib.load()
    _require_cuda(ret_pos, peer_dst, *srcs)
    n = len(srcs)
    assert n == len(dims) == len(cols) and all(s.dtype == torch.float32 and s.stride(-1) == 1 for s in srcs)
    a_src = (ctypes.c_void_p * n)(*[s.data_ptr() for s in srcs])
    a_str = (ctypes.c_int64 * n)(*[s.stride(0) for s in srcs])
    a_dim = (ctypes.c_int32 * n)(*dims)
    a_col = (ctypes.c_int64 * n)(*cols)
    _lib.check(lib.ptrec_a2a_scatter_rows_peer_multi(a_src, a_str, a_dim, a_col, n, _ptr(ret_pos), B, F, float(scale),
                                                     _ptr(peer_dst), dst_row_stride, C, G, rank,
                                                     _stream(ret_pos.device)), "ptrec_a2a_scatter_rows_peer_multi")


def a2a_scatter_rows_peer_ordered(srcs, dims, cols, slot_b: torch.Tensor, F: int, scale: float, peer_dst: torch.Tensor,
                                  dst_row_stride: int, C: int, G: int, rank: int) -> None:
    """``a2a_scatter_rows_peer_multi`` in destination order (``slot_b`` from the pack call): whole slots, contiguous
    NVLink stores."""
    lib = _lib.load()
    _require_cuda(slot_b, peer_dst, *srcs)
    n = len(srcs)
    assert n == len(dims) == len(cols) and all(s.dtype == torch.float32 and s.stride(-1) == 1 for s in srcs)
    assert slot_b.dtype == torch.int32 and slot_b.numel() == G * F * C
    a_src = (ctypes.c_void_p * n)(*[s.data_ptr() for s in srcs])
    a_str = (ctypes.c_int64 * n)(*[s.stride(0) for s in srcs])
    a_dim = (ctypes.c_int32 * n)(*dims)
    a_col = (ctypes.c_int64 * n)(*cols)
    _lib.check(lib.ptrec_a2a_scatter_rows_peer_ordered(a_src, a_str, a_dim, a_col, n, _ptr(slot_b), F, float(scale),
                                                       _ptr(peer_dst), dst_row_stride, C, G, rank,
                                                       _stream(slot_b.device)), "ptrec_a2a_scatter_rows_peer_ordered")


# ----------------------------------------------------------------------------------------------
# C2 cross-GPU ordering over symmetric memory (csrc/peer_sync.cu)
# ----------------------------------------------------------------------------------------------
class PeerSync:
    """Barrier between the ranks of a process group, made of one tiny kernel of this library per call (release /
    acquire flags in symmetric memory) instead of a NCCL collective.  ``barrier(slot)``: everything every rank
    enqueued before it on its current stream is complete and visible before anything any rank enqueues after it
    starts.  Every rank must issue the same sequence of calls per slot; slots are independent, so two streams may
    each run their own.  Graph-capturable."""
    FENCE, IDS, GRADS, DENSE, ROWS = 0, 1, 2, 3, 4

    def __init__(self, group, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        lib = _lib.load()
        self.group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n = lib.ptrec_peer_sync_slots() * lib.ptrec_peer_sync_max_ranks()
        self.flags = symm.empty(n, dtype=torch.int32, device=device)
        self.flags.zero_()
        self._hdl = symm.rendezvous(self.flags, self.group)
        self.peer_flags = torch.tensor([int(p) for p in self._hdl.buffer_ptrs], dtype=torch.int64).to(device)
        self.epoch = torch.zeros(lib.ptrec_peer_sync_slots(), dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)  # every rank's flags are zero before anyone signals

    def barrier(self, slot: int) -> None:
        dev = self.flags.device
        _lib.check(_lib.load().ptrec_peer_barrier(_ptr(self.peer_flags), _ptr(self.epoch), int(slot), self.world,
                                                  self.rank, _stream(dev)), "ptrec_peer_barrier")


def dense_pack(descs: torch.Tensor, chunk_start: torch.Tensor, n_tensors: int, n_chunks: int, stage: torch.Tensor) -> None:
    """Gradients named by a K7 descriptor table -> the contiguous ``stage`` (n_chunks * chunk floats)."""
    _require_cuda(descs, chunk_start, stage)
    assert stage.dtype == torch.float32 and stage.is_contiguous()
    assert stage.numel() >= n_chunks * _lib.load().ptrec_dense_optim_chunk()
    _lib.check(_lib.load().ptrec_dense_pack(_ptr(descs), _ptr(chunk_start), n_tensors, n_chunks, _ptr(stage),
                                            _stream(stage.device)), "ptrec_dense_pack")


# ----------------------------------------------------------------------------------------------
# K6 fp32-faithful Linear on tcgen05 (bf16 x 3 split operands)
# ----------------------------------------------------------------------------------------------
def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def tc_split3(src: torch.Tensor, relu_ref: Optional[torch.Tensor] = None, want_planes: bool = True,
              want_t: bool = False, want_colsum: bool = False):
    """fp32 [R, C] -> (planes [3, R, pad8(C)] bf16, planes_t [3, C, pad8(R)] bf16, colsum [C]); each None unless asked.
    ``relu_ref`` zeroes src where relu_ref <= 0 first (ReLU backward)."""
    lib = _lib.load()
    _require_cuda(src, relu_ref)
    assert src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1
    R, C = src.shape
    dev = src.device
    if relu_ref is not None:
        assert relu_ref.shape == src.shape and relu_ref.dtype == torch.float32 and relu_ref.stride(1) == 1
    planes = torch.empty(3, R, _pad8(C), dtype=torch.bfloat16, device=dev) if want_planes else None
    planes_t = torch.empty(3, C, _pad8(R), dtype=torch.bfloat16, device=dev) if want_t else None
    colsum = torch.empty(C, dtype=torch.float32, device=dev) if want_colsum else None
    ws = _workspace("tc_split3", lib.ptrec_tc_split3_workspace_bytes(R, C), dev) if want_colsum else None
    _lib.check(lib.ptrec_tc_split3(_ptr(src), src.stride(0), R, C, _ptr(relu_ref),
                                   relu_ref.stride(0) if relu_ref is not None else 0, _ptr(planes), _pad8(C),
                                   _ptr(planes_t), _pad8(R), _ptr(colsum), _ptr(ws), ws.numel() if ws is not None else 0,
                                   _stream(dev)), "ptrec_tc_split3")
    return planes, planes_t, colsum


def tc_gemm_split3(a_planes: torch.Tensor, b_planes: torch.Tensor, K: int, bias: Optional[torch.Tensor] = None,
                   relu: bool = False, splits: int = 1, want_planes: bool = False):
    """out [M, N] fp32 = A[M, K] B[N, K]^T (+ bias) (ReLU) from bf16 planes [3, M, lda], [3, N, ldb].  The result is a
    [:, :N] view of a buffer whose pitch is N rounded up to 4.  ``splits=0`` picks the split-K count for a K-heavy
    product.  ``want_planes``: also return the result as planes [3, M, pad8(N)] (written by the same epilogue)."""
    lib = _lib.load()
    _require_cuda(a_planes, b_planes, bias)
    assert a_planes.dtype == torch.bfloat16 and b_planes.dtype == torch.bfloat16
    assert a_planes.dim() == 3 and b_planes.dim() == 3 and a_planes.is_contiguous() and b_planes.is_contiguous()
    M, lda = a_planes.shape[1], a_planes.shape[2]
    N, ldb = b_planes.shape[1], b_planes.shape[2]
    dev = a_planes.device
    if splits == 0:
        splits = lib.ptrec_tc_gemm_split3_default_splits(M, N, K)
    ldo = (N + 3) // 4 * 4
    out = torch.empty(M, ldo, dtype=torch.float32, device=dev)
    nbytes = lib.ptrec_tc_gemm_split3_workspace_bytes(M, ldo, splits)
    ws = _workspace("tc_gemm_split3", nbytes, dev) if nbytes else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    planes = torch.empty(3, M, _pad8(N), dtype=torch.bfloat16, device=dev) if want_planes else None
    if want_planes and _pad8(N) != ldo:
        planes[:, :, N:].zero_()  # the epilogue writes columns < ldo only
    _lib.check(lib.ptrec_tc_gemm_split3(_ptr(a_planes), M, lda, _ptr(b_planes), N, ldb, K, _ptr(bias), int(relu),
                                        _ptr(out), ldo, _ptr(planes), _pad8(N), splits, _ptr(ws),
                                        ws.numel() if ws is not None else 0, _stream(dev)), "ptrec_tc_gemm_split3")
    y = out[:, :N] if ldo != N else out
    return (y, planes) if want_planes else y


def tc_gemm_split3_tn(a_planes: torch.Tensor, M: int, b_planes: torch.Tensor, N: int, splits: int = 0) -> torch.Tensor:
    """out [M, N] fp32 = A^T B from row-major planes A [3, K, pad8(M)], B [3, K, pad8(N)] (reduction over the rows:
    the weight-gradient form dW = g^T x with K = batch).  ``splits=0`` picks the split-K count."""
    lib = _lib.load()
    _require_cuda(a_planes, b_planes)
    assert a_planes.dtype == torch.bfloat16 and b_planes.dtype == torch.bfloat16
    assert a_planes.dim() == 3 and b_planes.dim() == 3 and a_planes.is_contiguous() and b_planes.is_contiguous()
    K, lda = a_planes.shape[1], a_planes.shape[2]
    assert b_planes.shape[1] == K and lda >= M and b_planes.shape[2] >= N
    ldb = b_planes.shape[2]
    dev = a_planes.device
    if splits == 0:
        splits = lib.ptrec_tc_gemm_split3_default_splits(M, N, K)
    ldo = (N + 3) // 4 * 4
    out = torch.empty(M, ldo, dtype=torch.float32, device=dev)
    nbytes = lib.ptrec_tc_gemm_split3_workspace_bytes(M, ldo, splits)
    ws = _workspace("tc_gemm_split3", nbytes, dev) if nbytes else None
    _lib.check(lib.ptrec_tc_gemm_split3_tn(_ptr(a_planes), M, lda, _ptr(b_planes), N, ldb, K, _ptr(out), ldo, splits,
                                           _ptr(ws), ws.numel() if ws is not None else 0, _stream(dev)),
               "ptrec_tc_gemm_split3_tn")
    return out[:, :N] if ldo != N else out


# ---- fp16 x 2 operand mode of K6 ("h2"): x * s = h0 + h1 / 2^11 with a per-tensor power-of-two scale s --------------
_TC_MODES = ("bf16x3", "fp16x2")
_tc_mode = os.environ.get("PTREC_TC_MODE", "fp16x2")
if _tc_mode not in _TC_MODES:
    raise ValueError(f"PTREC_TC_MODE must be one of {_TC_MODES}, got {_tc_mode!r}")


def tc_mode() -> str:
    """Operand format of the K6 Linear GEMMs: ``fp16x2`` (default: two fp16 planes of the power-of-two-scaled operand,
    3 MMAs per product, 22 mantissa bits) or ``bf16x3`` (three exact bf16 planes, 6 MMAs per product)."""
    return _tc_mode


def set_tc_mode(mode: str) -> None:
    global _tc_mode
    if mode not in _TC_MODES:
        raise ValueError(f"invalid K6 operand mode {mode!r}; one of {_TC_MODES}")
    _tc_mode = mode


def tc_split2h(src: torch.Tensor, relu_ref: Optional[torch.Tensor] = None, want_planes: bool = True,
               want_t: bool = False, want_colsum: bool = False, absmax_in: Optional[torch.Tensor] = None):
    """fp32 [R, C] -> (planes [2, R, pad8(C)] fp16, planes_t [2, C, pad8(R)] fp16, colsum [C], scale [1] fp32).
    ``scale`` is the power of two the tensor was multiplied by before the split (largest magnitude -> [2^13, 2^14)),
    computed on the device from the tensor's absolute maximum; the GEMM divides it out again.  ``absmax_in``: a
    one-element fp32 tensor that already holds max |src| (``tc_gemm_split2h(..., want_absmax=True)``): skips the
    pass that finds it."""
    lib = _lib.load()
    _require_cuda(src, relu_ref)
    assert src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1
    R, C = src.shape
    dev = src.device
    if relu_ref is not None:
        assert relu_ref.shape == src.shape and relu_ref.dtype == torch.float32 and relu_ref.stride(1) == 1
    planes = torch.empty(2, R, _pad8(C), dtype=torch.float16, device=dev) if want_planes else None
    planes_t = torch.empty(2, C, _pad8(R), dtype=torch.float16, device=dev) if want_t else None
    colsum = torch.empty(C, dtype=torch.float32, device=dev) if want_colsum else None
    scale = torch.empty(1, dtype=torch.float32, device=dev)
    ws = _workspace("tc_split2h", lib.ptrec_tc_split2h_workspace_bytes(R, C), dev)
    if absmax_in is not None:
        _require_cuda(absmax_in)
        assert absmax_in.dtype == torch.float32 and absmax_in.numel() == 1
    _lib.check(lib.ptrec_tc_split2h(_ptr(src), src.stride(0), R, C, _ptr(relu_ref),
                                    relu_ref.stride(0) if relu_ref is not None else 0, _ptr(planes), _pad8(C),
                                    _ptr(planes_t), _pad8(R), _ptr(colsum), _ptr(scale), _ptr(absmax_in), _ptr(ws),
                                    ws.numel(), _stream(dev)), "ptrec_tc_split2h")
    return planes, planes_t, colsum, scale


def _check_h2(a_planes, scale_a, b_planes, scale_b):
    _require_cuda(a_planes, b_planes, scale_a, scale_b)
    assert a_planes.dtype == torch.float16 and b_planes.dtype == torch.float16
    assert a_planes.dim() == 3 and b_planes.dim() == 3 and a_planes.shape[0] == 2 and b_planes.shape[0] == 2
    assert a_planes.is_contiguous() and b_planes.is_contiguous()
    assert scale_a.dtype == torch.float32 and scale_b.dtype == torch.float32 and scale_a.numel() == 1 == scale_b.numel()


def tc_gemm_split2h(a_planes: torch.Tensor, scale_a: torch.Tensor, b_planes: torch.Tensor, scale_b: torch.Tensor,
                    K: int, bias: Optional[torch.Tensor] = None, relu: bool = False, splits: int = 1,
                    want_absmax: bool = False):
    """out [M, N] fp32 = A[M, K] B[N, K]^T (+ bias) (ReLU) from fp16 planes [2, M, lda], [2, N, ldb] and their scales.
    ``want_absmax`` (splits == 1): also return a one-element fp32 tensor that the epilogue raises to max |out|."""
    lib = _lib.load()
    _check_h2(a_planes, scale_a, b_planes, scale_b)
    M, lda = a_planes.shape[1], a_planes.shape[2]
    N, ldb = b_planes.shape[1], b_planes.shape[2]
    dev = a_planes.device
    if splits == 0:
        splits = lib.ptrec_tc_gemm_split3_default_splits(M, N, K)
    ldo = _pad8(N)  # 32-byte aligned rows: the epilogue stores 256 bits per thread
    out = torch.empty(M, ldo, dtype=torch.float32, device=dev)
    nbytes = lib.ptrec_tc_gemm_split3_workspace_bytes(M, ldo, splits)
    ws = _workspace("tc_gemm_split3", nbytes, dev) if nbytes else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    absmax = torch.zeros(1, dtype=torch.float32, device=dev) if want_absmax else None
    _lib.check(lib.ptrec_tc_gemm_split2h(_ptr(a_planes), _ptr(scale_a), M, lda, _ptr(b_planes), _ptr(scale_b), N, ldb, K,
                                         _ptr(bias), int(relu), _ptr(out), ldo, _ptr(absmax), splits, _ptr(ws),
                                         ws.numel() if ws is not None else 0, _stream(dev)), "ptrec_tc_gemm_split2h")
    y = out[:, :N] if ldo != N else out
    return (y, absmax) if want_absmax else y


def tc_gemm_split2h_tn(a_planes: torch.Tensor, scale_a: torch.Tensor, M: int, b_planes: torch.Tensor,
                       scale_b: torch.Tensor, N: int, splits: int = 0) -> torch.Tensor:
    """out [M, N] fp32 = A^T B from row-major fp16 planes A [2, K, pad8(M)], B [2, K, pad8(N)] (dW = g^T x)."""
    lib = _lib.load()
    _check_h2(a_planes, scale_a, b_planes, scale_b)
    K, lda = a_planes.shape[1], a_planes.shape[2]
    assert b_planes.shape[1] == K and lda >= M and b_planes.shape[2] >= N
    ldb = b_planes.shape[2]
    dev = a_planes.device
    if splits == 0:
        splits = lib.ptrec_tc_gemm_split3_default_splits(M, N, K)
    ldo = _pad8(N)
    out = torch.empty(M, ldo, dtype=torch.float32, device=dev)
    nbytes = lib.ptrec_tc_gemm_split3_workspace_bytes(M, ldo, splits)
    ws = _workspace("tc_gemm_split3", nbytes, dev) if nbytes else None
    _lib.check(lib.ptrec_tc_gemm_split2h_tn(_ptr(a_planes), _ptr(scale_a), M, lda, _ptr(b_planes), _ptr(scale_b), N, ldb,
                                            K, _ptr(out), ldo, splits, _ptr(ws), ws.numel() if ws is not None else 0,
                                            _stream(dev)), "ptrec_tc_gemm_split2h_tn")
    return out[:, :N] if ldo != N else out


# ---- the fused tower: carried scales, GEMM epilogues that write their consumer's operand planes -------------------
def tc_scale_roll(slots: torch.Tensor, err: torch.Tensor) -> torch.Tensor:
    """``slots`` fp32 [n, 2] = {scale, max since the last roll}: every slot's maximum becomes its next scale (8 binades
    of headroom), the maxima restart at zero; returns the scales as a fresh [n] tensor — what ONE forward / backward
    reads.  ``err`` (int32 word) is raised if a maximum left the fp16 range under the scale it was split with."""
    lib = _lib.load()
    _require_cuda(slots, err)
    assert slots.dtype == torch.float32 and slots.dim() == 2 and slots.shape[1] == 2 and slots.is_contiguous()
    assert err.dtype == torch.int32
    out = torch.empty(slots.shape[0], dtype=torch.float32, device=slots.device)
    _lib.check(lib.ptrec_tc_scale_roll(_ptr(slots), slots.shape[0], _ptr(out), _ptr(err), _stream(slots.device)),
               "ptrec_tc_scale_roll")
    return out


def tc_split2h_prescaled(src: torch.Tensor, scale: torch.Tensor, max_out: Optional[torch.Tensor],
                         relu_ref: Optional[torch.Tensor] = None, want_planes: bool = True, want_t: bool = False,
                         want_colsum: bool = False):
    """``tc_split2h`` with the scale given (a one-element view of ``tc_scale_roll``'s result): one kernel, no maximum
    pass.  ``max_out`` (one-element fp32 view of the slot's maximum) is raised to max |masked src|.
    Returns (planes, planes_t, colsum)."""
    lib = _lib.load()
    _require_cuda(src, relu_ref, scale, max_out)
    assert src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1
    assert scale.dtype == torch.float32 and scale.numel() == 1
    R, C = src.shape
    dev = src.device
    if relu_ref is not None:
        assert relu_ref.shape == src.shape and relu_ref.dtype == torch.float32 and relu_ref.stride(1) == 1
    planes = torch.empty(2, R, _pad8(C), dtype=torch.float16, device=dev) if want_planes else None
    planes_t = torch.empty(2, C, _pad8(R), dtype=torch.float16, device=dev) if want_t else None
    colsum = torch.empty(C, dtype=torch.float32, device=dev) if want_colsum else None
    ws = _workspace("tc_split2h", lib.ptrec_tc_split2h_workspace_bytes(R, C), dev) if want_colsum else None
    _lib.check(lib.ptrec_tc_split2h_prescaled(_ptr(src), src.stride(0), R, C, _ptr(relu_ref),
                                              relu_ref.stride(0) if relu_ref is not None else 0, _ptr(planes), _pad8(C),
                                              _ptr(planes_t), _pad8(R), _ptr(colsum), _ptr(scale), _ptr(max_out),
                                              _ptr(ws), ws.numel() if ws is not None else 0, _stream(dev)),
               "ptrec_tc_split2h_prescaled")
    return planes, planes_t, colsum


def tc_fused_supported() -> bool:
    """The fused tower runs on the CTA-pair kernel with 256-wide tiles (the default configuration)."""
    lib = _lib.load()
    return bool(lib.ptrec_tc_2sm_enabled()) and lib.ptrec_tc_get_bn() == 256


def mask_words(n: int) -> int:
    """Words per row of a ReLU bit mask over n columns (multiple of 4: rows are read 16 bytes at a time)."""
    return ((n + 31) // 32 + 3) // 4 * 4


def tc_gemm_split2h_fused(a_planes: torch.Tensor, scale_a: torch.Tensor, b_planes: torch.Tensor, scale_b: torch.Tensor,
                          K: int, bias: Optional[torch.Tensor] = None, relu: bool = False, want_out: bool = True,
                          out_scale: Optional[torch.Tensor] = None, mask_in: Optional[torch.Tensor] = None,
                          want_mask: bool = False, want_colsum: bool = False, max_out: Optional[torch.Tensor] = None,
                          ws_tag: str = ""):
    """A[M, K] B[N, K]^T (+ bias) (ReLU) (* mask_in) on the CTA-pair fp16 x 2 kernel, handed over in the consumer's
    format.  Returns (out fp32 [M, N] or None, planes [2, M, pad16(N)] fp16 split with ``out_scale`` or None,
    mask int32 [M, mask_words(N)] of the positive entries or None, colsum [N] or None)."""
    lib = _lib.load()
    _check_h2(a_planes, scale_a, b_planes, scale_b)
    M, lda = a_planes.shape[1], a_planes.shape[2]
    N, ldb = b_planes.shape[1], b_planes.shape[2]
    dev = a_planes.device
    ldo = _pad8(N)            # 32-byte aligned rows of both outputs: the epilogue stores 256 bits per thread
    pld = (N + 15) // 16 * 16
    out = torch.empty(M, ldo, dtype=torch.float32, device=dev) if want_out else None
    planes = torch.empty(2, M, pld, dtype=torch.float16, device=dev) if out_scale is not None else None
    mw = mask_words(N)
    mask = torch.empty(M, mw, dtype=torch.int32, device=dev) if want_mask else None
    if mask_in is not None:
        assert mask_in.dtype == torch.int32 and tuple(mask_in.shape) == (M, mw) and mask_in.is_contiguous()
    colsum = torch.empty(N, dtype=torch.float32, device=dev) if want_colsum else None
    # ws_tag: calls whose column-sum reduce may still be in flight on another stream keep separate partial buffers
    ws = _workspace("tc_gemm_fused" + ws_tag, lib.ptrec_tc_gemm_fused_workspace_bytes(M, N), dev) if want_colsum else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    _lib.check(lib.ptrec_tc_gemm_split2h_fused(
        _ptr(a_planes), _ptr(scale_a), M, lda, _ptr(b_planes), _ptr(scale_b), N, ldb, K, _ptr(bias), int(relu),
        _ptr(out), ldo, _ptr(planes), pld, _ptr(out_scale), _ptr(mask_in), _ptr(mask), mw, _ptr(colsum),
        _ptr(max_out), _ptr(ws), ws.numel() if ws is not None else 0, _stream(dev)), "ptrec_tc_gemm_split2h_fused")
    y = None if out is None else (out[:, :N] if ldo != N else out)
    return y, planes, mask, colsum


# ----------------------------------------------------------------------------------------------
# K9 BCEWithLogitsLoss(reduction='mean'): forward + gradient in one launch
# ----------------------------------------------------------------------------------------------
_BCE_WS: Dict[tuple, torch.Tensor] = {}


class _BCELogitsMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        lib = _lib.load()
        dev = logits.device
        key = (dev, torch.cuda.current_stream(dev).cuda_stream)
        ws = _BCE_WS.get(key)
        if ws is None:   # zero once: the kernel leaves its ticket counter at zero
            ws = _BCE_WS[key] = torch.zeros(lib.ptrec_bce_logits_workspace_bytes(), dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        grad = torch.empty_like(logits) if ctx.needs_input_grad[0] else None
        _lib.check(lib.ptrec_bce_logits_mean(_ptr(logits), _ptr(target), logits.numel(), _ptr(loss), _ptr(grad),
                                             _ptr(ws), ws.numel(), _stream(dev)), "ptrec_bce_logits_mean")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def bce_logits_mean(logits: torch.Tensor, target: torch.Tensor) -> Optional[torch.Tensor]:
    """``BCEWithLogitsLoss()(logits, target)`` (mean reduction, no weights) in one launch; None if the inputs are not
    contiguous fp32 CUDA tensors of one shape (the caller then uses the module)."""
    if not (logits.is_cuda and target.is_cuda and logits.dtype == torch.float32 and target.dtype == torch.float32
            and logits.shape == target.shape and logits.is_contiguous() and target.is_contiguous()
            and logits.numel() >= 1 and not target.requires_grad):
        return None
    return _BCELogitsMean.apply(logits, target)


# ----------------------------------------------------------------------------------------------
# K5 DCN-v2 cross layers (tcgen05)
# ----------------------------------------------------------------------------------------------
def _check_bf16_2d(*ts):
    for t in ts:
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and t.data_ptr() % 16 == 0


def set_dcn_2sm(enabled: bool) -> None:
    """K5 kernel choice: the CTA-pair GEMM of K6 with one bf16 plane (default) or the 128 x 128 single-CTA kernel of
    round 1 (csrc/dcn_cross.cu); same results up to summation order."""
    _lib.load().ptrec_set_dcn_2sm(1 if enabled else 0)


def dcn_2sm_enabled() -> bool:
    return bool(_lib.load().ptrec_dcn_2sm_enabled())


def dcn_cross_fwd(x_l: torch.Tensor, x0: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                  want_u: bool = True):
    """out = x0 * (x_l @ weight.T + bias) + x_l  (bf16 [B, d]); returns (out, u or None)."""
    lib = _lib.load()
    _require_cuda(x_l, x0, weight, bias)
    _check_bf16_2d(x_l, x0, weight)
    B, d = x_l.shape
    assert x0.shape == x_l.shape and x0.stride(0) == x_l.stride(0) and weight.shape == (d, d) and weight.is_contiguous()
    out = torch.empty_strided((B, d), (x_l.stride(0), 1), dtype=torch.bfloat16, device=x_l.device)
    u = torch.empty_strided((B, d), (x_l.stride(0), 1), dtype=torch.bfloat16, device=x_l.device) if want_u else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == d
    _lib.check(lib.ptrec_dcn_cross_fwd(_ptr(x_l), _ptr(x0), _ptr(weight), _ptr(bias), B, d, x_l.stride(0), _ptr(out),
                                       _ptr(u), _stream(x_l.device)), "ptrec_dcn_cross_fwd")
    return out, u


def dcn_cross_dgrad(g_u: torch.Tensor, weight_t: torch.Tensor, g_out: torch.Tensor, x0: torch.Tensor,
                    want_prev: bool = True):
    """g_x = g_u @ W + g_out ; optionally also g_x * x0 (next g_u).  weight_t = W.T contiguous."""
    lib = _lib.load()
    _require_cuda(g_u, weight_t, g_out, x0)
    _check_bf16_2d(g_u, weight_t, g_out, x0)
    B, d = g_u.shape
    ld = g_u.stride(0)
    assert g_out.stride(0) == ld and x0.stride(0) == ld and weight_t.is_contiguous()
    g_x = torch.empty_strided((B, d), (ld, 1), dtype=torch.bfloat16, device=g_u.device)
    prev = torch.empty_strided((B, d), (ld, 1), dtype=torch.bfloat16, device=g_u.device) if want_prev else None
    _lib.check(lib.ptrec_dcn_cross_dgrad(_ptr(g_u), _ptr(weight_t), _ptr(g_out), _ptr(x0), B, d, ld, _ptr(g_x),
                                         _ptr(prev), _stream(g_u.device)), "ptrec_dcn_cross_dgrad")
    return g_x, prev


def dcn_cross_wgrad(g_u: torch.Tensor, x_l: torch.Tensor) -> torch.Tensor:
    """grad_W [d, d] fp32 = g_u.T @ x_l."""
    lib = _lib.load()
    _require_cuda(g_u, x_l)
    _check_bf16_2d(g_u, x_l)
    B, d = g_u.shape
    assert x_l.stride(0) == g_u.stride(0)
    gw = torch.empty(d, d, dtype=torch.float32, device=g_u.device)
    ws = _workspace("dcn_wgrad", lib.ptrec_dcn_cross_wgrad_workspace_bytes(B, d), g_u.device)
    _lib.check(lib.ptrec_dcn_cross_wgrad(_ptr(g_u), _ptr(x_l), B, d, g_u.stride(0), _ptr(gw), _ptr(ws), ws.numel(),
                                         _stream(g_u.device)), "ptrec_dcn_cross_wgrad")
    return gw


class _CrossNet(torch.autograd.Function):
    """All cross layers in one autograd node so that each dgrad epilogue can emit the g_u of the layer below.  Every
    element-wise step around the GEMMs (padding / casting, g_out * x0, the x0-gradient accumulation, the bias
    gradients, the transposed weights) is one launch of csrc/dcn_glue.cu."""

    @staticmethod
    def forward(ctx, x0f, n_layers, head_w, *params):
        """head_w: None -> the fp32 [B, d] output of the last layer; an fp32 [d] vector -> y [B] = x_L . head_w (the cross
        half of DCN's closing Linear) computed from the bf16 x_L, with its backward fused into the chain's first step."""
        lib = _lib.load()
        weights, biases = params[:n_layers], params[n_layers:]
        B, d = x0f.shape
        dp = (d + 7) // 8 * 8
        dev = x0f.device
        st = _stream(dev)
        if x0f.stride(1) != 1:
            x0f = x0f.contiguous()
        x0 = torch.empty(B, dp, dtype=torch.bfloat16, device=dev)
        _lib.check(lib.ptrec_dcn_pack_input(_ptr(x0f), x0f.stride(0), B, d, dp, _ptr(x0), st), "ptrec_dcn_pack_input")
        wts, xs, us = [], [x0], []
        x = x0
        for W, b in zip(weights, biases):
            w16 = torch.empty(dp, dp, dtype=torch.bfloat16, device=dev)
            w16t = torch.empty(dp, dp, dtype=torch.bfloat16, device=dev)
            bp = torch.empty(dp, dtype=torch.float32, device=dev)
            Wc, bc = W.detach().contiguous(), b.detach().contiguous()
            _lib.check(lib.ptrec_dcn_prep_weight(_ptr(Wc), _ptr(bc), d, dp, _ptr(w16), _ptr(w16t), _ptr(bp), st),
                       "ptrec_dcn_prep_weight")
            x, u = dcn_cross_fwd(x, x0, w16, bp, want_u=True)
            wts.append(w16t)
            xs.append(x)
            us.append(u)
        ctx.d, ctx.n, ctx.head = d, n_layers, head_w is not None
        ctx.params = params
        if head_w is not None:
            hw = head_w.detach().float().contiguous()
            y = torch.empty(B, dtype=torch.float32, device=dev)
            _lib.check(lib.ptrec_dcn_head_fwd(_ptr(x), B, d, dp, _ptr(hw), _ptr(y), st), "ptrec_dcn_head_fwd")
            ctx.save_for_backward(*xs[:-1], *us, *wts, x, hw)
            return y.to(x0f.dtype)
        out = torch.empty(B, d, dtype=torch.float32, device=dev)
        _lib.check(lib.ptrec_dcn_unpack(_ptr(x), B, d, dp, _ptr(out), st), "ptrec_dcn_unpack")
        ctx.save_for_backward(*xs[:-1], *us, *wts)
        return out.to(x0f.dtype)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        n, d = ctx.n, ctx.d
        saved = ctx.saved_tensors
        xs, us, wts = saved[:n], saved[n:2 * n], saved[2 * n:3 * n]
        x0 = xs[0]
        B, dp = x0.shape
        dev = g.device
        st = _stream(dev)
        g = g.float()
        g_out = torch.empty(B, dp, dtype=torch.bfloat16, device=dev)
        g_u = torch.empty(B, dp, dtype=torch.bfloat16, device=dev)
        ws = _workspace("dcn_bwd_layer", lib.ptrec_dcn_bwd_layer_workspace_bytes(B, dp), dev)
        g_head = None
        if ctx.head:
            x_last, hw = saved[3 * n], saved[3 * n + 1]
            g = g.contiguous()
            g_head = torch.empty(d, dtype=torch.float32, device=dev)
            _lib.check(lib.ptrec_dcn_head_bwd(_ptr(g), _ptr(hw), _ptr(x_last), _ptr(x0), B, d, dp, _ptr(g_out), _ptr(g_u),
                                              _ptr(g_head), _ptr(ws), ws.numel(), st), "ptrec_dcn_head_bwd")
        else:
            if g.stride(1) != 1:
                g = g.contiguous()
            _lib.check(lib.ptrec_dcn_bwd_init(_ptr(g), g.stride(0), _ptr(x0), B, d, dp, _ptr(g_out), _ptr(g_u), st),
                       "ptrec_dcn_bwd_init")
        g_x0 = torch.empty(B, dp, dtype=torch.float32, device=dev)
        gws, gbs = [None] * n, [None] * n
        # the weight-gradient GEMMs feed nothing but the optimizer: inside an IModel train step they run on the tower's
        # weight-gradient side stream beside the chain (joined before the optimizer step), as in dense._TcMLP.backward
        from .model.layer import dense as _dense
        side = None
        if _dense._DEFER_JOIN[0] and all(p.grad is None for p in ctx.params):
            side = _dense._wgrad_stream(dev)
        if side is not None:
            from .model.layer.embedding import register_join_stream
            main = torch.cuda.current_stream(dev)
            register_join_stream(dev, side)
        for l in range(n - 1, -1, -1):
            gbs[l] = torch.empty(d, dtype=torch.float32, device=dev)
            # d out / d x0 = u: g_x0 (+)= g_out * u_l; bias gradient = column sums of g_u
            _lib.check(lib.ptrec_dcn_bwd_layer(_ptr(g_out), _ptr(us[l]), _ptr(g_u), B, d, dp, _ptr(g_x0), int(l != n - 1),
                                               _ptr(gbs[l]), _ptr(ws), ws.numel(), st), "ptrec_dcn_bwd_layer")
            if side is not None:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    gws[l] = dcn_cross_wgrad(g_u, xs[l])[:d, :d].contiguous()
                g_u.record_stream(side)
                xs[l].record_stream(side)
                gws[l].record_stream(main)
            else:
                gws[l] = dcn_cross_wgrad(g_u, xs[l])[:d, :d]
            g_out, g_u = dcn_cross_dgrad(g_u, wts[l], g_out, x0, want_prev=l > 0)
        out = torch.empty(B, d, dtype=torch.float32, device=dev)  # x0 is also layer 0's x_l: + the chain's g_out
        _lib.check(lib.ptrec_dcn_bwd_final(_ptr(g_x0), _ptr(g_out), B, d, dp, _ptr(out), st), "ptrec_dcn_bwd_final")
        return (out, None, g_head, *gws, *gbs)


def cross_net(x0: torch.Tensor, weights, biases) -> torch.Tensor:
    """L DCN-v2 cross layers ``x <- x0 * (x W_l^T + b_l) + x`` on tensor cores (bf16, fp32 accumulate)."""
    _require_cuda(x0)
    return _CrossNet.apply(x0.contiguous(), len(weights), None, *weights, *biases)


def cross_net_head(x0: torch.Tensor, weights, biases, head_w: torch.Tensor) -> torch.Tensor:
    """``cross_net(x0) @ head_w`` ([B]; head_w fp32 [d]) without the fp32 [B, d] output: the row dot reads the last
    layer's bf16 result, its backward writes the chain's bf16 gradient operands directly."""
    _require_cuda(x0, head_w)
    assert head_w.dim() == 1 and head_w.numel() == x0.shape[1]
    return _CrossNet.apply(x0.contiguous(), len(weights), head_w, *weights, *biases)


# ----------------------------------------------------------------------------------------------
# K4 DIN attention pooling
# ----------------------------------------------------------------------------------------------
def _din_args(q, keys, lens, params):
    W1, b1, W2, b2, W3, b3 = params
    B, L, DQ = keys.shape
    H1, H2 = W1.shape[0], W2.shape[0]
    assert q.shape == (B, DQ) and q.stride(1) == 1 and keys.stride(2) == 1
    assert W1.shape == (H1, 4 * DQ) and W2.shape == (H2, H1) and W3.shape == (1, H2)
    for p in params:
        assert p.dtype == torch.float32 and p.is_contiguous()
    if lens is not None:
        lens = lens.to(torch.int32).contiguous()
    return B, L, DQ, H1, H2, lens


def din_attn_pool_fwd(q, keys, lens, params, want_scores: bool = False):
    lib = _lib.load()
    _require_cuda(q, keys, lens, *params)
    B, L, DQ, H1, H2, lens = _din_args(q, keys, lens, params)
    out = torch.empty(B, DQ, dtype=torch.float32, device=q.device)
    scores = torch.empty(B, L, dtype=torch.float32, device=q.device) if want_scores else None
    _lib.check(lib.ptrec_din_attn_pool_fwd(_ptr(q), q.stride(0), _ptr(keys), keys.stride(0), keys.stride(1), _ptr(lens),
                                           B, L, DQ, H1, H2, *[_ptr(p) for p in params], _ptr(out), _ptr(scores),
                                           _stream(q.device)), "ptrec_din_attn_pool_fwd")
    return out, scores


def din_attn_pool_bwd(q, keys, lens, params, g_pooled):
    lib = _lib.load()
    _require_cuda(q, keys, lens, g_pooled, *params)
    B, L, DQ, H1, H2, lens = _din_args(q, keys, lens, params)
    g_pooled = g_pooled.contiguous()
    g_q = torch.empty(B, DQ, dtype=torch.float32, device=q.device)
    g_keys = torch.empty(B, L, DQ, dtype=torch.float32, device=q.device)
    n = lib.ptrec_din_attn_pool_grad_floats(DQ, H1, H2)
    flat = torch.empty(n, dtype=torch.float32, device=q.device)
    ws = _workspace("din_bwd", lib.ptrec_din_attn_pool_bwd_workspace_bytes(B, DQ, H1, H2), q.device)
    _lib.check(lib.ptrec_din_attn_pool_bwd(_ptr(q), q.stride(0), _ptr(keys), keys.stride(0), keys.stride(1), _ptr(lens),
                                           B, L, DQ, H1, H2, *[_ptr(p) for p in params], _ptr(g_pooled), _ptr(g_q),
                                           _ptr(g_keys), g_keys.stride(0), g_keys.stride(1), _ptr(flat), _ptr(ws),
                                           ws.numel(), _stream(q.device)), "ptrec_din_attn_pool_bwd")
    sizes = [H1 * 4 * DQ, H1, H2 * H1, H2, H2, 1]
    gW1, gb1, gW2, gb2, gW3, gb3 = torch.split(flat, sizes)
    return g_q, g_keys, (gW1.view(H1, 4 * DQ), gb1, gW2.view(H2, H1), gb2, gW3.view(1, H2), gb3)


def _din_id_args(tables, ids, DQ):
    t0, t1 = tables
    i0, i1 = ids
    for t in (t0, t1):
        assert t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 and t.shape[1] == DQ // 2
    for i in (i0, i1):
        assert i.dtype == torch.int64 and i.is_contiguous()
    return (_ptr(t0), _ptr(t1), t0.stride(0), t1.stride(0), t0.shape[0], t1.shape[0], _ptr(i0), _ptr(i1))


def din_ids_supported(DQ: int, H1: int, H2: int) -> bool:
    """K4 reads its keys by id (no gathered [B, L, DQ] tensor) where forward AND backward have a tensor-core build."""
    return _lib.load().ptrec_din_tc_enabled() == 3 and DQ == 32 and (H1, H2) in ((80, 40), (64, 32))


def din_attn_pool_fwd_ids(q, tables, ids, ids_stride_b: int, ids_offset: int, err_flag, lens, L: int, params):
    """``din_attn_pool_fwd`` with key (b, l) = [tables[0][ids[0][b * ids_stride_b + ids_offset + l]] | tables[1][...]]."""
    lib = _lib.load()
    _require_cuda(q, *tables, *ids, lens, *params)
    W1, b1, W2, b2, W3, b3 = params
    B, DQ = q.shape
    H1, H2 = W1.shape[0], W2.shape[0]
    if lens is not None:
        lens = lens.to(torch.int32).contiguous()
    out = torch.empty(B, DQ, dtype=torch.float32, device=q.device)
    _lib.check(lib.ptrec_din_attn_pool_fwd_ids(_ptr(q), q.stride(0), *_din_id_args(tables, ids, DQ), ids_stride_b, ids_offset,
                                               _ptr(err_flag), _ptr(lens), B, L, DQ, H1, H2, *[_ptr(p) for p in params],
                                               _ptr(out), None, _stream(q.device)), "ptrec_din_attn_pool_fwd_ids")
    return out


def din_attn_pool_bwd_ids(q, tables, ids, ids_stride_b: int, ids_offset: int, lens, L: int, params, g_pooled, g_keys):
    """Backward of the above; ``g_keys`` is a caller-provided [B, L, DQ] view (strides honoured) that receives the key
    gradients as dense rows.  Returns (g_q, parameter gradients)."""
    lib = _lib.load()
    _require_cuda(q, g_pooled, g_keys, *tables, *ids, lens, *params)
    W1, b1, W2, b2, W3, b3 = params
    B, DQ = q.shape
    H1, H2 = W1.shape[0], W2.shape[0]
    if lens is not None:
        lens = lens.to(torch.int32).contiguous()
    g_pooled = g_pooled.contiguous()
    g_q = torch.empty(B, DQ, dtype=torch.float32, device=q.device)
    n = lib.ptrec_din_attn_pool_grad_floats(DQ, H1, H2)
    flat = torch.empty(n, dtype=torch.float32, device=q.device)
    ws = _workspace("din_bwd", lib.ptrec_din_attn_pool_bwd_workspace_bytes(B, DQ, H1, H2), q.device)
    _lib.check(lib.ptrec_din_attn_pool_bwd_ids(_ptr(q), q.stride(0), *_din_id_args(tables, ids, DQ), ids_stride_b, ids_offset,
                                               _ptr(lens), B, L, DQ, H1, H2, *[_ptr(p) for p in params], _ptr(g_pooled),
                                               _ptr(g_q), _ptr(g_keys), g_keys.stride(0), g_keys.stride(1), _ptr(flat),
                                               _ptr(ws), ws.numel(), _stream(q.device)), "ptrec_din_attn_pool_bwd_ids")
    sizes = [H1 * 4 * DQ, H1, H2 * H1, H2, H2, 1]
    gW1, gb1, gW2, gb2, gW3, gb3 = torch.split(flat, sizes)
    return g_q, (gW1.view(H1, 4 * DQ), gb1, gW2.view(H2, H1), gb2, gW3.view(1, H2), gb3)


class _DinAttnPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, keys, lens, W1, b1, W2, b2, W3, b3):
        params = (W1, b1, W2, b2, W3, b3)
        out, _ = din_attn_pool_fwd(q, keys, lens, params)
        ctx.save_for_backward(q, keys, lens, *params)
        return out

    @staticmethod
    def backward(ctx, g):
        q, keys, lens, *params = ctx.saved_tensors
        g_q, g_keys, gp = din_attn_pool_bwd(q, keys, lens, tuple(params), g)
        return (g_q, g_keys, None, *gp)


def din_attn_pool(q, keys, lens, W1, b1, W2, b2, W3, b3):
    """DIN attention pooling ``sum_{l<len} a_l k_l`` with the fused activation unit (CUDA forward + backward).
    ``q [B, DQ]``, ``keys [B, L, DQ]`` (may be strided views), ``lens [B]`` or None."""
    _require_cuda(q, keys)
    if q.stride(-1) != 1 or q.data_ptr() % 16 or q.stride(0) % 4:
        q = q.contiguous()
    if keys.stride(-1) != 1 or keys.data_ptr() % 16 or keys.stride(0) % 4 or keys.stride(1) % 4:
        keys = keys.contiguous()
    return _DinAttnPool.apply(q, keys, lens, W1, b1, W2, b2, W3, b3)
