"""Dense tower blocks.  Same structure and parameter names as the reference
(torchrec/model/layer/Dense.py:4-24, MLP.py:8-23): ``Dense`` = Linear -> ReLU -> Dropout (the
``activation`` argument is accepted and, as in the reference, always resolves to ReLU);
``MLP`` = ``Sequential`` of ``dense_{i}`` blocks under the attribute ``mlp``.

On a CUDA device the Linear -> ReLU pair and its backward run on the tcgen05 tensor cores through K6
(csrc/tc_linear.cu) at fp32-level error (the reference computes these GEMMs in fp32 and the north star asks for
1e-5), 3-5x faster than the fp32 SIMT sgemm behind ``nn.Linear``.  Default operand format ``fp16x2``: each fp32
operand is scaled by a per-tensor power of two and split into two fp16 planes (22 mantissa bits), three MMAs per
product.  ``PTREC_TC_MODE=bf16x3`` (``ops.set_tc_mode``) selects the exact three-plane bf16 split instead (six plane
pairs accumulated in fp32; include/ptrec_b200.h).  The ReLU
mask and the bias gradient are fused into the pass that splits the incoming gradient.  ``PTREC_TC_LINEAR=0`` selects
the stock ``nn.Linear`` path (cuBLAS fp32) for A/B measurements."""
import os
from typing import List

import torch
from torch.nn import Dropout, Linear, Module, ReLU, Sequential

from ... import ops


class _TcLinearReLU(torch.autograd.Function):
    """y = relu(x W^T + b) on K6.  Saves y (ReLU mask), the planes of x (weight gradient: read MN-major, i.e.
    transposed by the tensor core itself) and the transposed planes of W (input gradient).  ``px``: the planes of x if
    the producer already wrote them (K8 head, previous layer's epilogue); ``emit_planes``: have the epilogue write
    the planes of y for the next layer.  Returns (y, planes of y or an empty tensor).  In the fp16 x 2 operand mode
    ``px`` is instead the device word holding max |x| when the producer of x left one, and the second output is the
    word holding max |y| (reduced by the GEMM epilogue), so that only the first layer's input and the gradient
    entering the last layer need a pass of their own to find their scale."""

    @staticmethod
    def forward(ctx, x, weight, bias, px, emit_planes):
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        N, K = weight.shape
        ctx.h2 = ops.tc_mode() == "fp16x2"
        if ctx.h2:  # two fp16 planes of the scaled operand + the device-side scale word (ops.tc_split2h)
            # px here = the word holding max |x| if the producer of x (the previous layer's epilogue) left one
            px, _, _, sx = ops.tc_split2h(x, absmax_in=px)
            pw, pwt, _, sw = ops.tc_split2h(weight, want_planes=True, want_t=need_dx)
            y, py = ops.tc_gemm_split2h(px, sx, pw, sw, K, bias=bias, relu=True, want_absmax=True)  # py: max |y|
            ctx.save_for_backward(y, px if need_dw else None, pwt, sx, sw)
        else:
            if px is None:
                px, _, _ = ops.tc_split3(x)
            pw, pwt, _ = ops.tc_split3(weight, want_planes=True, want_t=need_dx)
            if emit_planes:
                y, py = ops.tc_gemm_split3(px, pw, K, bias=bias, relu=True, want_planes=True)
            else:
                y, py = ops.tc_gemm_split3(px, pw, K, bias=bias, relu=True), x.new_empty(0, dtype=torch.bfloat16)
            ctx.save_for_backward(y, px if need_dw else None, pwt)
        ctx.dims = (x.shape[0], N, K)
        ctx.has_bias = bias is not None
        ctx.mark_non_differentiable(py)
        return y, py

    @staticmethod
    def backward(ctx, gy, _gpy):
        B, N, K = ctx.dims
        need_dx, need_dw, need_db = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        if gy.stride(-1) != 1:
            gy = gy.contiguous()
        dw = None
        if ctx.h2:
            y, px, pwt, sx, sw = ctx.saved_tensors
            pg, _, db, sg = ops.tc_split2h(gy, relu_ref=y, want_planes=need_dx or need_dw, want_colsum=need_db,
                                           absmax_in=_carried_absmax(gy))
            dx = None
            if need_dx:                                                             # g W        [B, K]
                dx, am = ops.tc_gemm_split2h(pg, sg, pwt, sw, N, want_absmax=True)
                _carry_absmax(dx, am)   # the layer below splits dx next: it finds max |dx| already reduced
            if need_dw:
                dw = ops.tc_gemm_split2h_tn(pg, sg, N, px, sx, K)                   # g^T x      [N, K]
        else:
            y, px, pwt = ctx.saved_tensors
            pg, _, db = ops.tc_split3(gy, relu_ref=y, want_planes=need_dx or need_dw, want_colsum=need_db)
            dx = ops.tc_gemm_split3(pg, pwt, N) if need_dx else None                # g W        [B, K]
            if need_dw:
                dw = ops.tc_gemm_split3_tn(pg, N, px, K)                            # g^T x      [N, K]
        if dw is not None and not dw.is_contiguous():
            dw = dw.contiguous()
        return dx, dw, db, None, None


def _carry_absmax(t: torch.Tensor, word: torch.Tensor) -> None:
    """Attach the device word holding max |t| (written by the GEMM epilogue that produced ``t``) to the tensor object,
    with the storage address and version it describes."""
    t._ptrec_absmax = (word, t.data_ptr(), t._version, tuple(t.shape))


def _carried_absmax(t: torch.Tensor):
    """The word attached by ``_carry_absmax`` if it still describes ``t`` (same storage, not modified since); else None
    — the split then finds the maximum itself."""
    rec = getattr(t, "_ptrec_absmax", None)
    if rec is None or rec[1] != t.data_ptr() or rec[2] != t._version or rec[3] != tuple(t.shape):
        return None
    return rec[0]


def tc_linear_enabled() -> bool:
    return os.environ.get("PTREC_TC_LINEAR", "1") != "0"


# Below this many multiply-accumulates (batch * in * out) the three launches of the K6 path (split x, split W, GEMM)
# cost more than the fp32 cuBLAS sgemm they replace (measured: DIN's 8192 x ~100 x 200 layers), so small layers stay
# on nn.Linear.  Tests set it to 0 to push small models through K6.
TC_MIN_MACS = int(os.environ.get("PTREC_TC_LINEAR_MIN_MACS", str(1 << 28)))


class Dense(Module):
    def __init__(self, input_units: int, output_units: int, activation: str, dropout: float):
        super().__init__()
        self.linear = Linear(input_units, output_units)
        self.activation = ReLU()
        self.dropout = Dropout(dropout)
        self.emit_planes = False  # set by MLP: the next Dense consumes this layer's output planes (no split pass)

    def forward(self, x):
        w = self.linear.weight
        if x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32 and tc_linear_enabled() \
                and not torch.is_autocast_enabled() and x.numel() * w.shape[0] >= TC_MIN_MACS:
            lead = x.shape[:-1]
            x2 = x if x.dim() == 2 else x.reshape(-1, x.shape[-1])
            if x2.stride(-1) != 1:
                x2 = x2.contiguous()
            h2 = ops.tc_mode() == "fp16x2"
            if h2:   # fp16 x 2: what travels with a tensor is the word holding its absolute maximum
                px = _carried_absmax(x2)
            else:
                px = getattr(x, "_ptrec_planes", None)  # written by the producer of x (K8 head / previous layer)
                if px is not None and not (x.dim() == 2 and px.dtype == torch.bfloat16 and px.device == x2.device
                                           and tuple(px.shape) == (3, x2.shape[0], (x2.shape[1] + 7) // 8 * 8)):
                    px = None
            y, py = _TcLinearReLU.apply(x2, w, self.linear.bias, px, self.emit_planes and x.dim() == 2 and not h2)
            if self.training and self.dropout.p > 0:
                return self.dropout(y.reshape(*lead, w.shape[0]))
            if h2:
                _carry_absmax(y, py)
            elif py.numel():
                y._ptrec_planes = py
            return y.reshape(*lead, w.shape[0]) if x.dim() != 2 else y
        return self.dropout(self.activation(self.linear(x)))


class MLP(Module):
    def __init__(self, input_units: int, hidden_units_list: List[int], activation: str, dropout: float):
        super().__init__()
        self.mlp = Sequential()
        units = input_units
        for index, hidden_units in enumerate(hidden_units_list):
            self.mlp.add_module(f"dense_{index}", Dense(units, hidden_units, activation, dropout))
            units = hidden_units
        # Dense.emit_planes (the GEMM epilogue writes the next layer's planes) stays off: the epilogue stores are
        # per-thread rows (8-byte pieces, uncoalesced) and measured slower than the separate split pass
        # (cfg2 step 0.72 -> 0.85 ms); it needs a shared-memory-staged epilogue first.

    def forward(self, x):
        return self.mlp(x)
