"""Feature-interaction layers (fused CUDA forward + backward)."""
import torch
from torch import Tensor, nn

from ... import ops


class FMSecondOrder(nn.Module):
    """``y[b] = 0.5 * sum_k ((sum_f v[b,f,k])^2 - sum_f v[b,f,k]^2)`` for ``v [B, F, D]``.
    The F = 2 case is the reference's matrix-factorisation dot ``(u * i).sum(-1)``
    (torchrec/model/FunkSVD.py:51,62)."""

    def forward(self, v: Tensor) -> Tensor:
        return ops.fm2(v)


class _FMHead(torch.autograd.Function):
    """K8: first-order sum + FM second order + dense linear term + bias -> logit, and the tower input [v | x]."""

    @staticmethod
    def forward(ctx, v2d, w1, x, wd, bias, F, D, want_deep_in, want_planes, call=None):
        planes = None
        if call is not None:
            # the K6 fused tower reads its input as fp16 planes: write those, and no fp32 copy (deep_in is an
            # allocation that only carries the autograd edge; MLP.forward pairs it with the call record)
            sl = call.mlp._scales
            logit, call.in_planes = ops.fm_head_fwd_h2(v2d, w1, x, wd, bias, F, D, call.cs[sl.i_x():sl.i_x() + 1],
                                                       sl.max_word(sl.i_x()))
            nd = 0 if x is None else x.shape[1]
            deep_in = v2d.new_empty(v2d.shape[0], (F * D + nd + 3) // 4 * 4)[:, :F * D + nd]
        elif want_deep_in and want_planes:
            logit, deep_in, planes = ops.fm_head_fwd(v2d, w1, x, wd, bias, F, D, True, True)
        else:
            logit, deep_in = ops.fm_head_fwd(v2d, w1, x, wd, bias, F, D, want_deep_in)
        ctx.save_for_backward(v2d, x, wd)
        ctx.meta = (F, D)
        if deep_in is None:
            deep_in = v2d.new_empty(0)
            ctx.mark_non_differentiable(deep_in)
        if planes is None:
            planes = v2d.new_empty(0, dtype=torch.bfloat16)
        ctx.mark_non_differentiable(planes)
        return logit, deep_in, planes

    @staticmethod
    def backward(ctx, g_logit, g_deep_in, _g_planes):
        v2d, x, wd = ctx.saved_tensors
        F, D = ctx.meta
        need = ctx.needs_input_grad
        if g_deep_in is not None and g_deep_in.numel() == 0:
            g_deep_in = None
        if g_deep_in is not None and not (g_deep_in.stride(1) == 1 and g_deep_in.stride(0) % 4 == 0
                                          and g_deep_in.data_ptr() % 16 == 0):
            g_deep_in = g_deep_in.contiguous()
            if g_deep_in.stride(0) % 4 != 0:  # odd width: re-pitch
                buf = g_deep_in.new_empty(g_deep_in.shape[0], (g_deep_in.shape[1] + 3) // 4 * 4)
                buf[:, :g_deep_in.shape[1]] = g_deep_in
                g_deep_in = buf[:, :g_deep_in.shape[1]]
        from .dense import side_reductions
        with side_reductions(g_logit.device) as sr:   # the partial reduction (grad wd, grad bias) leaves the chain
            gv, gw1, gwd, gb = ops.fm_head_bwd(v2d, x, wd, g_logit.contiguous(), g_deep_in, F, D, want_w1=need[1],
                                               want_wd=need[3] and wd is not None, want_bias=need[4])
            sr.adopt(gwd, gb)
        return gv if need[0] else None, gw1, None, gwd, gb, None, None, None, None, None


class _RowDot(torch.autograd.Function):
    """K8: y[b] = h[b, :] . w — the Linear(H, 1, bias=False) closing a tower.  ``call``: h is the output of a K6 fused
    tower (dense._TowerCall) and has no other consumer: the backward writes the gradient of the tower's last
    pre-activation as that tower's operand planes (+ its bias gradient) instead of an fp32 gradient of h."""

    @staticmethod
    def forward(ctx, h, w, call=None):
        ctx.save_for_backward(h, w)
        ctx.call = call
        return ops.rowdot_fwd(h, w)

    @staticmethod
    def backward(ctx, g):
        h, w = ctx.saved_tensors
        call = ctx.call
        if call is not None and ctx.needs_input_grad[0]:
            sl = call.mlp._scales
            i = sl.i_g(sl.L - 1)
            from .dense import side_reductions
            with side_reductions(g.device) as sr:   # the two partial reductions (last bias gradient, grad w) leave the chain
                planes, db, gw = ops.rowdot_bwd_h2(h, w, g.contiguous(), call.cs[i:i + 1], sl.max_word(i), True,
                                                   ctx.needs_input_grad[1])
                sr.adopt(db, gw)
            call.grad = (planes, db)
            return h.new_empty(h.shape), gw, None   # not materialised: the tower's backward reads call.grad
        gh, gw = ops.rowdot_bwd(h, w, g.contiguous(), ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return gh, gw, None


def fm_head_enabled() -> bool:
    import os
    return os.environ.get("PTREC_FM_HEAD", "1") != "0"


def fm_head(v: Tensor, w1, x, wd, bias, want_deep_in: bool, tower_units: int = 0, tower=None):
    """``v [B, F, D]``, ``w1 [B, F, 1]`` or None, ``x [B, nd]`` or None, ``wd`` = Linear(nd, 1).weight or None,
    ``bias`` scalar parameter or None -> ``(logit [B], deep_in [B, F*D + nd] or None)``; None if K8 does not cover
    the shape (the caller composes K3 + library ops instead).  ``tower_units``: width of the Dense layer that will
    consume ``deep_in``; when that layer runs on K6 the head also writes its bf16 operand planes
    (``deep_in._ptrec_planes``), sparing the layer its split pass.  ``tower``: the ``MLP`` that is the ONLY consumer of
    ``deep_in``; when its next forward runs fused (K6 fused tower) the head writes the tower's fp16 operand planes and
    no fp32 ``deep_in`` at all (the returned tensor then only carries the autograd edge to that MLP)."""
    B, F, D = v.shape
    nd = 0 if x is None else x.shape[1]
    if not (v.is_cuda and v.dtype == torch.float32 and fm_head_enabled() and ops.fm_head_supported(F, D, nd)):
        return None
    v2d = v.reshape(B, F * D)
    if not (v2d.stride(1) == 1 and v2d.stride(0) % 4 == 0 and v2d.data_ptr() % 16 == 0):
        v2d = v2d.contiguous()
    from . import dense
    nd_ = 0 if x is None else x.shape[1]
    call = None
    if want_deep_in and tower is not None and hasattr(tower, "tower_call"):
        call = tower.tower_call(_Meta(B, F * D + nd_, v))
    want_planes = (call is None and want_deep_in and tower_units > 0 and dense.tc_linear_enabled() and ops.tc_mode() == "bf16x3"
                   and B * (F * D + nd) * tower_units >= dense.TC_MIN_MACS)
    logit, deep_in, planes = _FMHead.apply(v2d, w1.reshape(B, F).contiguous() if w1 is not None else None,
                                           x.contiguous() if x is not None else None,
                                           wd.reshape(-1) if (wd is not None and x is not None) else None,
                                           bias.reshape(1) if bias is not None else None, F, D, want_deep_in,
                                           want_planes, call)
    if not want_deep_in:
        return logit, None
    if call is not None:
        deep_in._ptrec_tower_in = call
        return logit, deep_in
    if planes.numel():
        deep_in._ptrec_planes = planes
    return logit, deep_in


class _Meta:
    """Shape / dtype / device of a tensor that does not exist yet (what ``MLP.tower_call`` needs to decide)."""

    def __init__(self, rows: int, cols: int, like: Tensor):
        self.shape, self.dtype, self.device, self.is_cuda = (rows, cols), like.dtype, like.device, like.is_cuda

    def dim(self):
        return 2


def row_dot(h: Tensor, weight: Tensor, tower_handoff: bool = False):
    """``Linear(H, 1, bias=False)(h).squeeze(-1)`` as one pass; None if the shape is not covered.
    ``tower_handoff``: the caller guarantees that ``h`` — the output of an ``MLP`` — has no other consumer; when that
    MLP ran as a K6 fused tower, the backward then hands the gradient over in the tower's operand format (no fp32
    gradient of ``h`` is written)."""
    if not (h.is_cuda and h.dtype == torch.float32 and h.dim() == 2 and fm_head_enabled()
            and ops.rowdot_supported(h.shape[1]) and h.stride(1) == 1 and h.stride(0) % 4 == 0
            and h.data_ptr() % 16 == 0):
        return None
    call = getattr(h, "_ptrec_tower_out", None) if tower_handoff else None
    wv = weight.reshape(-1)
    if wv.data_ptr() % 16 != 0 or wv.stride(0) != 1:   # e.g. a column slice of a wider Linear: the kernels read 16 bytes at a time
        wv = wv.clone()
    return _RowDot.apply(h, wv, call)


class CrossNet(nn.Module):
    """DCN-v2 cross network: ``x_{l+1} = x0 * (x_l W_l^T + b_l) + x_l`` for ``l < num_layers``.
    Parameters are ``layers.{l}.weight [d, d]`` / ``layers.{l}.bias [d]`` (``nn.Linear`` modules, so the
    reference's ``_reset_weights_fn`` initialises them).  The contraction runs in bf16 on the tcgen05 tensor
    cores with fp32 accumulation and a fused ``x0 * u + x`` epilogue; fp32 master weights."""

    def __init__(self, input_units: int, num_layers: int):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(input_units, input_units) for _ in range(num_layers)])

    def forward(self, x0: Tensor) -> Tensor:
        return ops.cross_net(x0, [m.weight for m in self.layers], [m.bias for m in self.layers])

    def forward_head(self, x0: Tensor, head_w: Tensor) -> Tensor:
        """``forward(x0) @ head_w`` ([B]): the cross half of a closing ``Linear(concat(cross, deep), 1)`` fused with the
        last cross layer's bf16 output (no fp32 [B, d] tensor either way)."""
        return ops.cross_net_head(x0, [m.weight for m in self.layers], [m.bias for m in self.layers], head_w)


class AttentionPooling(nn.Module):
    """DIN attention pooling: ``pooled = sum_{l < len} a_l k_l`` with the activation unit
    ``a_l = fc3(relu(fc2(relu(fc1([q, k_l, q - k_l, q * k_l])))))`` (no softmax).  Parameters are three
    ``nn.Linear`` modules (``fc1: 4*dim -> hidden[0]``, ``fc2``, ``fc3: hidden[1] -> 1``) so the reference's init
    applies; forward and backward are one fused CUDA kernel each."""

    def __init__(self, dim: int, hidden=(80, 40)):
        super().__init__()
        self.dim = dim
        self.fc1 = nn.Linear(4 * dim, hidden[0])
        self.fc2 = nn.Linear(hidden[0], hidden[1])
        self.fc3 = nn.Linear(hidden[1], 1)

    def forward(self, q: Tensor, keys: Tensor, lens: Tensor = None) -> Tensor:
        return ops.din_attn_pool(q, keys, lens, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                                 self.fc3.weight, self.fc3.bias)
