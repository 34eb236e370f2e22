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
from typing import List, Optional

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


class _TowerScales:
    """Carried power-of-two scales of one MLP on the fused K6 path (include/ptrec_b200.h, "the fused tower"): one slot
    {scale, max since the last roll} per tensor — input, every weight, every hidden activation, every pre-activation
    gradient.  The slots are seeded with measured maxima by one forward / backward on the per-layer path
    (``MLP._forward_recording``); from then on every kernel that writes a tensor's planes raises the slot's maximum and
    ``roll`` (once per forward) turns it into the next step's scale."""

    def __init__(self, n_layers: int, device: torch.device):
        self.L = n_layers
        self.device = device
        self.slots = torch.zeros(3 * n_layers, 2, dtype=torch.float32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.fwd_ready = False
        self.bwd_ready = False

    # slot indices
    def i_x(self): return 0
    def i_w(self, l): return 1 + l
    def i_y(self, l): return self.L + 1 + l        # output of hidden layer l (l < L - 1) = input of layer l + 1
    def i_g(self, l): return 2 * self.L + l        # gradient w.r.t. the pre-activation of layer l

    def i_in(self, l): return self.i_x() if l == 0 else self.i_y(l - 1)

    def seed(self, i: int, t: torch.Tensor) -> None:
        m = self.slots[i, 1]
        torch.maximum(m, t.detach().abs().max().to(torch.float32), out=m)

    def max_word(self, i: int) -> torch.Tensor:
        return self.slots[i, 1:2]

    def roll(self) -> torch.Tensor:
        cs = ops.tc_scale_roll(self.slots, self.err)
        self.err_host.copy_(self.err, non_blocking=True)  # polled without synchronising (MLP.poll_errors)
        return cs


_W_STREAMS = {}   # device -> side stream of the ahead-of-time weight splits
_G_STREAMS = {}   # device -> side stream of the tower's weight-gradient GEMMs
_DEFER_JOIN = [False]   # True inside IModel's train step: the join happens once, before the optimizer step


class side_reductions:
    """``with side_reductions(device) as sr:`` around ONE library call that ends in a gradient-finalising reduction (bias /
    weight-vector gradients from per-CTA partials: ``tc_gemm_split2h_fused(want_colsum=True)``, ``rowdot_bwd_h2``,
    ``fm_head_bwd``): inside an ``IModel`` train step the reduction runs on the weight-gradient side stream behind an event
    on the producer's stream (``ptrec_set_reduce_stream``) and leaves the backward's critical path; ``sr.adopt(t, ...)``
    for the tensors it writes.  Every call in flight needs its own partial buffer (``ws_tag``).  Elsewhere: a no-op.
    ``PTREC_REDUCE_STREAM=0`` switches it off."""

    def __init__(self, device):
        self.device, self.on, self.side = device, False, None

    def __enter__(self):
        import os
        dev = self.device
        if (_DEFER_JOIN[0] and dev is not None and torch.device(dev).type == "cuda"
                and os.environ.get("PTREC_REDUCE_STREAM", "1") != "0"):
            side = _wgrad_stream(dev)
            if side is not None:
                from ... import _lib
                from .embedding import register_join_stream
                register_join_stream(dev, side)
                _lib.load().ptrec_set_reduce_stream(side.cuda_stream)
                self.on, self.side = True, side
        return self

    def __exit__(self, *exc):
        if self.on:
            from ... import _lib
            _lib.load().ptrec_set_reduce_stream(None)
        return False

    def adopt(self, *tensors) -> None:
        if self.on:
            for t in tensors:
                if t is not None:
                    t.record_stream(self.side)


def _wgrad_stream(dev):
    """The fused tower's weight-gradient GEMMs (and their split-K reductions) feed nothing but the optimizer, so they run
    on a side stream beside the input-gradient chain, whose result the rest of the backward (head, embedding update)
    waits for; ``join_aux_streams`` (IModel, between backward and the optimizer step) joins it.  ``PTREC_WGRAD_STREAM=0``
    keeps them on the main stream."""
    import os
    if os.environ.get("PTREC_WGRAD_STREAM", "1") == "0":
        return None
    s = _G_STREAMS.get(dev)
    if s is None:
        s = _G_STREAMS[dev] = torch.cuda.Stream(dev)
    return s


class _TowerCall:
    """One forward (+ backward) of a fused tower: the scales it reads, and what its neighbours hand over in the tower's
    own operand format — the producer of its input (``fm_head(..., tower=mlp)``: planes of x) and the consumer of its
    output (``row_dot(..., tower_handoff=True)``: planes of the masked output gradient + the last bias gradient)."""

    def __init__(self, mlp, cs):
        self.mlp, self.cs = mlp, cs
        self.in_planes = None      # planes of the tower input, written by its producer
        self.grad = None           # (planes of the last pre-activation gradient, last bias gradient)
        self.w_planes = None       # [(planes of W_l, transposed planes or None)] when prepared ahead (MLP.tower_call)
        self.w_event = None        # ... on a side stream: recorded when they are complete


class _TcMLP(torch.autograd.Function):
    """The whole Linear -> ReLU stack on K6 with nothing between two GEMMs: the epilogue of layer l writes relu(x W^T + b)
    as the fp16 planes layer l + 1 consumes (and its weight gradient re-reads MN-major) plus one bit per element (> 0);
    the input-gradient GEMM of layer l + 1 applies that bit mask, writes the planes of layer l's pre-activation gradient
    and its column sums (bias gradient).  fp32 is materialised only for the tower's output and its input gradient."""

    @staticmethod
    def forward(ctx, x, mlp, call, *params):
        sl: _TowerScales = mlp._scales
        L = sl.L
        if call is None:
            call = _TowerCall(mlp, sl.roll())
        cs = call.cs
        sc = lambda i: cs[i:i + 1]
        need_dx = ctx.needs_input_grad[0]
        any_grad = any(ctx.needs_input_grad)
        px = call.in_planes   # the producer of x wrote them (x itself is then not materialised)
        if px is None:
            px, _, _ = ops.tc_split2h_prescaled(x, sc(sl.i_x()), sl.max_word(sl.i_x()))
        acts, masks, pwts = [px], [], []
        y = None
        if call.w_planes is not None:   # split ahead of time beside the producer of x (MLP.tower_call)
            main = torch.cuda.current_stream(x.device)
            main.wait_event(call.w_event)
            for pw, pwt in call.w_planes:
                pw.record_stream(main)
                if pwt is not None:
                    pwt.record_stream(main)
        for l in range(L):
            W, b = params[2 * l], params[2 * l + 1]
            if call.w_planes is not None:
                pw, pwt = call.w_planes[l]
            else:
                pw, pwt, _ = ops.tc_split2h_prescaled(W, sc(sl.i_w(l)), sl.max_word(sl.i_w(l)),
                                                      want_t=any_grad and (l > 0 or need_dx))
            pwts.append(pwt)
            last = l == L - 1
            y, py, mask, _ = ops.tc_gemm_split2h_fused(
                acts[l], sc(sl.i_in(l)), pw, sc(sl.i_w(l)), W.shape[1], bias=b, relu=True, want_out=last,
                out_scale=None if last else sc(sl.i_y(l)), want_mask=any_grad and not last,
                max_out=None if last else sl.max_word(sl.i_y(l)))
            if not last:
                acts.append(py)
                masks.append(mask)
        if getattr(mlp, "_keep_masks", False):  # test introspection: the ReLU decisions this forward took
            mlp._last_masks = list(masks)
        ctx.scales = sl
        ctx.call = call
        ctx.params = params if any_grad else None
        ctx.dims = [(params[2 * l].shape[0], params[2 * l].shape[1]) for l in range(L)]
        ctx.has_bias = [params[2 * l + 1] is not None for l in range(L)]
        if any_grad:
            ctx.n_masks = len(masks)
            ctx.save_for_backward(y, cs, *acts, *masks, *[t for t in pwts if t is not None])
            ctx.pwt_present = [t is not None for t in pwts]
        return y

    @staticmethod
    def backward(ctx, gy):
        sl: _TowerScales = ctx.scales
        L = sl.L
        saved = ctx.saved_tensors
        y, cs = saved[0], saved[1]
        acts = list(saved[2:2 + L])
        masks = list(saved[2 + L:2 + L + ctx.n_masks])
        rest = list(saved[2 + L + ctx.n_masks:])
        pwts = [rest.pop(0) if present else None for present in ctx.pwt_present]
        sc = lambda i: cs[i:i + 1]
        need = ctx.needs_input_grad
        if gy.stride(-1) != 1:
            gy = gy.contiguous()
        grads = [None] * (2 * L)
        if ctx.call.grad is not None:   # the consumer of y handed its gradient over as planes (gy is not materialised)
            pg, db = ctx.call.grad
            ctx.call.grad = None
        else:
            pg, _, db = ops.tc_split2h_prescaled(gy, sc(sl.i_g(L - 1)), sl.max_word(sl.i_g(L - 1)), relu_ref=y,
                                                 want_colsum=ctx.has_bias[L - 1] and need[3 + 2 * (L - 1) + 1])
        dx = None
        # weight gradients on a side stream — only when autograd will adopt them as fresh .grad tensors (an existing
        # .grad would be accumulated into on the main stream, without waiting for the side stream)
        side = None
        if gy.is_cuda and all(p is None or p.grad is None for p in (ctx.params or ())):
            side = _wgrad_stream(gy.device)
        if side is not None:
            main = torch.cuda.current_stream(gy.device)
            if _DEFER_JOIN[0]:
                from .embedding import register_join_stream
                register_join_stream(gy.device, side)
        for l in range(L - 1, -1, -1):
            N, K = ctx.dims[l]
            pg_prev = db_prev = None
            if need[3 + 2 * l] and side is not None:   # g^T x  [N, K], forked before this layer's input-gradient GEMM
                side.wait_stream(main)       # pg of this layer (and, the first time, everything before the backward)
                with torch.cuda.stream(side):
                    dw = ops.tc_gemm_split2h_tn(pg, sc(sl.i_g(l)), N, acts[l], sc(sl.i_in(l)), K)
                    dw = dw if dw.is_contiguous() else dw.contiguous()
                pg.record_stream(side)
                acts[l].record_stream(side)
                cs.record_stream(side)
                dw.record_stream(main)       # consumed by the optimizer on the main stream after the join
                grads[2 * l] = dw
            if l > 0:    # g W, masked by the ReLU of layer l - 1, as planes + bias gradient of layer l - 1
                with side_reductions(gy.device) as sr:   # the column-sum reduce (bias gradient) leaves the chain
                    _, pg_prev, _, db_prev = ops.tc_gemm_split2h_fused(
                        pg, sc(sl.i_g(l)), pwts[l], sc(sl.i_w(l)), N, want_out=False, out_scale=sc(sl.i_g(l - 1)),
                        mask_in=masks[l - 1], want_colsum=ctx.has_bias[l - 1] and need[3 + 2 * (l - 1) + 1],
                        max_out=sl.max_word(sl.i_g(l - 1)), ws_tag=str(l) if sr.on else "")
                    sr.adopt(db_prev)
            elif need[0]:
                dx, _, _, _ = ops.tc_gemm_split2h_fused(pg, sc(sl.i_g(0)), pwts[0], sc(sl.i_w(0)), N)
            if need[3 + 2 * l] and side is None:
                dw = ops.tc_gemm_split2h_tn(pg, sc(sl.i_g(l)), N, acts[l], sc(sl.i_in(l)), K)   # g^T x  [N, K]
                grads[2 * l] = dw if dw.is_contiguous() else dw.contiguous()
            grads[2 * l + 1] = db
            pg, db = pg_prev, db_prev
        if side is not None and not _DEFER_JOIN[0]:
            main.wait_stream(side)   # a caller that reads .grad right after backward() (no IModel step around it)
        sl.bwd_ready = True
        return (dx, None, None, *grads)


def tc_fused_enabled() -> bool:
    return os.environ.get("PTREC_TC_FUSED", "1") != "0"


class MLP(Module):
    def __init__(self, input_units: int, hidden_units_list: List[int], activation: str, dropout: float):
        super().__init__()
        self.mlp = Sequential()
        units = input_units
        for index, hidden_units in enumerate(hidden_units_list):
            self.mlp.add_module(f"dense_{index}", Dense(units, hidden_units, activation, dropout))
            units = hidden_units
        self._scales = None  # _TowerScales, created on the first CUDA forward that qualifies for the fused path

    def _fused_eligible(self, x) -> bool:
        if not (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and tc_linear_enabled() and tc_fused_enabled()
                and ops.tc_mode() == "fp16x2" and ops.tc_fused_supported() and not torch.is_autocast_enabled()
                and len(self.mlp) >= 1):
            return False
        for d in self.mlp:
            if not isinstance(d, Dense) or d.linear.weight.dtype != torch.float32 or d.linear.bias is None \
                    or (self.training and d.dropout.p > 0) or x.shape[0] * d.linear.weight.numel() < TC_MIN_MACS:
                return False
        return True

    def _forward_recording(self, x, want_grad: bool):
        """One pass on the per-layer path that seeds the carried scales with measured maxima (a hidden activation's
        gradient bounds its ReLU-masked version, which is what the slot describes)."""
        sl = self._scales
        L = sl.L
        sl.seed(sl.i_x(), x)
        h = x
        for l, d in enumerate(self.mlp):
            sl.seed(sl.i_w(l), d.linear.weight)
            h = d(h)
            if l < L - 1:
                sl.seed(sl.i_y(l), h)
            if want_grad and h.requires_grad:
                def hook(g, i=sl.i_g(l), first=(l == 0)):
                    sl.seed(i, g)
                    if first:
                        sl.bwd_ready = True
                h.register_hook(hook)
        sl.fwd_ready = True
        return h

    def tower_call(self, x_like) -> Optional["_TowerCall"]:
        """For the producer of this tower's input (``fm_head(..., tower=mlp)``): if the next ``forward`` will run fused
        — ``x_like`` has the input's shape / dtype / device — roll the scales now and return the call record, so the
        producer can write the input directly as planes (``call.cs[0:1]`` = its scale, ``call.mlp._scales.max_word(0)``
        = where its maximum goes, ``call.in_planes`` = the planes).  None otherwise: the producer writes fp32."""
        self._pending = None
        if not self._fused_eligible(x_like):
            return None
        sl = self._scales
        if sl is None or sl.device != x_like.device or sl.L != len(self.mlp):
            return None
        want_grad = torch.is_grad_enabled()
        if not sl.fwd_ready or (want_grad and not sl.bwd_ready):
            return None
        call = self._pending = _TowerCall(self, sl.roll())
        # the weights' planes depend on nothing the producer of x computes: split them on a side stream while it runs
        dev = x_like.device
        main = torch.cuda.current_stream(dev)
        side = _W_STREAMS.get(dev)
        if side is None:
            side = _W_STREAMS[dev] = torch.cuda.Stream(dev)
        side.wait_stream(main)   # the roll above (and the optimizer step that wrote the weights)
        with torch.cuda.stream(side):
            call.w_planes = []
            for l, d in enumerate(self.mlp):
                pw, pwt, _ = ops.tc_split2h_prescaled(d.linear.weight.detach(), call.cs[sl.i_w(l):sl.i_w(l) + 1],
                                                      sl.max_word(sl.i_w(l)), want_t=want_grad)
                call.w_planes.append((pw, pwt))
            call.w_event = torch.cuda.Event()
            call.w_event.record(side)
        return call

    def forward(self, x):
        call, self._pending = getattr(self, "_pending", None), None
        if call is not None and getattr(x, "_ptrec_tower_in", None) is not call:
            call = None        # someone else's tensor: a call record only pairs with the input written for it
        if call is None and getattr(x, "_ptrec_tower_in", None) is not None:
            raise RuntimeError("this tensor was written as the planes of another fused-tower call; its fp32 values do not exist")
        if call is None and not self._fused_eligible(x):
            return self.mlp(x)
        if self._scales is None or self._scales.device != x.device or self._scales.L != len(self.mlp):
            self._scales = _TowerScales(len(self.mlp), x.device)
        sl = self._scales
        want_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if call is None and (not sl.fwd_ready or (want_grad and not sl.bwd_ready)):
            return self._forward_recording(x, want_grad)
        x2 = x if x.stride(-1) == 1 else x.contiguous()
        params = []
        for d in self.mlp:
            params += [d.linear.weight, d.linear.bias]
        if call is None:
            call = _TowerCall(self, sl.roll())
        y = _TcMLP.apply(x2, self, call, *params)
        y._ptrec_tower_out = call   # row_dot(..., tower_handoff=True) hands the gradient of y back as planes
        return y

    # picked up by IModel.train_step (non-synchronising) and IModel._check_device_flags (once per epoch)
    def poll_errors(self) -> None:
        if self._scales is not None and int(self._scales.err_host[0]):
            self._raise_scale_error()

    def check_errors(self) -> None:
        if self._scales is not None and int(self._scales.err.item()):
            self._raise_scale_error()

    def reset_scales(self) -> None:
        """Forget the carried scales: the next forward / backward measures them again on the per-layer path."""
        self._scales = None

    def _raise_scale_error(self):
        raise RuntimeError(
            "K6 fused tower: a tensor grew more than 256-fold between two consecutive steps and left the fp16 range of "
            "its carried scale — results since then are invalid (the run is diverging, or the inputs changed scale "
            "abruptly).  MLP.reset_scales() re-measures; PTREC_TC_FUSED=0 selects the per-layer path, which derives "
            "every scale from the tensor itself")
