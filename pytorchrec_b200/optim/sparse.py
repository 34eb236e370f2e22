"""Sparse optimizers whose embedding-table update is fused into the backward scatter.

The reference updates tables densely: ``torch.optim.SGD/Adam`` or its own ``AdamW`` sweep every row
of every table each step (torchrec/optim/optimizers.py:7-11, AdamW.py:21-61, driven from
IModel.py:122-124).  Here the tables' rows are updated inside ``loss.backward()`` by
``ptrec_embedding_bwd_fused_*`` — only rows touched by the batch are read or written — while the
dense tower parameters of the same param groups go through the stock torch optimizer of the same
family.  Constructor convention is the reference's: ``Optimizer(params=model.get_parameters(), **kw)``
(RepeatTask.py:96).

Semantics: SGD and Adagrad equal their dense torch counterparts on touched rows and
leave untouched rows alone, which is what the dense update does for a zero gradient when
``weight_decay == 0``.  ``weight_decay`` on tables applies to touched rows only.  ``SparseAdam`` is
*lazy* Adam with the arithmetic of ``torch.optim.SparseAdam``.
"""
import ctypes
from typing import Dict, Optional

import torch
from torch.optim import Optimizer

from .. import _lib
from .._lib import OptimArgs


class _FusedSparseOptimizer(Optimizer):
    KIND = -1
    N_STATE = 0

    def __init__(self, params, defaults, dense_cls, dense_kwargs):
        super().__init__(params, defaults)
        self._fused = set()
        for group in self.param_groups:
            for p in group["params"]:
                if getattr(p, "_ptrec_table", None) is not None:
                    p._ptrec_optim = (self, group)
                    self._fused.add(p)
        dense_groups = []
        self._dense_links = []
        for group in self.param_groups:
            dense_params = [p for p in group["params"] if p not in self._fused]
            if dense_params:
                g = {k: v for k, v in group.items() if k != "params" and k in dense_kwargs}
                g["params"] = dense_params
                dense_groups.append(g)
                self._dense_links.append((group, len(dense_groups) - 1))
        # The companion is built at the first step(): stock optimizers such as torch.optim.Adagrad allocate
        # their state in __init__, i.e. before IModel.compile() moves the model to its device.
        self._dense = None
        self._dense_spec = (dense_cls, dense_groups, dense_kwargs) if dense_groups else None
        self._step_count_fused = 0
        self._ptr_cache: Dict[int, tuple] = {}
        self._interleaved: Dict[int, torch.Tensor] = {}  # id(param) -> [rows, stride] buffer holding weight | state

    def graph_safe(self) -> bool:
        """True when a captured step stays valid on replay (no host-side step count in the arithmetic)."""
        return False

    def _interleave(self, p: torch.Tensor, names, fills, slots: int) -> None:
        """Re-house table ``p`` and its element-wise state in ONE ``[rows, slots*D]`` buffer (weight in columns
        [0, D), state k in [k*D, (k+1)*D)).  On B200 a random row access costs a full 128-byte DRAM line whatever the
        row size (profiles/r1_ubench_dram_granularity.csv), so keeping a 64-byte weight row and its 64-byte state row
        in the same line halves the row traffic of the fused update.  ``p.data`` becomes a strided view of the
        buffer; ``state[p][name]`` are views too, so state_dict / load_state_dict keep working."""
        st = self.state[p]
        D = p.shape[1]
        buf = self._interleaved.get(id(p))
        ok = (buf is not None and buf.device == p.device and p.data.data_ptr() == buf.data_ptr()
              and all(n in st and st[n].data_ptr() == buf.data_ptr() + 4 * D * (k + 1) for k, n in enumerate(names)))
        if ok:
            return
        alloc = getattr(p, "_ptrec_alloc", None)  # e.g. symmetric memory for tables that peers read over NVLink
        new = (alloc(p.shape[0], slots * D) if alloc is not None
               else torch.empty(p.shape[0], slots * D, dtype=torch.float32, device=p.device))
        new[:, :D].copy_(p.data)
        for k, (n, fill) in enumerate(zip(names, fills)):
            view = new[:, (k + 1) * D:(k + 2) * D]
            if n in st and st[n].shape == view.shape:
                view.copy_(st[n].to(p.device))
            else:
                view.fill_(fill)
            st[n] = view
        if slots > len(names) + 1:
            new[:, (len(names) + 1) * D:].zero_()
        p.data = new[:, :D]
        self._interleaved[id(p)] = new
        self._ptr_cache.clear()

    # ---- called by EmbeddingGroup.apply_backward -------------------------------------------------
    def _state_tensors(self, p: torch.Tensor):
        """(state1, state2) for table ``p`` — allocated on first use on the table's device."""
        raise NotImplementedError

    def _optim_args(self, group: dict) -> OptimArgs:
        raise NotImplementedError

    def _fused_prepare(self, emb_group, group: dict):
        weights = [t.weight for t in emb_group.tables]
        key = tuple(w.data_ptr() for w in weights)
        cached = self._ptr_cache.get(id(emb_group))
        if cached is None or cached[0] != key:
            s1, s2 = [], []
            for w in weights:
                a, b = self._state_tensors(w)
                s1.append(a)
                s2.append(b)
            dev = weights[0].device
            key = tuple(w.data_ptr() for w in weights)  # _state_tensors may have re-housed the tables
            p1 = torch.tensor([t.data_ptr() for t in s1], dtype=torch.int64).to(dev) if self.N_STATE >= 1 else None
            p2 = torch.tensor([t.data_ptr() for t in s2], dtype=torch.int64).to(dev) if self.N_STATE >= 2 else None
            cached = (key, p1, p2)
            self._ptr_cache[id(emb_group)] = cached
        return cached[1], cached[2], self._optim_args(group)

    # ---- torch.optim.Optimizer surface ------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._step_count_fused += 1
        self._ensure_dense()
        if self._dense is not None:
            for group, di in self._dense_links:
                dg = self._dense.param_groups[di]
                for k in dg:
                    if k != "params" and k in group:
                        dg[k] = group[k]
            self._dense.step()
        return loss

    def _ensure_dense(self):
        if self._dense is None and self._dense_spec is not None:
            cls, groups, kwargs = self._dense_spec
            self._dense = cls(groups, **kwargs)

    def state_dict(self):
        self._ensure_dense()
        return {"fused": super().state_dict(), "dense": self._dense.state_dict() if self._dense else None,
                "step": self._step_count_fused}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict["fused"])
        self._ensure_dense()
        if self._dense is not None and state_dict.get("dense") is not None:
            self._dense.load_state_dict(state_dict["dense"])
        self._step_count_fused = int(state_dict.get("step", 0))
        self._ptr_cache.clear()


class SparseSGD(_FusedSparseOptimizer):
    """w[r] -= lr * g[r] on touched rows (``torch.optim.SGD`` without momentum for dense params)."""
    KIND = _lib.OPT_SGD
    N_STATE = 0

    def __init__(self, params, lr=1e-3, weight_decay=0.0):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        super().__init__(params, dict(lr=lr, weight_decay=weight_decay), torch.optim.SGD,
                         dict(lr=lr, weight_decay=weight_decay))

    def graph_safe(self):
        return True

    def _state_tensors(self, p):
        return None, None

    def _optim_args(self, group):
        return OptimArgs(kind=self.KIND, step=self._step_count_fused + 1, lr=group["lr"], eps=0.0, beta1=0.0,
                         beta2=0.0, weight_decay=group["weight_decay"], lr_decay=0.0)


class SparseAdagrad(_FusedSparseOptimizer):
    """Element-wise Adagrad with ``torch.optim.Adagrad`` arithmetic on touched rows:
    ``sum += g*g; w -= clr * g / (sqrt(sum) + eps)``, ``clr = lr / (1 + (step-1)*lr_decay)``."""
    KIND = _lib.OPT_ADAGRAD
    N_STATE = 1

    def __init__(self, params, lr=1e-2, lr_decay=0.0, weight_decay=0.0, initial_accumulator_value=0.0, eps=1e-10,
                 interleave: bool = True):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        defaults = dict(lr=lr, lr_decay=lr_decay, weight_decay=weight_decay,
                        initial_accumulator_value=initial_accumulator_value, eps=eps)
        self.interleave = interleave
        super().__init__(params, defaults, torch.optim.Adagrad, dict(defaults))

    def graph_safe(self):
        return all(g["lr_decay"] == 0 for g in self.param_groups)

    def _init_acc(self, p):
        for g in self.param_groups:
            if any(q is p for q in g["params"]):
                return g["initial_accumulator_value"]
        return 0.0

    def _state_tensors(self, p):
        st = self.state[p]
        if self.interleave and type(self) is SparseAdagrad:
            self._interleave(p, ["sum"], [self._init_acc(p)], slots=2)
        elif "sum" not in st or st["sum"].device != p.device:
            st["sum"] = torch.full_like(p.data, self._init_acc(p)) if "sum" not in st else st["sum"].to(p.device)
        return st["sum"], None

    def _optim_args(self, group):
        return OptimArgs(kind=self.KIND, step=self._step_count_fused + 1, lr=group["lr"], eps=group["eps"],
                         beta1=0.0, beta2=0.0, weight_decay=group["weight_decay"], lr_decay=group["lr_decay"])


class SparseRowWiseAdagrad(SparseAdagrad):
    """One accumulator per row: ``sum[r] += mean_k g[r,k]^2``.  State is ``[rows]`` instead of ``[rows, D]``."""
    KIND = _lib.OPT_ROWWISE_ADAGRAD
    N_STATE = 1

    def _state_tensors(self, p):
        st = self.state[p]
        if "sum" not in st or st["sum"].device != p.device:
            init = 0.0
            for g in self.param_groups:
                if any(q is p for q in g["params"]):
                    init = g["initial_accumulator_value"]
            st["sum"] = (torch.full((p.shape[0],), init, dtype=torch.float32, device=p.device)
                         if "sum" not in st else st["sum"].to(p.device))
        return st["sum"], None


class SparseAdam(_FusedSparseOptimizer):
    """Lazy Adam: moments and weights move only on rows touched by the batch
    (``torch.optim.SparseAdam`` arithmetic); dense parameters use ``torch.optim.Adam``."""
    KIND = _lib.OPT_LAZY_ADAM
    N_STATE = 2

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, interleave: bool = True):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid betas: {betas}")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.interleave = interleave
        super().__init__(params, defaults, torch.optim.Adam, dict(defaults))

    def _state_tensors(self, p):
        st = self.state[p]
        if self.interleave:
            self._interleave(p, ["exp_avg", "exp_avg_sq"], [0.0, 0.0], slots=4)  # w | m | v | pad: 2 lines at D=16
            return st["exp_avg"], st["exp_avg_sq"]
        for k in ("exp_avg", "exp_avg_sq"):
            if k not in st:
                st[k] = torch.zeros_like(p.data)
            elif st[k].device != p.device:
                st[k] = st[k].to(p.device)
        return st["exp_avg"], st["exp_avg_sq"]

    def _optim_args(self, group):
        return OptimArgs(kind=self.KIND, step=self._step_count_fused + 1, lr=group["lr"], eps=group["eps"],
                         beta1=group["betas"][0], beta2=group["betas"][1], weight_decay=group["weight_decay"],
                         lr_decay=0.0)
