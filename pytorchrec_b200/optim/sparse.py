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
        self._dense_plan: Dict[tuple, tuple] = {}        # (dense group, step bucket) -> (pointer key, descriptors, chunk starts, n)
        self._retired_plans: list = []                   # plans superseded during a graph capture (their pinned buffers outlive it)
        # set by data-parallel models whose replicas exchange over NVLink peer memory (distributed/sharded.py):
        # object with .world, .stage(key, n_floats, device) -> (stage, peer pointer array), .barrier(), .fallback(params)
        self._peer_reduce = None

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
        # (no cache flush here: the pointer arrays cached per embedding group are keyed by the tables' addresses, so
        # the group this table belongs to rebuilds its own entry; flushing everything made the FIRST group of a model
        # with two widths rebuild its arrays one step later — inside the CUDA-graph capture when warmup == 1)

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
            if not self._fused_dense_step():
                if self._peer_reduce is not None:  # K7 not applicable: average the gradients with the library collective
                    self._peer_reduce.fallback([p for g in self._dense.param_groups for p in g["params"]])
                self._dense.step()
        return loss

    # ---- K7: the dense companion's arithmetic in one launch per parameter group ------------------------
    _DENSE_KIND = {torch.optim.SGD: _lib.OPT_SGD, torch.optim.Adagrad: _lib.OPT_ADAGRAD,
                   torch.optim.Adam: _lib.OPT_LAZY_ADAM}

    def _fused_dense_step(self) -> bool:
        """Step the dense parameters with ``ptrec_dense_optim_step`` (state lives in the torch optimizer object, so
        ``state_dict`` / ``load_state_dict`` are unchanged).  False = not applicable (CPU tensors, sparse or missing
        layouts): the caller runs the stock ``torch.optim`` step instead."""
        import os
        kind = self._DENSE_KIND.get(type(self._dense))
        if kind is None or os.environ.get("PTREC_FUSED_DENSE_OPTIM", "1") == "0":
            return False
        plans = []
        for gi, dg in enumerate(self._dense.param_groups):
            if dg.get("maximize") or dg.get("amsgrad") or dg.get("momentum") or dg.get("nesterov") \
                    or dg.get("differentiable") or dg.get("capturable"):
                return False
            ps = [p for p in dg["params"] if p.grad is not None]
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and not p.grad.is_sparse
                        and p.grad.is_contiguous() and p.grad.dtype == torch.float32):
                    return False
            plans.append((gi, dg, ps))
        lib = _lib.load()
        chunk = lib.ptrec_dense_optim_chunk()
        # phase 1 (host only, may still return False): states, step counters, descriptor tables
        staged = []
        for gi, dg, ps in plans:
            if not ps:
                continue
            states = []
            for p in ps:
                if kind == _lib.OPT_SGD:  # no state; do not create empty entries in the defaultdict
                    states.append(None)
                    continue
                st = self._dense.state[p]
                if kind == _lib.OPT_LAZY_ADAM and "exp_avg" not in st:  # torch.optim.Adam initialises lazily
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if kind != _lib.OPT_SGD:
                    if not torch.is_tensor(st.get("step")):
                        return False
                    for k in (("sum",) if kind == _lib.OPT_ADAGRAD else ("exp_avg", "exp_avg_sq")):
                        t = st[k]
                        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                            return False
                states.append(st)
            staged.append((gi, dg, ps, states))
        launches = []
        # Parameters are bucketed by what the kernel needs to be told once per launch: the hyper-parameters of their
        # group and their step counter (torch.optim keeps one per parameter; a parameter that skipped a step lags
        # behind).  Groups with identical hyper-parameters — the usual case: the reference's get_parameters() returns
        # several groups that differ in nothing the dense optimizer uses — share ONE launch.
        merged: Dict[tuple, list] = {}
        for gi, dg, ps, states in staged:
            betas = dg.get("betas", (0.0, 0.0))
            hyper = (float(dg["lr"]), float(dg.get("eps", 0.0)), float(betas[0]), float(betas[1]),
                     float(dg.get("weight_decay", 0.0)), float(dg.get("lr_decay", 0.0)))
            for i, st in enumerate(states):
                step = 1
                if kind != _lib.OPT_SGD:
                    st["step"] += 1
                    step = int(st["step"].item())
                merged.setdefault((hyper, step), []).append((ps[i], st))
        for bi, ((hyper, step), items) in enumerate(sorted(merged.items(), key=lambda kv: kv[0])):
            gi = "merged"
            dg = {"lr": hyper[0], "eps": hyper[1], "betas": (hyper[2], hyper[3]), "weight_decay": hyper[4],
                  "lr_decay": hyper[5]}
            bps = [p for p, _ in items]
            if kind == _lib.OPT_ADAGRAD:
                s1, s2 = [st["sum"] for _, st in items], [None] * len(items)
            elif kind == _lib.OPT_LAZY_ADAM:
                s1, s2 = [st["exp_avg"] for _, st in items], [st["exp_avg_sq"] for _, st in items]
            else:
                s1 = s2 = [None] * len(items)
            key = tuple((p.data_ptr(), p.grad.data_ptr(), a.data_ptr() if a is not None else 0,
                         b.data_ptr() if b is not None else 0) for p, a, b in zip(bps, s1, s2))
            cached = self._dense_plan.get((gi, bi))
            if cached is None or cached[0] != key:
                import numpy as np
                rec = np.zeros((len(bps), 5), dtype=np.int64)
                starts = np.zeros(len(bps) + 1, dtype=np.int32)
                for i, (p, k4) in enumerate(zip(bps, key)):
                    rec[i, :4] = k4
                    rec[i, 4] = p.numel()
                    starts[i + 1] = starts[i] + (p.numel() + chunk - 1) // chunk
                dev = bps[0].device
                # pinned staging: this may run inside a CUDA-graph capture (the captured copy node re-reads the
                # pinned buffers on replay, so they are kept alive with the plan)
                h_rec, h_starts = torch.from_numpy(rec).pin_memory(), torch.from_numpy(starts).pin_memory()
                old = self._dense_plan.get((gi, bi))
                if old is not None and torch.cuda.is_current_stream_capturing():
                    # freeing a pinned buffer makes the host allocator record an event on every stream that used it; if
                    # that stream is part of the running capture (the dense step may run on the weight-gradient side
                    # stream) the event is a captured one and the allocator's later cudaEventQuery fails
                    # (cudaErrorInvalidValue at the next pin_memory()): superseded plans outlive the capture
                    self._retired_plans.append(old)
                cached = (key, h_rec.to(dev, non_blocking=True), h_starts.to(dev, non_blocking=True),
                          int(starts[-1]), h_rec, h_starts)
                self._dense_plan[(gi, bi)] = cached
            betas = dg.get("betas", (0.0, 0.0))
            args = OptimArgs(kind=kind, step=step, lr=dg["lr"], eps=dg.get("eps", 0.0), beta1=betas[0],
                             beta2=betas[1], weight_decay=dg.get("weight_decay", 0.0),
                             lr_decay=dg.get("lr_decay", 0.0))
            launches.append(((gi, bi), cached, len(bps), args, bps[0].device))
        # phase 2: launches.  With a peer reducer attached (row-wise sharded models on NVLink peer memory) the
        # data-parallel mean of the gradients is fused in: pack every group's gradients into this rank's symmetric
        # stage, ONE barrier, then each K7 launch sums the ranks' stages itself (csrc/peer_sync.cu).
        red = self._peer_reduce
        if red is not None:
            from .. import ops
            stages = []
            for pkey, cached, n, args, dev in launches:
                stage, peer_ptrs = red.stage(pkey, cached[3] * chunk, dev)
                ops.dense_pack(cached[1], cached[2], n, cached[3], stage)
                stages.append(peer_ptrs)
            red.barrier()
        for li, (pkey, cached, n, args, dev) in enumerate(launches):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if red is not None:
                _lib.check(lib.ptrec_dense_optim_step_reduce(
                    ctypes.c_void_p(cached[1].data_ptr()), ctypes.c_void_p(cached[2].data_ptr()), n, cached[3],
                    ctypes.byref(args), ctypes.c_void_p(stages[li].data_ptr()), red.world, 1.0 / red.world, stream),
                    "ptrec_dense_optim_step_reduce")
            else:
                _lib.check(lib.ptrec_dense_optim_step(ctypes.c_void_p(cached[1].data_ptr()),
                                                      ctypes.c_void_p(cached[2].data_ptr()), n, cached[3],
                                                      ctypes.byref(args), stream), "ptrec_dense_optim_step")
        return True

    # ---- per-table optimizer state (N4: sharded checkpoints gather / scatter it row-wise) -----------------------
    def table_state(self, p: torch.Tensor) -> Dict[str, torch.Tensor]:
        """The state tensors of fused table ``p`` by name (``sum`` / ``exp_avg`` / ``exp_avg_sq``), created if the
        first step has not happened yet; views into the interleaved buffer where there is one."""
        if p not in self._fused:
            raise KeyError("not a table owned by this fused optimizer")
        if self.N_STATE:
            self._state_tensors(p)
        return {k: v for k, v in self.state[p].items() if torch.is_tensor(v)}

    @torch.no_grad()
    def load_table_state(self, p: torch.Tensor, state: Dict[str, torch.Tensor]) -> None:
        cur = self.table_state(p)
        for k, v in state.items():
            if k not in cur or cur[k].shape != v.shape:
                raise ValueError(f"optimizer state {k!r}: expected shape {tuple(cur[k].shape) if k in cur else None}, "
                                 f"got {tuple(v.shape)}")
            cur[k].copy_(v)

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
        if self.interleave and type(self) is SparseAdagrad and p.dtype == torch.float32:
            self._interleave(p, ["sum"], [self._init_acc(p)], slots=2)
        elif "sum" not in st or st["sum"].device != p.device:
            # bf16 tables: the accumulator stays fp32, in a tensor of its own with the weight's row stride in elements
            st["sum"] = (torch.full(p.shape, self._init_acc(p), dtype=torch.float32, device=p.device)
                         if "sum" not in st else st["sum"].to(p.device))
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
        if self.interleave and p.dtype == torch.float32:
            self._interleave(p, ["exp_avg", "exp_avg_sq"], [0.0, 0.0], slots=4)  # w | m | v | pad: 2 lines at D=16
            return st["exp_avg"], st["exp_avg_sq"]
        for k in ("exp_avg", "exp_avg_sq"):
            if k not in st:
                st[k] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)  # fp32 moments for bf16 tables too
            elif st[k].device != p.device:
                st[k] = st[k].to(p.device)
        return st["exp_avg"], st["exp_avg_sq"]

    def _optim_args(self, group):
        return OptimArgs(kind=self.KIND, step=self._step_count_fused + 1, lr=group["lr"], eps=group["eps"],
                         beta1=group["betas"][0], beta2=group["betas"][1], weight_decay=group["weight_decay"],
                         lr_decay=0.0)
