"""Training runtime with the reference's ``IModel`` surface (torchrec/model/IModel.py:34-321).

Lifecycle is the reference's: ``set_torch_seed`` -> ``Module.__init__`` -> ``_init_weights()``
(subclass builds its layers in declaration order) -> ``_reset_weights()`` re-draws N(0, 0.01) for
every Linear / Embedding weight and bias (IModel.py:37-71).  ``compile`` type-checks its arguments
the same way (:94-114) and ``train_step`` is the same five-line hot loop (:116-125).  What differs
is underneath: embedding modules launch the sm_100a kernels, and a fused sparse optimizer applies
the table update inside ``loss.backward()``.
"""
import copy
import pickle
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional

import numpy as np
import os

import torch
from torch.nn import Module
from torch.nn.modules.loss import _Loss  # noqa
from torch.optim.optimizer import Optimizer
from torch.utils.data import DataLoader, Dataset

from ..metric import IMetric, MetricList
from ..utils import set_torch_seed, tensor_to_device


class History:
    """Per-epoch log accumulator (subset of torchrec/callback/History.py:26-40)."""

    def __init__(self):
        self.epoch: List[int] = []
        self.history: Dict[str, List[Any]] = {}

    def on_epoch_end(self, epoch: int, logs: Dict[str, Any]):
        self.epoch.append(epoch)
        for k, v in logs.items():
            if isinstance(v, torch.Tensor):
                v = v.item()
            self.history.setdefault(k, []).append(v)


class _Callbacks:
    """The slice of torchrec/callback/CallbackList.py that ``fit`` drives (IModel.py:160-208): ``set_model`` /
    ``set_params`` at construction, then ``on_train_begin``, ``on_epoch_begin``, ``on_epoch_end``, ``on_train_end``.
    Callbacks are duck-typed (a reference-style ``ICallback`` works unchanged); missing hooks are skipped."""

    def __init__(self, callbacks, model, **params):
        self.callbacks = list(callbacks or [])
        for cb in self.callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(model)
            if hasattr(cb, "set_params"):
                cb.set_params(params)

    def _call(self, hook: str, *args):
        for cb in self.callbacks:
            fn = getattr(cb, hook, None)
            if fn is not None:
                fn(*args)

    def on_train_begin(self, logs=None):
        self._call("on_train_begin", logs)

    def on_train_end(self, logs=None):
        self._call("on_train_end", logs)

    def on_epoch_begin(self, epoch, logs=None):
        self._call("on_epoch_begin", epoch, logs)

    def on_epoch_end(self, epoch, logs=None):
        self._call("on_epoch_end", epoch, logs)


_LOOKUP_CACHE_KEY = "__ptrec_lookup_cache__"


def _drop_lookup_cache(module, args):
    """Forward pre-hook: the per-batch lookup cache (packed ids, shared sort) never outlives one forward."""
    if args and isinstance(args[0], dict):
        args[0].pop(_LOOKUP_CACHE_KEY, None)


import weakref

_LIVE_MODELS: "weakref.WeakSet" = weakref.WeakSet()   # for the at-exit drain of pytorchrec_b200/__init__.py


class _GraphedTrainStep:
    """Whole-step CUDA graph (forward, loss, zero_grad, backward with the fused sparse update, dense
    optimizer step) keyed by the batch signature.  The first ``warmup`` steps of a signature run eagerly —
    they are real training steps and initialise every lazily-allocated buffer (optimizer state, pointer
    arrays, workspaces) outside the capture.  Inputs are copied into static device tensors and the graph
    is replayed; nothing is skipped or cached across steps."""

    def __init__(self, model: "IModel", warmup: int = 2):
        self.model = model
        self.warmup = warmup
        self.entries: Dict[tuple, dict] = {}
        self.replayed_launches = 0  # libptrec kernel launches executed through graph replays

    def step(self, data: Dict[str, torch.Tensor]):
        m = self.model
        sig = tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in data.items() if isinstance(v, torch.Tensor)))
        e = self.entries.setdefault(sig, {"n": 0, "graph": None})
        if e["n"] < self.warmup:
            e["n"] += 1
            return m._eager_train_step(data)
        dev = m.compiled_device
        opt = m.compiled_optimizers
        on_host = all(v.device.type == "cpu" for v in data.values() if isinstance(v, torch.Tensor))
        if e["graph"] is None:
            if not getattr(opt, "graph_safe", lambda: False)():
                raise RuntimeError("CUDA-graph train steps need an optimizer whose update does not depend on the "
                                   "host-side step count (SparseSGD, SparseAdagrad / SparseRowWiseAdagrad with lr_decay=0)")
            # static inputs are typed views of ONE device buffer: host batches arrive with a single pinned,
            # packed, asynchronous copy (N1), device batches are copied key by key
            from ..utils.ingest import BatchPacker
            e["packer"] = BatchPacker(data, dev)
            e["static"] = e["packer"].views
            if on_host:
                e["packer"].load(data)
            else:
                for k, v in e["static"].items():
                    v.copy_(data[k])
            torch.cuda.synchronize(dev)
            m.train()
            from .. import _lib
            l0 = _lib.load().ptrec_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss = m._train_step_body(dict(e["static"]))  # forward may add private entries to its dict
            opt._step_count_fused -= 1  # capture records the step, it does not execute it
            e["graph"], e["loss"] = g, loss
            e["launches"] = int(_lib.load().ptrec_launch_count() - l0)  # recorded into the graph, not yet run
        elif on_host:
            pre = m._take_prefetched(data) if m._prefetched else None
            if pre is not None:   # already on the device (copy stream): one D2D copy into the graph's static inputs
                e["packer"].dev.copy_(pre[1].dev, non_blocking=True)
                pre[2]()
            else:
                e["packer"].load(data)
        elif getattr(data, "buffer", None) is not None and data.signature == e["packer"].signature():
            e["packer"].dev.copy_(data.buffer, non_blocking=True)  # IModel.stage(): one D2D copy for the whole batch
        else:  # loose device tensors: one copy per key
            for k, v in e["static"].items():
                v.copy_(data[k], non_blocking=True)
        e["graph"].replay()
        self.replayed_launches += e["launches"]
        opt._step_count_fused += 1
        return {"loss": e["loss"]}


class IModel(Module, ABC):
    """Model interface: subclasses implement ``_init_weights`` and ``forward(data) -> (prediction, target)``."""

    @classmethod
    def get_argument_descriptions(cls) -> list:
        """Hyper-parameter metadata in the reference's form (torchrec/model/IModel.py + utils/argument): subclasses
        extend the list (FunkSVD.py:13-25); ``random_seed`` is the one argument every model takes (IModel.py:37)."""
        from ..utils.argument import ArgumentDescription
        return [ArgumentDescription(name="random_seed", type_=int, help_info="seed of torch's generators (set_torch_seed)",
                                    default_value=2020)]

    @classmethod
    def check_argument_values(cls, arguments: Dict[str, Any]) -> None:
        """Validate ``arguments`` (name -> value) against the descriptions; unknown names are left to the caller."""
        for d in cls.get_argument_descriptions():
            if d.name in arguments:
                d.check(arguments[d.name])

    def __init__(self, random_seed: int, **kwargs):  # noqa
        set_torch_seed(random_seed)
        super().__init__()
        _LIVE_MODELS.add(self)
        self.stop_training = False
        self.best_state_dict = None
        self.history: Optional[History] = None
        self._is_compiled = False
        self.compiled_optimizers: Optional[Optimizer] = None
        self.compiled_loss: Optional[_Loss] = None
        self.compiled_metrics: Optional[MetricList] = None
        self.compiled_device: Optional[torch.device] = None
        self._graphed: Optional[_GraphedTrainStep] = None
        self._packers: Dict[tuple, Any] = {}
        self.packed_ingest = True  # N1: host batches move with one pinned packed copy instead of one copy per key
        self._prefetch: Dict[tuple, dict] = {}   # batch signature -> {"packers": [2], "i", "stream"}
        self._prefetched: list = []              # [(batch dict, device views, ready event, packer)], at most 2
        self.register_forward_pre_hook(_drop_lookup_cache)
        self._init_weights()
        self._reset_weights()

    @abstractmethod
    def _init_weights(self):
        pass

    @staticmethod
    def _reset_weights_fn(m):
        # same predicate and draw order as IModel.py:61-68
        if 'Linear' in str(type(m)):
            torch.nn.init.normal_(m.weight, mean=0.0, std=0.01)
            if m.bias is not None:
                torch.nn.init.normal_(m.bias, mean=0.0, std=0.01)
        elif 'Embedding' in str(type(m)):
            torch.nn.init.normal_(m.weight, mean=0.0, std=0.01)

    def _reset_weights(self):
        self.apply(self._reset_weights_fn)

    # ------------------------------------------------------------------ weights
    def load_weights(self, filepath: str, device: torch.device):
        # written by save_weights with pickle.HIGHEST_PROTOCOL (as the reference does); torch>=2.6's
        # weights-only unpickler rejects protocol 5, so load our own checkpoint with the full unpickler
        state_dict = torch.load(filepath, map_location=device, weights_only=False)
        self.load_state_dict(state_dict)
        self.to(device)

    def save_weights(self, filepath: str):
        # A table owned by a fused optimizer is a strided view of its interleaved weight | state buffer, and
        # torch.save serialises a view's WHOLE storage: clone to contiguous tensors so that the file holds the
        # weights only (the reference's checkpoint content, IModel.py:79-81) at their own size.
        sd = {k: (v.detach().clone(memory_format=torch.contiguous_format) if isinstance(v, torch.Tensor) else v)
              for k, v in self.state_dict().items()}
        torch.save(sd, filepath, pickle_protocol=pickle.HIGHEST_PROTOCOL)

    def get_parameters(self):
        """Two param groups: weights, and biases with ``weight_decay=0`` (IModel.py:83-92)."""
        weight_p, bias_p = [], []
        for name, p in self.named_parameters():
            if not p.requires_grad:
                continue
            (bias_p if 'bias' in name else weight_p).append(p)
        return [{'params': weight_p}, {'params': bias_p, 'weight_decay': 0.0}]

    # ------------------------------------------------------------------ compile / step
    def compile(self, optimizer: Optimizer, loss: _Loss, metrics: List[IMetric], device: torch.device):
        if not isinstance(optimizer, Optimizer):
            raise ValueError(f"invalid optimizer: {optimizer}")
        if not isinstance(loss, _Loss):
            raise ValueError(f"invalid loss: {loss}")
        if (not isinstance(metrics, list)) or not all(isinstance(m, IMetric) for m in metrics):
            raise ValueError(f"invalid metrics: {metrics}")
        if not isinstance(device, torch.device):
            raise ValueError(f"invalid device: {device}")
        self.compiled_optimizers = optimizer
        self.compiled_loss = loss
        self.compiled_metrics = MetricList(metrics)
        self.compiled_device = device
        self.to(device)
        # an optimizer that allocates its state at construction (torch.optim.Adagrad) did so before this move: the
        # reference builds the optimizer first and compiles second (RepeatTask.py:96), so follow the parameters
        for p_, st in optimizer.state.items():
            if torch.is_tensor(p_):
                for k, v in st.items():  # ("step" counters stay where the optimizer put them: host tensors by default)
                    if torch.is_tensor(v) and k != "step" and v.device != p_.device:
                        st[k] = v.to(p_.device)
        self._is_compiled = True

    def enable_cuda_graph(self, enabled: bool = True, warmup: int = 2) -> None:
        """Replay ``train_step`` as one CUDA graph per batch signature (CUDA devices only).  Do not keep
        outputs of a grad-enabled forward alive across the capturing step: their autograd graph pins the
        parameters' AccumulateGrad nodes to the default stream, which cannot join a capture."""
        self._graphed = _GraphedTrainStep(self, warmup) if enabled else None

    def train_step(self, data: Dict):
        for m in self._flag_pollers():  # non-synchronising: error words published by earlier steps (pinned host memory)
            m.poll_errors()
        if self._graphed is not None and self.compiled_device is not None and self.compiled_device.type == "cuda":
            return self._graphed.step(data)
        return self._eager_train_step(data)

    def prefetch(self, data: Dict) -> Dict:
        """N1: start the host -> device transfer of a batch that a later ``train_step`` / ``test_step`` call will
        consume, on a copy stream, so that it overlaps the step that is running (the data-loader pattern: call it
        for batch k+1 right after ``train_step(batch k)``).  Double-buffered; passing the same dict object to the
        step picks the transfer up.  Returns ``data``; a no-op for device batches or CPU models."""
        dev = self.compiled_device
        if not (self.packed_ingest and dev is not None and dev.type == "cuda" and isinstance(data, dict) and data
                and all(isinstance(v, torch.Tensor) and v.device.type == "cpu" for v in data.values())):
            return data
        from ..utils.ingest import BatchPacker
        sig = tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in data.items()))
        pf = self._prefetch.get(sig)
        if pf is None:
            pf = self._prefetch[sig] = {"packers": [BatchPacker(data, dev), BatchPacker(data, dev)], "i": 0,
                                        "stream": torch.cuda.Stream(dev), "consumed": [None, None]}
        i = pf["i"]
        pf["i"] ^= 1
        packer, stream = pf["packers"][i], pf["stream"]
        if pf["consumed"][i] is not None:
            stream.wait_event(pf["consumed"][i])  # the step that read this buffer two prefetches ago is done with it
        with torch.cuda.stream(stream):
            views = packer.load(data)
            ready = torch.cuda.Event()
            ready.record(stream)
        self._prefetched = [r for r in self._prefetched if r[0] is not data][-1:] + [(data, views, ready, packer, pf, i)]
        return data

    def stage(self, data: Dict) -> Dict:
        """A device-resident, packed copy of a batch (N1): ``train_step`` / ``test_step`` accept it like any batch
        dict; the CUDA-graph step moves it into its static inputs with one copy instead of one per key."""
        dev = self.compiled_device
        if dev is None or dev.type != "cuda":
            return tensor_to_device(data, dev)
        from ..utils.ingest import BatchPacker
        sig = tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in data.items()))
        packer = self._packers.get(sig)
        if packer is None:
            packer = self._packers[sig] = BatchPacker(data, dev)
        return packer.stage(data)

    def pack_host(self, data: Dict) -> Dict:
        """N1: ``data`` (host tensors) re-assembled in ONE pinned buffer laid out like the device-side packed batch —
        what a loader worker / collate function hands over when it builds the batch in page-locked memory.  Still a
        ``Dict[str, Tensor]``; ``prefetch`` / ``train_step`` then move it with a single DMA instead of one per key."""
        dev = self.compiled_device
        if dev is None or dev.type != "cuda":
            return data
        from ..utils.ingest import BatchPacker
        sig = tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in data.items()))
        packer = self._packers.get(sig)
        if packer is None:
            packer = self._packers[sig] = BatchPacker(data, dev)
        return packer.pack_host(data)

    def _take_prefetched(self, data: Dict):
        """(device views, packer, release) of a batch handed to ``prefetch`` earlier, or None.  The current stream is
        made to wait for the transfer; ``release()`` must be called once the step's reads of the views are enqueued."""
        for n, rec in enumerate(self._prefetched):
            if rec[0] is data:
                del self._prefetched[n]
                _, views, ready, packer, pf, i = rec
                cur = torch.cuda.current_stream(self.compiled_device)
                cur.wait_event(ready)

                def release():
                    ev = torch.cuda.Event()
                    ev.record(cur)
                    pf["consumed"][i] = ev
                return views, packer, release
        return None

    def _to_device(self, data: Dict):
        """Host -> device move of one batch.  CUDA target + host tensors: one packed pinned transfer (N1);
        otherwise the reference's per-key ``tensor_to_device`` (IModel.py:119)."""
        dev = self.compiled_device
        pre = self._take_prefetched(data) if self._prefetched else None
        if pre is not None:
            self._release_after_step = pre[2]
            return pre[0]
        if (self.packed_ingest and dev is not None and dev.type == "cuda" and isinstance(data, dict) and data
                and all(isinstance(v, torch.Tensor) and v.device.type == "cpu" for v in data.values())):
            sig = tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in data.items()))
            packer = self._packers.get(sig)
            if packer is None:
                from ..utils.ingest import BatchPacker
                packer = self._packers[sig] = BatchPacker(data, dev)
            return packer.load(data)
        return tensor_to_device(data, dev)

    def _eager_train_step(self, data: Dict):
        self.train()
        data = self._to_device(data)
        # detached: a loss that keeps its autograd graph alive would pin AccumulateGrad nodes to the stream of this
        # eager step and break a later CUDA-graph capture of the same parameters
        logs = {"loss": self._train_step_body(data).detach()}
        rel = getattr(self, "_release_after_step", None)
        if rel is not None:  # the prefetch buffer may be refilled once this step's kernels have read it
            self._release_after_step = None
            rel()
        return logs

    def _train_step_body(self, data: Dict):
        """forward, loss, zero_grad, backward, [gradient hook], step — IModel.py:120-124."""
        prediction, target = self(data)
        loss = self._apply_loss(prediction, target)
        self.compiled_optimizers.zero_grad()
        from .layer import dense as _dense
        _dense._DEFER_JOIN[0] = loss.is_cuda   # side-stream weight gradients are joined below, not inside backward()
        if loss.is_cuda:
            from .layer.embedding import reset_join_streams
            reset_join_streams(loss.device)
        try:
            loss.backward()
        finally:
            _dense._DEFER_JOIN[0] = False
        if loss.is_cuda:
            from .layer.embedding import join_aux_streams
            # lookups placed on the aux stream update their tables there; the tower's weight gradients run on theirs
            join_aux_streams(loss.device)
        self._before_optimizer_step()
        self.compiled_optimizers.step(closure=None)
        return loss

    def _apply_loss(self, prediction, target):
        """``self.compiled_loss(prediction, target)``; a plain ``BCEWithLogitsLoss()`` on CUDA runs as K9 (forward and
        gradient in one launch, ``PTREC_FUSED_LOSS=0`` keeps the module)."""
        lf = self.compiled_loss
        if (type(lf) is torch.nn.BCEWithLogitsLoss and lf.reduction == "mean" and lf.weight is None
                and lf.pos_weight is None and torch.is_tensor(prediction) and prediction.is_cuda
                and os.environ.get("PTREC_FUSED_LOSS", "1") != "0"):
            from .. import ops
            out = ops.bce_logits_mean(prediction, target)
            if out is not None:
                return out
        return lf(prediction, target)

    def _flag_pollers(self) -> list:
        p = self.__dict__.get("_pollers")
        if p is None:
            p = self.__dict__["_pollers"] = [m for m in self.modules() if m is not self and hasattr(m, "poll_errors")]
        return p

    def _before_optimizer_step(self) -> None:
        """Hook between backward and step (data-parallel models all-reduce dense gradients here)."""

    def test_step(self, data):
        self.eval()
        data = self._to_device(data)
        prediction, target = self(data)
        rel = getattr(self, "_release_after_step", None)
        if rel is not None:
            self._release_after_step = None
            rel()
        return prediction, target

    def predict_step(self, data):
        prediction, _ = self.test_step(data)
        return prediction

    # ------------------------------------------------------------------ fit / evaluate
    def _data_parallel_rank(self):
        """(rank, world_size) of this replica among the data-parallel ranks; (0, 1) for single-device models.  Row-wise
        sharded models (``distributed.sharded``) train one batch slice per rank."""
        sh = getattr(self, "sharded", None)
        return (int(sh.rank), int(sh.world)) if sh is not None and hasattr(sh, "world") else (0, 1)

    def _loader(self, dataset, batch_size: int, shuffle: bool, workers: int, drop_last: bool, train: bool = False):
        """Batches of ``dataset`` in ``DataLoader`` order.  A tensor-native split (``data.SplitDataset``, N2) assembles
        whole batches by index — same samples, same order under the same seed; when TRAINING a data-parallel model it
        yields this rank's strided slice of that order (evaluation runs the whole split on every rank, so metrics need
        no reduction).  Any other ``Dataset`` goes through ``torch.utils.data.DataLoader`` exactly as in the reference
        (IModel.py:183-186,242,294)."""
        if hasattr(dataset, "batches"):
            rank, world = self._data_parallel_rank() if train else (0, 1)
            return dataset.batches(batch_size, shuffle=shuffle, drop_last=drop_last, rank=rank, world_size=world)
        return DataLoader(dataset=dataset, batch_size=batch_size, shuffle=shuffle, num_workers=workers,
                          drop_last=drop_last)

    def fit(self, dataset: Dataset, batch_size: int, epochs: int, dev_dataset: Optional[Dataset] = None,
            train_mode=None, verbose: int = 0, callbacks: Optional[list] = None, shuffle: bool = True,
            workers: int = 0, drop_last: bool = False, dev_batch_size: Optional[int] = None,
            dev_freq: int = 1) -> History:
        self._assert_compile_was_called()
        if not (hasattr(callbacks, "on_train_begin") and hasattr(callbacks, "on_epoch_end")):  # a plain list
            n = len(dataset) if hasattr(dataset, "__len__") else 0
            batches = n // batch_size if drop_last else (n + batch_size - 1) // batch_size
            callbacks = _Callbacks(callbacks, self, verbose=verbose, epochs=epochs, batches=batches)
        self.history = History()
        self.stop_training = False
        logs: Dict[str, Any] = {}
        callbacks.on_train_begin()
        for epoch in range(epochs):
            callbacks.on_epoch_begin(epoch)
            if train_mode is not None and getattr(train_mode, "value", train_mode) == "pair_wise" \
                    and hasattr(dataset, "train_neg_sample"):
                dataset.train_neg_sample()
            it = iter(self._loader(dataset, batch_size, shuffle, workers, drop_last, train=True))
            data = next(it, None)
            if data is not None:
                self.prefetch(data)
            while data is not None:
                logs = self.train_step(data)
                data = next(it, None)       # the loader builds batch k+1 and its transfer starts while step k runs
                if data is not None:
                    self.prefetch(data)
            epoch_logs = copy.copy(logs)
            if dev_dataset is not None and (epoch + 1) % dev_freq == 0:
                epoch_logs.update(self.evaluate(dev_dataset, dev_batch_size or batch_size, workers=workers))
            self._check_device_flags()
            self.history.on_epoch_end(epoch, epoch_logs)
            callbacks.on_epoch_end(epoch, epoch_logs)
            if self.stop_training:
                break
        callbacks.on_train_end()
        return self.history

    def _check_device_flags(self) -> None:
        """Once per epoch (one host sync): raise on any error word the kernels set during it — an out-of-range id
        (the reference raises IndexError from index_select on the CPU) or an overflowed exchange list of a row-wise
        sharded table — instead of training on silently."""
        for m in self.modules():
            chk = getattr(m, "check_errors", None) or getattr(m, "check_index_errors", None)
            if chk is not None and m is not self:
                chk()

    @torch.no_grad()
    def evaluate(self, dataset: Dataset, batch_size: int, verbose: int = 0, callbacks=None, workers: int = 0):
        self._assert_compile_was_called()
        loader = self._loader(dataset, batch_size, False, workers, False)
        predictions, targets = [], []
        # N3: with ranking metrics only, scores stay on the device and ranks are computed there — one
        # synchronisation per evaluate() instead of the reference's .cpu().numpy() per batch (IModel.py:250-251)
        keep_on_device = self.compiled_metrics.rank_only() and self.compiled_device.type == "cuda"
        for data in loader:
            prediction, target = self.test_step(data)
            if keep_on_device:
                predictions.append(prediction.detach())
            else:
                predictions.append(prediction.detach().cpu().numpy())
                targets.append(target.detach().cpu().numpy())
        if keep_on_device:
            return self.compiled_metrics(torch.cat(predictions), None)
        return self.compiled_metrics(np.concatenate(predictions), np.concatenate(targets))

    @torch.no_grad()
    def predict(self, dataset: Dataset, batch_size: int, verbose: int = 0, callbacks=None, workers: int = 0):
        loader = self._loader(dataset, batch_size, False, workers, False)
        return np.concatenate([self.predict_step(d).detach().cpu().numpy() for d in loader])

    def _assert_compile_was_called(self):
        if not self._is_compiled:
            raise RuntimeError('compile() must be called before fit/evaluate')

    def save_best_weights(self):
        self.best_state_dict = copy.deepcopy(tensor_to_device(dict(self.state_dict()), device=torch.device("cpu")))

    def load_best_weights(self):
        assert self.best_state_dict is not None
        self.load_state_dict(self.best_state_dict)
        self.to(self.compiled_device)
