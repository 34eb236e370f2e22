"""Embedding modules backed by the sm_100a kernels.

The reference's embedding seam is ``torch.nn.Embedding(column.category_num, emb_size)`` created in
``_init_weights`` and indexed with ``column.get_feature_data(batch)`` (torchrec/model/FunkSVD.py:39-48,
SVDPP.py:36-61, NCF.py:38-65).  ``EmbeddingTable`` is a drop-in for that module (same constructor,
same ``weight`` parameter name, same RNG consumption at construction, caught by
``IModel._reset_weights_fn`` because its type name contains 'Embedding', IModel.py:61-68).
``MultiTableEmbedding`` is an ``nn.ModuleList`` of such tables built from a list of columns; it
runs all lookups + pooling of a batch in one gather launch and one fused backward.

Backward semantics: gradients never materialise as ``[rows, D]`` tensors.  When the tables are
owned by a ``pytorchrec_b200.optim`` sparse optimizer the row update happens inside ``backward()``
(sort -> dedup -> segment-sum -> update).  Otherwise (stock ``torch.optim`` / the reference's ``AdamW``)
``weight.grad`` is built from the same segment sums with ``nn.Embedding``'s own convention: a dense
``[rows, D]`` tensor by default (``sparse=False`` — what the reference's models produce, so ``adam`` / ``adamw`` /
``sgd`` with weight decay keep working unchanged), a coalesced ``torch.sparse_coo_tensor`` with ``sparse=True``.
"""
import os
from typing import Dict, List, Optional, Sequence, Union

import torch
from torch import Tensor, nn

from ... import ops
from ...feature_column import CategoricalColumn

_MASKED = -1  # 0xFFFFFFFF viewed as int32
_CACHE_KEY = "__ptrec_lookup_cache__"  # private per-batch entry (never a tensor: tensor_to_device skips it)


class EmbeddingGroup:
    """Non-module helper: a set of tables sharing ``dim`` + the feature layout reading them."""

    def __init__(self, tables: Sequence["EmbeddingTable"], dim: int):
        self.tables = list(tables)
        self.dim = dim
        self.table_set = ops.TableSet()
        self._err: Dict[torch.device, Tensor] = {}
        self._state_cache = None

    def weights(self) -> List[Tensor]:
        return [t.weight for t in self.tables]

    def err_flag(self, device) -> Tensor:
        e = self._err.get(device)
        if e is None:
            e = torch.zeros(1, dtype=torch.int32, device=device)
            self._err[device] = e
        return e

    def check_index_errors(self) -> None:
        """Synchronising check: raises IndexError if any lookup so far was out of range (the
        reference raises IndexError from ``index_select`` on CPU)."""
        for e in self._err.values():
            if int(e.item()) != 0:
                e.zero_()
                raise IndexError("embedding id out of range [0, category_num)")

    # -------------------------------------------------------------------------------- backward
    def binding(self):
        b = None
        for t in self.tables:
            tb = getattr(t.weight, "_ptrec_optim", None)
            if b is None:
                b = tb
            elif tb is not None and (tb[0] is not b[0] or tb[1] is not b[1]):
                raise RuntimeError("all tables of one MultiTableEmbedding must sit in the same param "
                                   "group of the same fused optimizer")
            elif tb is None and b is not None:
                raise RuntimeError("some tables of this embedding group are not owned by the fused optimizer")
        return b

    def apply_backward(self, layout, ids, lens, bag_scale, batch, grad_out, shared=None):
        bind = self.binding()
        prepared = None
        if bind is not None:
            # may re-house the tables (weight | state interleaving) on first use: do it before taking pointers
            prepared = bind[0]._fused_prepare(self, bind[1])
        tables = self.table_set.refresh([w.data for w in self.weights()])
        # modules reading the same columns (e.g. FM's embedding and first-order tables) share one sort
        srt = shared.get("sort") if shared is not None else None
        if srt is not None and shared.get("sort_event") is not None:
            # join the early sort; the event stays: a second consumer may run on another stream (aux_stream below)
            torch.cuda.current_stream(ids.device).wait_event(shared["sort_event"])
        if grad_out.is_cuda:
            # this backward may run on the aux stream (autograd runs a node on its forward's stream): everything it
            # reads that was allocated on another stream must not be recycled before its kernels have finished
            cur = torch.cuda.current_stream(grad_out.device)
            for t in (grad_out, ids, lens, bag_scale) + ((srt.sorted_keys, srt.perm, srt.seg_start, srt.seg_meta, srt.n_seg)
                                                         if srt is not None else ()):
                if t is not None:
                    t.record_stream(cur)
        if srt is None:
            srt = ops.sort_dedup(tables, layout, ids, lens, batch)
            if shared is not None:
                shared["sort"] = srt
        if bind is not None:
            s1, s2, args = prepared
            ops.bwd_fused(tables, s1, s2, layout, batch, srt, grad_out, bag_scale, args)
            return [None] * len(self.tables)
        # stock-optimizer mode: coalesced sparse gradients (one host sync for the segment count)
        row_grad = ops.segment_sum(layout, batch, srt, grad_out, bag_scale)
        n = int(srt.n_seg.item())
        keys = srt.sorted_keys[srt.seg_start[:n].long()]
        seg_table = srt.seg_table[:n]
        valid = keys != _MASKED
        grads = []
        for t, table in enumerate(self.tables):
            sel = valid & (seg_table == t)
            idx = keys[sel].long() & 0xFFFFFFFF
            rows_g = row_grad[:n][sel].to(table.weight.dtype)  # a bf16 table takes a bf16 gradient (autograd's rule)
            if getattr(table, "sparse", False):
                g = torch.sparse_coo_tensor(idx.unsqueeze(0), rows_g, size=table.weight.shape, check_invariants=False)
                grads.append(g._coalesced_(True))  # segments are unique sorted rows by construction
            else:  # nn.Embedding(sparse=False): the dense gradient every stock optimizer accepts
                g = torch.zeros(table.weight.shape, dtype=table.weight.dtype, device=row_grad.device)
                g[idx] = rows_g                  # unique rows: a plain scatter, no accumulation order involved
                grads.append(g)
        return grads


_SIDE_STREAMS: Dict[torch.device, "torch.cuda.Stream"] = {}
_AUX_STREAMS: Dict[torch.device, "torch.cuda.Stream"] = {}
_AUX_USED: Dict[torch.device, bool] = {}   # the aux stream was forked since the last join_aux_streams


class aux_stream:
    """``with aux_stream(device) as s:`` runs a small, independent lookup (FM's first-order tables: 4-byte rows) on a
    second stream forked from the current one, so that its latency-bound gather — and, because autograd runs a
    node's backward on its forward's stream, its fused update — overlaps the wide tables' kernels instead of queueing
    behind them.  ``s.join(t)`` before the first use of a result on the main stream.  A no-op object on CPU or with
    ``PTREC_AUX_STREAM=0``.  Captured in a step graph as a fork / join."""

    def __init__(self, device):
        import os
        self.on = device is not None and torch.device(device).type == "cuda" and os.environ.get("PTREC_AUX_STREAM", "1") != "0"
        self.device = device

    def __enter__(self):
        if self.on:
            self.main = torch.cuda.current_stream(self.device)
            self.side = _AUX_STREAMS.get(self.device)
            if self.side is None:
                self.side = _AUX_STREAMS[self.device] = torch.cuda.Stream(self.device)
            self.side.wait_stream(self.main)
            _AUX_USED[torch.device(self.device)] = True
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.ctx.__exit__(*exc)
        return False

    def join(self, *tensors) -> None:
        if self.on:
            self.main.wait_stream(self.side)
            for t in tensors:
                if t is not None:
                    t.record_stream(self.main)


_JOIN_STREAMS: Dict[torch.device, list] = {}   # other side streams forked since the last join (dense.py: weight gradients)


def register_join_stream(device, stream) -> None:
    """``stream`` was forked from the current stream during this step's backward: ``join_aux_streams`` joins it."""
    lst = _JOIN_STREAMS.setdefault(torch.device(device), [])
    if all(s is not stream for s in lst):
        lst.append(stream)


def reset_join_streams(device) -> None:
    """Start of a train step's backward: forget side streams registered by a step that did not reach its join (an
    exception between backward and the optimizer step) — joining a stream that is not part of a running graph capture
    would invalidate the capture."""
    if device is not None:
        _JOIN_STREAMS.pop(torch.device(device), None)


def join_aux_streams(device) -> None:
    """Make the current stream wait for whatever the side streams still run (a backward that autograd placed on the aux
    stream, the tower's weight-gradient GEMMs): called by ``IModel`` between ``backward()`` and ``optimizer.step()``."""
    if device is None:
        return
    for s in _JOIN_STREAMS.pop(torch.device(device), []):
        torch.cuda.current_stream(device).wait_stream(s)
    if not _AUX_USED.pop(torch.device(device), False):
        return   # nothing forked in this step (waiting on a stream outside a running graph capture would invalidate it)
    side = _AUX_STREAMS.get(torch.device(device))
    if side is not None:
        torch.cuda.current_stream(device).wait_stream(side)


def _early_sort_enabled() -> bool:
    import os
    return os.environ.get("PTREC_EARLY_SORT", "1") != "0"


def _start_early_sort(tables, layout, ids, lens, batch, shared) -> None:
    """K2a depends on the ids only, so it is started during the FORWARD on a side stream and overlaps the dense
    tower (its kernels are latency-bound and co-reside with the persistent GEMM CTAs); the backward joins it with
    an event.  Under CUDA-graph capture the fork / join becomes part of the graph."""
    dev = ids.device
    main = torch.cuda.current_stream(dev)
    side = _SIDE_STREAMS.get(dev)
    if side is None:
        side = _SIDE_STREAMS[dev] = torch.cuda.Stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        srt = ops.sort_dedup(tables, layout, ids, lens, batch)
        ev = torch.cuda.Event()
        ev.record(side)
    for t in (srt.sorted_keys, srt.perm, srt.seg_start, srt.seg_meta, srt.n_seg):
        t.record_stream(main)  # consumed on the main stream in backward
    shared["sort"], shared["sort_event"] = srt, ev


class _FusedLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, group: EmbeddingGroup, layout, ids, lens, batch, shared, *weights):
        tables = group.table_set.refresh([w.detach() for w in weights])
        needs_scale = any(fd.pooling != 0 for fd in layout.host)
        if (shared is not None and "sort" not in shared and any(ctx.needs_input_grad[6:]) and _early_sort_enabled()):
            _start_early_sort(tables, layout, ids, lens, batch, shared)
        out, bag_scale = ops.gather_pool_fwd(tables, layout, ids, lens, batch, want_scale=needs_scale,
                                             err_flag=group.err_flag(ids.device))
        ctx.group, ctx.layout, ctx.batch, ctx.shared = group, layout, batch, shared
        ctx.save_for_backward(ids, lens, bag_scale)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        ids, lens, bag_scale = ctx.saved_tensors
        if not (grad_out.stride(1) == 1 and grad_out.data_ptr() % 16 == 0 and grad_out.stride(0) % 4 == 0):
            grad_out = grad_out.contiguous()
        grads = ctx.group.apply_backward(ctx.layout, ids, lens, bag_scale, ctx.batch, grad_out, ctx.shared)
        return (None, None, None, None, None, None, *grads)


class _FusedLookupAttention(torch.autograd.Function):
    """DIN's history lookup and attention pooling as ONE node: K4 reads the keys from the item / category tables by id
    (no gathered [B, L, 2D] tensor), the candidate rows (the query) are the only rows gathered; the backward writes the
    key gradients and the query gradient into one [B, 1 + L, 2D] buffer that feeds the same sort / dedup / fused update
    as ``_FusedLookup`` (one update per table and step, as with the materialised lookup)."""

    @staticmethod
    def forward(ctx, group: EmbeddingGroup, layout, ids, shared, B, L, lens, blocks, *rest):
        att, weights = rest[:6], rest[6:]
        dev = ids.device
        tables = group.table_set.refresh([w.detach() for w in weights])
        batch = B * (1 + L)
        if (shared is not None and "sort" not in shared and any(ctx.needs_input_grad[14:]) and _early_sort_enabled()):
            _start_early_sort(tables, layout, ids, None, batch, shared)
        F = len(blocks)
        q_ids = ids.view(F, B, 1 + L)[:, :, 0].contiguous().view(-1)
        q, _ = ops.gather_pool_fwd(tables, layout, q_ids, None, B, want_scale=False, err_flag=group.err_flag(dev))
        # user column p: its ids are block blocks[p][0] of `ids`, its table is weights[blocks[p][1]]
        tabs = [weights[t].detach() for _, t in blocks]
        idl = [ids[k * batch:(k + 1) * batch] for k, _ in blocks]
        params = tuple(p.detach() for p in att)
        pooled = ops.din_attn_pool_fwd_ids(q, tabs, idl, 1 + L, 1, group.err_flag(dev), lens, L, params)
        ctx.group, ctx.layout, ctx.shared, ctx.B, ctx.L, ctx.blocks = group, layout, shared, B, L, blocks
        ctx.save_for_backward(ids, lens, q, *att, *weights)
        return q, pooled

    @staticmethod
    def backward(ctx, g_q_out, g_pooled):
        ids, lens, q = ctx.saved_tensors[:3]
        att = ctx.saved_tensors[3:9]
        weights = ctx.saved_tensors[9:]
        B, L = ctx.B, ctx.L
        DQ = q.shape[1]
        batch = B * (1 + L)
        if g_pooled is None:
            g_pooled = torch.zeros_like(q)
        g_seq = torch.empty(B, 1 + L, DQ, dtype=torch.float32, device=q.device)
        tabs = [weights[t].detach() for _, t in ctx.blocks]
        idl = [ids[k * batch:(k + 1) * batch] for k, _ in ctx.blocks]
        g_q, gp = ops.din_attn_pool_bwd_ids(q, tabs, idl, 1 + L, 1, lens, L, tuple(p.detach() for p in att), g_pooled,
                                            g_seq[:, 1:])
        if g_q_out is not None:
            torch.add(g_q, g_q_out, out=g_seq[:, 0])
        else:
            g_seq[:, 0] = g_q
        grads = ctx.group.apply_backward(ctx.layout, ids, None, None, batch, g_seq.view(batch, DQ), ctx.shared)
        return (None,) * 8 + tuple(gp) + tuple(grads)


class EmbeddingTable(nn.Module):
    """``nn.Embedding``-compatible table: ``EmbeddingTable(num_embeddings, embedding_dim)``,
    parameter ``weight`` ``[num_embeddings, embedding_dim]`` fp32, N(0,1) at construction like
    ``nn.Embedding.reset_parameters``.  ``forward(ids)`` returns ``weight[ids]`` for ids of any shape.
    ``sparse`` has ``nn.Embedding``'s meaning and matters only without a fused optimizer (module docstring).
    ``dtype=torch.bfloat16`` stores the rows in bf16 (half the HBM bytes per lookup; ``include/ptrec_b200.h``
    PTREC_BF16): lookups widen exactly to fp32, the fused optimizers keep fp32 state, compute the step in fp32 and
    round the row back to bf16 (nearest even).  The initial draw is the fp32 one, rounded."""

    def __init__(self, num_embeddings: int, embedding_dim: int, device=None, sparse: bool = False,
                 dtype: torch.dtype = torch.float32):
        super().__init__()
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("EmbeddingTable dtype must be torch.float32 or torch.bfloat16")
        self.num_embeddings = int(num_embeddings)
        self.embedding_dim = int(embedding_dim)
        self.sparse = bool(sparse)
        w = torch.empty(self.num_embeddings, self.embedding_dim, device=device)
        nn.init.normal_(w)
        self.weight = nn.Parameter(w if dtype == torch.float32 else w.to(dtype))
        self._group = None
        self._layout = None
        self._tag()

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._tag()
        return out

    def _tag(self):
        self.weight._ptrec_table = self

    def forward(self, ids: Tensor) -> Tensor:
        if self._group is None:
            self._group = EmbeddingGroup([self], self.embedding_dim)
            self._layout = ops.FeatureLayout([{"table": 0, "bag_len": 1}], self.embedding_dim, 1)
        self._tag()
        flat = ids.reshape(-1)
        if flat.dtype != torch.int64:
            flat = flat.long()
        flat = flat.contiguous()
        out = _FusedLookup.apply(self._group, self._layout, flat, None, flat.numel(), None, self.weight)
        return out.view(*ids.shape, self.embedding_dim)

    def pooled(self, ids: Tensor, pooling: str = "sum", mask: str = "pad", lens: Optional[Tensor] = None) -> Tensor:
        """Masked bag pooling of padded ``ids [B, L]`` -> ``[B, D]`` without materialising ``[B, L, D]``:
        the fused form of SVDPP.py:49-55 (``pooling='sqrtn', mask='pad'``) and SASRec.py:109-110
        (``pooling='mean'``, ``mask='pad_keep_first'`` or ``'lens'``)."""
        assert ids.dim() == 2
        if self._group is None:
            self._group = EmbeddingGroup([self], self.embedding_dim)
            self._layout = ops.FeatureLayout([{"table": 0, "bag_len": 1}], self.embedding_dim, 1)
        B, L = ids.shape
        key = (L, pooling, mask)
        cache = self.__dict__.setdefault("_bag_layouts", {})
        lay = cache.get(key)
        if lay is None:
            lay = ops.FeatureLayout([{"table": 0, "bag_len": L, "pooling": pooling, "mask": mask,
                                      "lens_col": 0 if mask == "lens" else -1}], self.embedding_dim, 1)
            cache[key] = lay
        self._tag()
        flat = ids.reshape(-1)
        flat = (flat if flat.dtype == torch.int64 else flat.long()).contiguous()
        lens32 = None
        if mask == "lens":
            if lens is None:
                raise ValueError("mask='lens' needs lens")
            lens32 = lens.reshape(1, B).to(torch.int32).contiguous()
        return _FusedLookup.apply(self._group, lay, flat, lens32, B, None, self.weight)

    def check_index_errors(self) -> None:
        if self._group is not None:
            self._group.check_index_errors()

    def extra_repr(self):
        return f"{self.num_embeddings}, {self.embedding_dim}"


class MultiTableEmbedding(nn.ModuleList):
    """All sparse features of a model behind one fused lookup.

    Children are ``EmbeddingTable`` modules, one per distinct table, registered in column order —
    so ``state_dict()`` keys (``<name>.<i>.weight``) and seeded initialisation match an
    ``nn.ModuleList([nn.Embedding(c.category_num, emb_size) for c in columns])`` written in the
    reference's idiom.

    :param columns: categorical columns, one per sparse feature.
    :param emb_size: embedding dim D (1, 2 or a multiple of 4 up to 128).
    :param pooling: 'sum' | 'mean' | 'sqrtn' for multi-hot features ([B, L] ids); one-hot ([B]) features ignore it.
    :param mask: 'none' | 'pad' (id != 0, SVDPP.py:49) | 'pad_keep_first' (model/utils.py:5-10) | 'lens'.
    :param lens_columns: for mask='lens', ``{feature_name: CategoricalColumn}`` giving the valid length.
    :param share: ``{feature_name: feature_name_of_table_owner}`` — several features reading one table.
    :param sparse: ``nn.Embedding``'s flag for the gradient layout without a fused optimizer (dense by default).
    :param dtype: ``torch.float32`` (default) or ``torch.bfloat16`` table storage (see ``EmbeddingTable``).
    ``pooling`` / ``mask`` may also be dicts keyed by feature name.
    ``forward(batch)`` returns ``[B, F, D]`` in ``columns`` order.
    """

    def __init__(self, columns: Sequence[CategoricalColumn], emb_size: int,
                 pooling: Union[str, Dict[str, str]] = "sum", mask: Union[str, Dict[str, str]] = "none",
                 lens_columns: Optional[Dict[str, CategoricalColumn]] = None,
                 share: Optional[Dict[str, str]] = None, device=None, sparse: bool = False,
                 dtype: torch.dtype = torch.float32):
        super().__init__()
        self.columns = list(columns)
        self.emb_size = int(emb_size)
        names = [getattr(c, "feature_name", str(i)) for i, c in enumerate(self.columns)]
        self.feature_names = names
        share = share or {}
        self._table_of: List[int] = []
        owner_index: Dict[str, int] = {}
        for i, (c, name) in enumerate(zip(self.columns, names)):
            owner = share.get(name, name)
            if owner not in owner_index:
                if owner != name:
                    raise ValueError(f"feature {name} shares the table of {owner}, which must come first")
                owner_index[owner] = len(self)
                self.append(EmbeddingTable(c.category_num, self.emb_size, device=device, sparse=sparse, dtype=dtype))
            self._table_of.append(owner_index[owner])
        self._pooling = {n: (pooling.get(n, "sum") if isinstance(pooling, dict) else pooling) for n in names}
        self._mask = {n: (mask.get(n, "none") if isinstance(mask, dict) else mask) for n in names}
        self._lens_columns = lens_columns or {}
        # internal feature order: by table, then by position
        self._order = sorted(range(len(names)), key=lambda i: (self._table_of[i], i))
        self._layouts: Dict[tuple, ops.FeatureLayout] = {}
        self._group: Optional[EmbeddingGroup] = None

    @property
    def tables(self) -> List[EmbeddingTable]:
        return [m for m in self]

    @property
    def weight(self) -> Tensor:
        """Zero-size placeholder.  The reference's ``_reset_weights_fn`` (IModel.py:61-68) calls
        ``normal_(m.weight)`` on every module whose type name contains 'Embedding'; the container
        itself owns no parameter (its children do), and ``normal_`` on an empty tensor draws nothing."""
        return torch.empty(0)

    def _layout_for(self, bag_lens: tuple) -> ops.FeatureLayout:
        lay = self._layouts.get(bag_lens)
        if lay is None:
            specs = []
            lens_cols: List[str] = []
            for i in self._order:
                name = self.feature_names[i]
                m = self._mask[name] if bag_lens[i] > 1 or self._mask[name] == "lens" else "none"
                lens_col = -1
                if m == "lens":
                    if name not in self._lens_columns:
                        raise ValueError(f"feature {name}: mask='lens' needs lens_columns[{name!r}]")
                    if name not in lens_cols:
                        lens_cols.append(name)
                    lens_col = lens_cols.index(name)
                specs.append({"table": self._table_of[i], "bag_len": bag_lens[i],
                              "pooling": self._pooling[name], "mask": m, "lens_col": lens_col})
            lay = ops.FeatureLayout(specs, self.emb_size, len(self))
            # output columns follow the user's column order, not the internal (table-sorted) one
            for k, i in enumerate(self._order):
                lay.host[k].out_col = i * self.emb_size
            lay.lens_names = lens_cols
            self._layouts[bag_lens] = lay
        return lay

    def forward(self, batch: Dict[str, Tensor]) -> Tensor:
        if self._group is None:
            self._group = EmbeddingGroup(self.tables, self.emb_size)
        id_list = [c.get_feature_data(batch) for c in self.columns]
        B = id_list[0].shape[0]
        bag_lens = tuple(1 if t.dim() == 1 else int(t.shape[1]) for t in id_list)
        layout = self._layout_for(bag_lens)
        # Per-batch cache kept inside the batch dict: modules reading the same columns with the same
        # masks and table heights share the packed id tensor and, in backward, ONE sort/dedup.
        cache = batch.get(_CACHE_KEY) if isinstance(batch, dict) else None
        if cache is None:
            cache = {}
            if isinstance(batch, dict):
                batch[_CACHE_KEY] = cache
        key = (tuple(self.feature_names[i] for i in self._order), bag_lens,
               tuple(self._mask[self.feature_names[i]] for i in self._order),
               tuple(t.num_embeddings for t in self.tables), tuple(self._table_of),
               tuple((t.data_ptr(), t._version) for t in id_list))
        shared = cache.get(key)
        if shared is None:
            ids = torch.cat([id_list[i].reshape(-1) for i in self._order]) if len(id_list) > 1 \
                else id_list[0].reshape(-1).contiguous()
            lens = None
            if layout.lens_names:
                lens = torch.stack([self._lens_columns[n].get_feature_data(batch).reshape(-1)
                                    for n in layout.lens_names]).to(torch.int32).contiguous()
            shared = {"ids": ids, "lens": lens}
            cache[key] = shared
        ids, lens = shared["ids"], shared["lens"]
        for t in self.tables:
            t._tag()
        out = _FusedLookup.apply(self._group, layout, ids, lens, B, shared, *[t.weight for t in self.tables])
        return out.view(B, len(self.columns), self.emb_size)

    def check_index_errors(self) -> None:
        if self._group is not None:
            self._group.check_index_errors()

    def lookup_attention(self, cand: Dict[str, Tensor], hist: Dict[str, Tensor], lens: Optional[Tensor], attention):
        """DIN: ``seq = self({cand_f || hist_f}) -> q = seq[:, 0], keys = seq[:, 1:]; pooled = attention(q, keys, lens)``
        with the history rows read by K4 straight from the tables (``_FusedLookupAttention``) where that build exists
        (two fp32 tables of width D with 2D = 32, tensor-core forward and backward; ``PTREC_DIN_FUSED_GATHER=0``
        switches it off), else the materialised lookup.  ``cand``: name -> [B] ids, ``hist``: name -> [B, L] ids (this
        module's column names).  Returns (q [B, F*D], pooled [B, F*D])."""
        names = list(self.feature_names)
        first = cand[names[0]]
        B, L = hist[names[0]].shape
        flat = {n: torch.cat([cand[n].unsqueeze(1), hist[n]], dim=1).reshape(-1) for n in names}
        fc = (attention.fc1, attention.fc2, attention.fc3)
        fused = (first.is_cuda and len(names) == 2 and self.emb_size * 2 == 32 and len(set(self._table_of)) == 2
                 and all(t.weight.dtype == torch.float32 for t in self.tables)
                 and ops.din_ids_supported(2 * self.emb_size, fc[0].out_features, fc[1].out_features)
                 and os.environ.get("PTREC_DIN_FUSED_GATHER", "1") != "0")
        if not fused:
            seq = self(flat).view(B, 1 + L, len(names) * self.emb_size)
            q, keys = seq[:, 0], seq[:, 1:]
            return q, attention(q, keys, lens)
        if self._group is None:
            self._group = EmbeddingGroup(self.tables, self.emb_size)
        layout = self._layout_for((1,) * len(names))
        ids = torch.cat([flat[names[i]].reshape(-1) for i in self._order]).long().contiguous()
        shared = {"ids": ids, "lens": None}
        for t in self.tables:
            t._tag()
        blocks = tuple((self._order.index(i), self._table_of[i]) for i in range(len(names)))
        if lens is not None:
            lens = lens.to(torch.int32).contiguous()
        return _FusedLookupAttention.apply(self._group, layout, ids, shared, B, L, lens, blocks,
                                           fc[0].weight, fc[0].bias, fc[1].weight, fc[1].bias, fc[2].weight, fc[2].bias,
                                           *[t.weight for t in self.tables])
