"""Row-wise sharded embedding tables across the GPUs of one node (C1) + dense-tower gradient
allreduce (C2).  One process per GPU (torchrun); NCCL over NVLink / NVSwitch through
``torch.distributed``.  The reference is single-device (torchrec/task/Task.py:187-190): nothing here
has a reference counterpart, it widens the same hot path to the 8 x B200 box.

Plan:  owner(id) = id mod G,  local_row = id div G.

Push path (default on NCCL process groups; PTREC_EXCHANGE=push): nothing but small id lists and whole rows crosses
NVLink, every random access stays on the GPU that owns the row, and no NCCL collective is issued:
  forward   pack ids (+ the sample index of each lookup) by owner, stored straight into the owners' lists -> barrier
            -> every OWNER gathers the rows of the lists it received and stores each one straight into the
            requester's output row (``ptrec_gather_push``: the fused form of gather -> all-to-all -> gather by slot)
            -> barrier.  The owner-side sort / dedup of the received lists runs on a side stream meanwhile.
  backward  as the pull path below: gradient rows stored into the owners' buffers -> barrier -> fused update.
  dense     as below (all-reduce fused into the K7 optimizer launch).

Pull path (PTREC_EXCHANGE=pull or ``peer=True``; NCCL process groups while the shards stay below PTREC_PEER_MAX_GB = 16 GB per GPU;
``peer=True/False`` or PTREC_PEER_GATHER=1/0 force one or the other):
the shards and the owners' receive buffers live in symmetric memory, every rank holds the peers' pointers, and
the exchange happens INSIDE the kernels over NVLink / NVSwitch:
  forward   ONE gather launch per width; its 128-bit row loads go to whichever GPU owns the row  (no collective).
            Meanwhile, on a side stream: pack ids by owner, stored straight into the owners' lists -> barrier ->
            owner-side sort / dedup of the received lists — all of it overlaps the forward and the dense tower.
  backward  scatter gradient rows straight into the owners' buffers -> barrier ("every push has landed")
            -> owner-side segment-sum + fused optimizer update (joins the early sort).
  dense     the replicated tower's gradients are packed into a symmetric stage, one barrier, and the optimizer
            kernel (K7) sums every rank's stage itself: the all-reduce is fused into the update.  That barrier is also
            the fence that orders this step's table updates and buffer reuse before the next step's peer accesses.
The barriers are kernels of this library on symmetric-memory flags (csrc/peer_sync.cu, ``ops.PeerSync``), not NCCL
launches: a step of the peer path issues no NCCL collective at all (PTREC_PEER_SYNC=0 restores the NCCL fence and
all_reduce; PTREC_EARLY_EXCHANGE=0 keeps the id exchange and the sort in the backward).

All-to-all path (NCCL only; any backend that offers all_to_all_single):
  forward   pack ids by owner (fixed-capacity lists, no host sync)  -> all_to_all(ids)
            -> owner-side fused gather straight into the return layout -> all_to_all(rows)
            -> local gather by slot  -> [B, F, D]
  backward  scatter gradient rows into the send layout -> all_to_all(grads)
            -> owner-side sort / dedup / segment-sum / fused optimizer update (no gradient returns)
"""
import copy
import math
import os
import weakref
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor, nn

from .. import ops
from ..feature_column import CategoricalColumn
from ..model.ctr import DeepFM
from ..model.layer.embedding import EmbeddingGroup, EmbeddingTable


def shard_rows(total_rows: int, world: int, rank: int) -> int:
    """Number of rows r in [0, total_rows) with r mod world == rank."""
    return (total_rows - rank + world - 1) // world if total_rows > rank else 0


def owner_share(category_nums: Sequence[int], world: int) -> float:
    """Largest fraction of one field's lookups that a single owner can expect under ``owner = id mod G`` with ids
    uniform over the field's categories: ``max_f ceil(R_f / G) / R_f``.  ``1 / G`` for tall tables; a field with
    fewer categories than ranks sends ``1 / R_f`` of the batch to each of its ``R_f`` owners (Criteo has fields with
    3-10 categories)."""
    share = 1.0 / world
    for r in category_nums:
        r = max(int(r), 1)
        share = max(share, ((r + world - 1) // world) / r)
    return share


def list_capacity(batch: int, world: int, factor: float = 1.25, share: Optional[float] = None) -> int:
    """Slots per (owner, field) list: the expected ``batch * share`` (``share`` = ``owner_share`` of the columns,
    ``1 / world`` when not given) plus slack, a multiple of 16, never more than the batch (which cannot overflow).
    Hot ids can still exceed it (every duplicate of an id lands on the same owner): the pack kernels then raise the
    overflow word, which ``RowWiseShardedEmbedding`` turns into a RuntimeError (``poll_errors`` / ``check_errors``)."""
    share = 1.0 / world if share is None else share
    c = int(math.ceil(batch * share * factor)) + 64
    return min((c + 15) // 16 * 16, (batch + 15) // 16 * 16)


_PEER_MODULES: "weakref.WeakSet" = weakref.WeakSet()  # RowWiseShardedEmbedding instances on the peer-memory path


def _mark_fenced(group) -> None:
    """A collective over ``group`` was just enqueued on the current stream: everything the ranks enqueued before
    it (table updates, receive-buffer resets) is ordered before whatever any rank enqueues after it."""
    for m in _PEER_MODULES:
        if (m.group or dist.group.WORLD) is (group or dist.group.WORLD):
            m._dirty = False


def allreduce_dense_grads(params: Sequence[Tensor], group=None) -> None:
    """Average ``.grad`` of the replicated dense parameters over the ranks with ONE flat all_reduce."""
    grads = [p.grad for p in params if p.grad is not None and not p.grad.is_sparse]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    _mark_fenced(group)
    flat.div_(dist.get_world_size(group))
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


class _ShardedLookup(torch.autograd.Function):
    """One exchange for every embedding width of the same fields: ids travel once, the rows of all widths share one
    all-to-all (interleaved per slot), the gradients likewise, and the owner sorts the received lookups once."""

    @staticmethod
    def forward(ctx, mod: "RowWiseShardedEmbedding", ids: Tensor, *weights):
        F, B = ids.shape
        G, dev = mod.world, ids.device
        C = mod.capacity(B)
        send_ids, ret_pos = ops.a2a_pack_by_owner(ids, F, B, G, C, mod.overflow_flag(dev))
        mod.publish_overflow(dev)
        recv_ids = torch.empty_like(send_ids)
        dist.all_to_all_single(recv_ids, send_ids, group=mod.group)            # [G_src, F, C]
        own_ids = recv_ids.permute(1, 0, 2).contiguous().view(-1)              # [F, G_src, C]
        bufs = mod.buffers(C, dev)  # persistent: fixed addresses (pointer arrays built once, graph-capturable)
        rows_out, recv_rows = bufs["rows_out"], bufs["recv_rows"]             # [G*F*C, S]: every width of a slot side by side
        S = mod.slot_width
        n_t = len(mod.columns)
        for k, D in enumerate(mod.dims):
            tables = mod.egroups[k].table_set.refresh([w.detach() for w in weights[k * n_t:(k + 1) * n_t]])
            ops.gather_pool_fwd(tables, mod.owner_layout(C, k), own_ids, None, C, out=rows_out,
                                err_flag=mod.egroups[k].err_flag(dev), out_row_stride=S)
        dist.all_to_all_single(recv_rows, rows_out, group=mod.group)           # [G_owner, F, C, S]
        outs = []
        pos64 = ret_pos.view(-1).long()
        for k, D in enumerate(mod.dims):
            # local gather by slot: out[b, f] = recv_rows[ret_pos[f, b], col_k : col_k + D]
            out, _ = ops.gather_pool_fwd(bufs["slot_tables"][k], mod.slot_layout(k), pos64, None, B)
            outs.append(out.view(B, F, D))
        ctx.mod, ctx.C, ctx.shape = mod, C, (F, B)
        ctx.save_for_backward(own_ids, ret_pos)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        mod, C = ctx.mod, ctx.C
        F, B = ctx.shape
        G, S = mod.world, mod.slot_width
        own_ids, ret_pos = ctx.saved_tensors
        dev = own_ids.device
        bufs = mod.buffers(C, dev)
        send_g, recv_g = bufs["send_g"], bufs["recv_g"]
        # slots that carry no lookup are -1 on the owner (masked in the sort): no need to clear them
        for k, D in enumerate(mod.dims):
            g = grads[k]
            if g is None:
                send_g[:, mod.col_of[k]:mod.col_of[k] + D].zero_()
                continue
            g = g.reshape(B, F * D)
            if not g.is_contiguous():
                g = g.contiguous()
            ops.a2a_scatter_rows(g, ret_pos, B, F, D, mod.grad_scale, send_g[:, mod.col_of[k]:mod.col_of[k] + D])
        dist.all_to_all_single(recv_g, send_g, group=mod.group)                # [G_src, F, C, S]
        srt = None
        for k, D in enumerate(mod.dims):
            eg = mod.egroups[k]
            bind = eg.binding()
            if bind is None:
                raise RuntimeError("row-wise sharded tables need a pytorchrec_b200.optim sparse optimizer")
            s1, s2, args = bind[0]._fused_prepare(eg, bind[1])  # may interleave weight | state: before taking pointers
            tables = eg.table_set.refresh([t.weight.data for t in eg.tables])
            layout = mod.owner_layout(C, k)
            if srt is None:
                srt = ops.sort_dedup(tables, layout, own_ids, None, C)  # same ids, same shard heights for every width
            ops.bwd_fused(tables, s1, s2, layout, C, srt, recv_g, None, args, grad_row_stride=S)
        return (None, None) + (None,) * (len(mod.dims) * len(mod.columns))


_SIDE: Dict[torch.device, "torch.cuda.Stream"] = {}


def _side_stream(dev) -> "torch.cuda.Stream":
    st = _SIDE.get(dev)
    if st is None:
        st = _SIDE[dev] = torch.cuda.Stream(dev)
    return st


_AUX: Dict[torch.device, "torch.cuda.Stream"] = {}


def _aux_stream(dev) -> "torch.cuda.Stream":
    """Second compute stream: the narrow widths of a slot (the first-order column: 4-byte rows, latency-bound) are
    gathered and updated beside the wide one instead of behind it (PTREC_AUX_STREAM=0: same stream)."""
    st = _AUX.get(dev)
    if st is None:
        st = _AUX[dev] = torch.cuda.Stream(dev)
    return st


def _aux_enabled() -> bool:
    import os
    return os.environ.get("PTREC_AUX_STREAM", "1") != "0"


def _owner_updates(mod, C, srt, recv_g, dev) -> None:
    """Owner-side segment-sum + fused optimizer update of every width; widths after the first on the aux stream."""
    main = torch.cuda.current_stream(dev)
    use_aux = _aux_enabled() and len(mod.dims) > 1
    aux = _aux_stream(dev) if use_aux else None
    prepared = []
    for k, D in enumerate(mod.dims):
        eg = mod.egroups[k]
        bind = eg.binding()
        if bind is None:
            raise RuntimeError("row-wise sharded tables need a pytorchrec_b200.optim sparse optimizer")
        s1, s2, args = bind[0]._fused_prepare(eg, bind[1])  # may interleave weight | state: before taking pointers
        tables = eg.table_set.refresh([t.weight.data for t in eg.tables])
        prepared.append((tables, s1, s2, args))
    if use_aux:
        aux.wait_stream(main)   # gradient rows landed (barrier), sort joined
        for t in (srt.sorted_keys, srt.perm, srt.seg_start, srt.seg_meta, srt.n_seg):
            t.record_stream(aux)   # allocated on another stream: not to be recycled before the aux update has read them
    for k, (tables, s1, s2, args) in enumerate(prepared):
        if use_aux and k > 0:
            with torch.cuda.stream(aux):
                ops.bwd_fused(tables, s1, s2, mod.owner_layout(C, k), C, srt, recv_g, None, args, grad_row_stride=mod.slot_width)
        else:
            ops.bwd_fused(tables, s1, s2, mod.owner_layout(C, k), C, srt, recv_g, None, args, grad_row_stride=mod.slot_width)
    if use_aux:
        main.wait_stream(aux)


class _PeerLookup(torch.autograd.Function):
    """The same lookup with the exchange inside the kernels (peer loads / stores over NVLink), see module docstring."""

    @staticmethod
    def _exchange_ids(mod, ids, F, B, dev):
        """pack ids by owner into the owners' lists -> barrier -> owner-side sort / dedup of the lists received here.
        Runs on whatever stream is current.  Returns (ret_pos, sort result, capacity)."""
        C = mod.capacity(B)
        pb = mod.peer_buffers(C, dev)
        ret_pos = ops.a2a_pack_by_owner_peer(ids, F, B, mod.world, C, mod.rank, pb["peer_ids"], mod.overflow_flag(dev),
                                             slot_b=pb["slot_b"] if mod.ordered_scatter else None)
        mod.publish_overflow(dev)
        mod.barrier(ops.PeerSync.IDS, dev)  # every rank's lists have landed in every owner's buffer
        eg = mod.egroups[0]                 # same ids and the same shard heights for every width: one sort
        tables = eg.table_set.refresh([t.weight.data for t in eg.tables])
        srt = ops.sort_dedup(tables, mod.owner_layout(C, 0), pb["recv_ids"], None, C)
        return ret_pos, srt, C

    @staticmethod
    def forward(ctx, mod: "RowWiseShardedEmbedding", ids: Tensor, *weights):
        F, B = ids.shape
        dev = ids.device
        if mod._dirty:
            mod.fence(dev)
        ctx.early = None
        if mod.early_exchange and any(ctx.needs_input_grad[2:]):
            # the id exchange and the owner-side sort depend on the ids only: start them now on a side stream, where
            # they overlap the gather, the interaction and the dense tower; the backward joins them with an event
            main, side = torch.cuda.current_stream(dev), _side_stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ret_pos, srt, C = _PeerLookup._exchange_ids(mod, ids, F, B, dev)
                ev = torch.cuda.Event()
                ev.record(side)
            ids.record_stream(side)
            for t in (ret_pos, srt.sorted_keys, srt.perm, srt.seg_start, srt.seg_meta, srt.n_seg):
                t.record_stream(main)  # consumed on the main stream in backward
            ctx.early = (ret_pos, srt, C, ev)
        flat = ids.reshape(-1)
        outs = []
        main = torch.cuda.current_stream(dev)
        use_aux = _aux_enabled() and len(mod.dims) > 1
        if use_aux:
            aux = _aux_stream(dev)
            aux.wait_stream(main)
        for k, D in enumerate(mod.dims):
            if use_aux and k > 0:
                with torch.cuda.stream(aux):
                    out = ops.gather_fwd_sharded(mod._shard_ptrs[k], mod._rows_global, mod.world, mod._row_stride[k],
                                                 mod.plain_layout(k), flat, B, err_flag=mod.egroups[k].err_flag(dev))
                out.record_stream(main)
            else:
                out = ops.gather_fwd_sharded(mod._shard_ptrs[k], mod._rows_global, mod.world, mod._row_stride[k],
                                             mod.plain_layout(k), flat, B, err_flag=mod.egroups[k].err_flag(dev))
            outs.append(out.view(B, F, D))
        if use_aux:
            main.wait_stream(aux)
        ctx.mod, ctx.shape = mod, (F, B)
        ctx.save_for_backward(ids)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        mod = ctx.mod
        F, B = ctx.shape
        G, S = mod.world, mod.slot_width
        (ids,) = ctx.saved_tensors
        dev = ids.device
        if ctx.early is not None:
            ret_pos, srt, C, ev = ctx.early
            torch.cuda.current_stream(dev).wait_event(ev)  # join the early exchange + sort
        else:
            if mod._dirty:  # the owners may still be consuming their buffers from an earlier backward
                mod.fence(dev)
            ret_pos, srt, C = _PeerLookup._exchange_ids(mod, ids, F, B, dev)
        pb = mod.peer_buffers(C, dev)
        gs = []
        for k, D in enumerate(mod.dims):
            g = grads[k]
            g = torch.zeros(B, F * D, device=dev) if g is None else g.reshape(B, F * D)
            gs.append(g if g.is_contiguous() else g.contiguous())
        mod.scatter_grads(gs, ret_pos, pb, B, F, C)
        mod.barrier(ops.PeerSync.GRADS, dev)  # every rank's gradient rows have landed in every owner's buffer
        _owner_updates(mod, C, srt, pb["recv_g"], dev)
        mod._dirty = True   # tables changed, buffers recycled: a fence must precede the next peer access
        return (None, None) + (None,) * (len(mod.dims) * len(mod.columns))


class _PushLookup(torch.autograd.Function):
    """Push-mode lookup (module docstring): owners deliver rows into the requesters' outputs over NVLink."""

    @staticmethod
    def forward(ctx, mod: "RowWiseShardedEmbedding", ids: Tensor, *weights):
        F, B = ids.shape
        dev = ids.device
        G = mod.world
        if mod._dirty:
            mod.fence(dev)
        C = mod.capacity(B)
        pb = mod.peer_buffers(C, dev)
        po = mod.push_outputs(B, dev)
        ret_pos = ops.a2a_pack_by_owner_push(ids, F, B, G, C, mod.rank, pb["peer_ids"], pb["peer_b"], po["outs"], mod.dims,
                                             mod.overflow_flag(dev), slot_b=pb["slot_b"] if mod.ordered_scatter else None)
        mod.publish_overflow(dev)
        mod.barrier(ops.PeerSync.IDS, dev)   # every rank's lists have landed in every owner's buffers
        n_t = len(mod.columns)
        tabs = [mod.egroups[k].table_set.refresh([w.detach() for w in weights[k * n_t:(k + 1) * n_t]])
                for k in range(len(mod.dims))]
        ctx.early = None
        if any(ctx.needs_input_grad[2:]):
            # the owner-side sort depends on the received ids only: side stream, joined by the backward
            main, side = torch.cuda.current_stream(dev), _side_stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                srt = ops.sort_dedup(tabs[0], mod.owner_layout(C, 0), pb["recv_ids"], None, C)
                ev = torch.cuda.Event()
                ev.record(side)
            for t in (srt.sorted_keys, srt.perm, srt.seg_start, srt.seg_meta, srt.n_seg):
                t.record_stream(main)
            ctx.early = (srt, ev)
        ops.gather_push([t.ptrs for t in tabs], po["peer_outs"], [t.row_stride or d for t, d in zip(tabs, mod.dims)],
                        [o.stride(0) for o in po["outs"]], mod.dims, pb["recv_ids"], pb["recv_b"], tabs[0].rows, F, G, C,
                        err_flag=mod.egroups[0].err_flag(dev))
        mod.barrier(ops.PeerSync.ROWS, dev)  # every owner's rows have landed in this rank's outputs
        ctx.mod, ctx.shape, ctx.C = mod, (F, B), C
        ctx.save_for_backward(ret_pos)
        return tuple(o.view(B, F, d) for o, d in zip(po["outs"], mod.dims))

    @staticmethod
    def backward(ctx, *grads):
        mod, C = ctx.mod, ctx.C
        F, B = ctx.shape
        G, S = mod.world, mod.slot_width
        (ret_pos,) = ctx.saved_tensors
        dev = ret_pos.device
        pb = mod.peer_buffers(C, dev)
        gs = []
        for k, D in enumerate(mod.dims):
            g = grads[k]
            g = torch.zeros(B, F * D, device=dev) if g is None else g.reshape(B, F * D)
            gs.append(g if g.is_contiguous() else g.contiguous())
        mod.scatter_grads(gs, ret_pos, pb, B, F, C)
        mod.barrier(ops.PeerSync.GRADS, dev)  # every rank's gradient rows have landed in every owner's buffer
        srt, ev = ctx.early
        torch.cuda.current_stream(dev).wait_event(ev)  # join the early sort
        _owner_updates(mod, C, srt, pb["recv_g"], dev)
        mod._dirty = True   # tables changed, buffers recycled: a fence must precede the next exchange
        return (None, None) + (None,) * (len(mod.dims) * len(mod.columns))


class _DenseReducer:
    """What a fused optimizer needs to average the replicated dense gradients over NVLink peer memory
    (``optim/sparse.py::_fused_dense_step``): per-launch stages in symmetric memory, one barrier, a fallback."""

    def __init__(self, mod: "RowWiseShardedEmbedding"):
        self.mod = mod
        self.world = mod.world
        self._stages: Dict[tuple, tuple] = {}

    def stage(self, key, n_floats: int, device):
        rec = self._stages.get(key)
        if rec is None or rec[0].numel() < n_floats:
            import torch.distributed._symmetric_memory as symm
            group = self.mod.group or dist.group.WORLD
            buf = symm.empty(max(n_floats, 1024), dtype=torch.float32, device=device)
            buf.zero_()
            hdl = symm.rendezvous(buf, group)
            ptrs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64).to(device)
            self.mod.sync_peers()
            rec = self._stages[key] = (buf, ptrs, hdl)
        return rec[0], rec[1]

    def barrier(self) -> None:
        dev = torch.device("cuda", torch.cuda.current_device())
        self.mod.barrier(ops.PeerSync.DENSE, dev)

    def fallback(self, params) -> None:
        allreduce_dense_grads(params, self.mod.group)


class RowWiseShardedEmbedding(nn.Module):
    """``MultiTableEmbedding`` whose tables are sharded row-wise over the process group.  ``emb_size`` may be a list
    (e.g. ``[16, 1]`` for DeepFM's embedding and first-order tables): every width gets its own table per column
    (``groups[k][f]`` = LOCAL shard ``EmbeddingTable(shard_rows(category_num), dim_k)``) and all widths share one
    id exchange, one row exchange and one gradient exchange per step.  One-hot fields only (``[B]`` ids per
    column) — the Criteo-shaped configs; pooled bags stay on ``MultiTableEmbedding``."""

    def __init__(self, columns: Sequence[CategoricalColumn], emb_size, group=None,
                 capacity_factor: float = 1.25, device=None, peer: Optional[bool] = None):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("RowWiseShardedEmbedding needs an initialised torch.distributed process group")
        self.columns = list(columns)
        self.single = isinstance(emb_size, int)
        self.dims = [int(emb_size)] if self.single else [int(d) for d in emb_size]
        self.col_of, w = [], 0
        for d in self.dims:  # column of each width inside a slot, 16-byte aligned
            self.col_of.append(w)
            w += (d + 3) // 4 * 4
        self.slot_width = w
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.capacity_factor = capacity_factor
        # Exchange lists are sized for the fullest owner of the lowest-cardinality field (ids uniform over the
        # field's categories); a field with very few categories makes EVERY list that long, so it is better kept out
        # of the sharded module (replicated) — say so once.
        self.owner_share = owner_share([c.category_num for c in self.columns], self.world)
        if self.owner_share > 2.0 / self.world:
            import warnings
            small = [getattr(c, "feature_name", "?") for c in self.columns if c.category_num < 2 * self.world]
            warnings.warn(f"row-wise sharding over {self.world} ranks: columns {small} have fewer categories than "
                          f"2 x ranks, so every exchange list is sized for {self.owner_share:.0%} of the batch instead "
                          f"of {1.0 / self.world:.0%}; replicate such tiny tables instead of sharding them")
        self._overflow_host: Dict[torch.device, Tensor] = {}
        self.grad_scale = 1.0 / self.world  # mean loss over the GLOBAL batch
        if self.world * len(self.columns) > 256:
            raise ValueError("world_size * n_fields must be <= 256")
        self.groups = nn.ModuleList([
            nn.ModuleList([EmbeddingTable(max(shard_rows(c.category_num, self.world, self.rank), 1), d, device=device)
                           for c in self.columns]) for d in self.dims])
        self.egroups: List[EmbeddingGroup] = []
        self._owner_layouts: Dict[tuple, ops.FeatureLayout] = {}
        self._slot_layouts: Dict[int, ops.FeatureLayout] = {}
        self._overflow: Dict[torch.device, Tensor] = {}
        self._bufs: Dict[tuple, dict] = {}
        # ---- peer-memory path state ----
        # exchange mode: "push" (owners deliver rows; default on NCCL groups), "pull" (requesters read the owners'
        # shards in symmetric memory: the round-1 peer path; ``peer=True`` / PTREC_PEER_GATHER=1), "a2a" (NCCL
        # all-to-all; any backend; ``peer=False`` / PTREC_PEER_GATHER=0)
        nccl = dist.get_backend(group) == "nccl"
        mode = os.environ.get("PTREC_EXCHANGE", "auto")
        legacy = os.environ.get("PTREC_PEER_GATHER", "auto")
        if peer is not None:
            mode = "pull" if peer else "a2a"
        elif mode == "auto" and legacy in ("0", "1"):
            mode = "pull" if legacy == "1" else "a2a"
        elif mode == "auto":
            # Small shards: PULL (no barrier on the forward's critical path; measured at cfg2, 2 x B200: 0.845 ms / step
            # against 0.876 for push).  Large shards: PUSH — random row reads over 7 x 83 GB of peer mappings ran
            # 10.1 ms / step at cfg5 on 8 GPUs (round 1), while push keeps every random access local (cfg5, 2 GPUs:
            # 4.71 ms against 5.74 for NCCL all-to-all).  The estimate uses the rank-independent shard height so that
            # every rank takes the same decision.
            cap_gb = float(os.environ.get("PTREC_PEER_MAX_GB", "16"))
            per_gpu = sum((c.category_num + self.world - 1) // self.world for c in self.columns) * sum(self.dims) * 8
            if nccl and os.environ.get("PTREC_PEER_SYNC", "1") != "0":
                mode = "pull" if per_gpu <= cap_gb * 2 ** 30 else "push"
            else:
                mode = "a2a"
        if mode not in ("push", "pull", "a2a"):
            raise ValueError(f"PTREC_EXCHANGE must be push, pull, a2a or auto, got {mode!r}")
        if mode != "a2a" and not nccl:
            raise RuntimeError(f"exchange mode {mode!r} needs an NCCL process group (symmetric memory over NVLink)")
        self.exchange = mode
        peer = mode == "pull"
        self.peer = bool(peer)
        # cross-rank ordering by this library's barrier kernels on symmetric-memory flags (default) or by NCCL
        # (also serves the all-to-all path's dense-gradient reduction: any NCCL group can hold symmetric memory)
        self.use_peer_sync = dist.get_backend(group) == "nccl" and os.environ.get("PTREC_PEER_SYNC", "1") != "0"
        # NCCL collectives from a side stream would have to keep one issue order on every rank: early exchange only
        # with the barrier kernels
        self.early_exchange = self.peer and self.use_peer_sync and os.environ.get("PTREC_EARLY_EXCHANGE", "1") != "0"
        if self.exchange == "push" and not self.use_peer_sync:
            raise RuntimeError("the push exchange orders its stores with the barrier kernels: PTREC_PEER_SYNC=0 is not compatible")
        self._push_out: Dict[tuple, dict] = {}
        self.ordered_scatter = os.environ.get("PTREC_ORDERED_SCATTER", "1") != "0" and self.slot_width <= 128
        self._sync: Optional[ops.PeerSync] = None
        self._reducer: Optional[_DenseReducer] = None
        self._dirty = False                        # a table update / buffer reset not yet ordered by a collective
        self._peer_key = None                      # (data_ptr, row stride) of every table when pointers were exchanged
        self._symm: Dict[int, Tensor] = {}         # data_ptr -> symmetric-memory allocation that backs a table
        self._symm_handles: list = []
        self._shard_ptrs: List[Tensor] = []        # per width: int64 [T, G] shard base addresses seen from this GPU
        self._row_stride: List[int] = []
        self._rows_global: Optional[Tensor] = None
        self._plain_layouts: Dict[int, ops.FeatureLayout] = {}
        self._peer_bufs: Dict[tuple, dict] = {}
        self._fence_buf: Optional[Tensor] = None
        if self.exchange != "a2a":
            _PEER_MODULES.add(self)

    # ---- peer-memory path --------------------------------------------------------------------------------------
    def peer_sync(self, device) -> ops.PeerSync:
        if self._sync is None:  # collective on first use (symmetric allocation + rendezvous)
            self._sync = ops.PeerSync(self.group, device)
        return self._sync

    def barrier(self, slot: int, device) -> None:
        """Orders everything every rank enqueued before it (on the stream that is current there) ahead of everything
        any rank enqueues after it: one tiny kernel of this library (``ops.PeerSync``), or — PTREC_PEER_SYNC=0 — a
        1-element NCCL all_reduce."""
        if self.use_peer_sync:
            self.peer_sync(device).barrier(slot)
        else:
            if self._fence_buf is None:
                self._fence_buf = torch.zeros(1, dtype=torch.float32, device=device)
            dist.all_reduce(self._fence_buf, group=self.group)
        if slot in (ops.PeerSync.FENCE, ops.PeerSync.DENSE):  # issued on the main stream after the step's updates
            _mark_fenced(self.group)

    def fence(self, device) -> None:
        self.barrier(ops.PeerSync.FENCE, device)

    def dense_reducer(self) -> Optional["_DenseReducer"]:
        """The peer all-reduce hooks for the fused optimizers, or None when ordering goes through NCCL."""
        if not self.use_peer_sync:
            return None
        if self._reducer is None:
            self._reducer = _DenseReducer(self)
        return self._reducer

    def sync_peers(self) -> None:
        """Host-synchronising fence: call (on every rank) after writing table rows outside the training step."""
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self._dirty = False

    def _symm_alloc(self, cap_rows: int):
        import torch.distributed._symmetric_memory as symm

        def alloc(rows: int, width: int) -> Tensor:
            dev = torch.device("cuda", torch.cuda.current_device())
            buf = symm.empty(max(cap_rows, rows), width, dtype=torch.float32, device=dev)  # same size on all ranks
            self._symm[buf.data_ptr()] = buf
            return buf[:rows]
        return alloc

    def _ensure_peer(self, device) -> None:
        """(Re)exchange the shard pointers when a table's storage moved (first use; the optimizer interleaving
        weight | state at its first step; a loaded optimizer state).  Collective, and identical on every rank
        because the same events move the same tables everywhere."""
        key = tuple((t.weight.data_ptr(), t.weight.stride(0)) for t in self.tables)
        if key == self._peer_key:
            return
        import torch.distributed._symmetric_memory as symm
        group = self.group or dist.group.WORLD
        G = self.world
        self._shard_ptrs, self._row_stride, handles = [], [], []
        for g, D in zip(self.groups, self.dims):
            ptrs = []
            strides = set()
            for col, t in zip(self.columns, g):
                w = t.weight
                buf = self._symm.get(w.data_ptr())
                if buf is None:  # plain allocation: move the rows into symmetric memory
                    if w.stride(0) != w.shape[1]:
                        raise RuntimeError("a sharded table was interleaved outside symmetric memory")
                    view = w._ptrec_alloc(w.shape[0], w.shape[1])
                    view.copy_(w.data)
                    w.data = view
                    buf = self._symm[w.data_ptr()]
                hdl = symm.rendezvous(buf, group)
                handles.append(hdl)
                ptrs.append([int(p) for p in hdl.buffer_ptrs])
                strides.add(int(w.stride(0)))
            if len(strides) != 1:
                raise RuntimeError(f"tables of one width must share a row stride, got {sorted(strides)}")
            self._shard_ptrs.append(torch.tensor(ptrs, dtype=torch.int64).to(device))
            self._row_stride.append(strides.pop())
        live = {t.weight.data_ptr() for t in self.tables}
        self._symm = {p: b for p, b in self._symm.items() if p in live}
        self._symm_handles = handles
        self._rows_global = torch.tensor([c.category_num for c in self.columns], dtype=torch.int64).to(device)
        self._peer_key = tuple((t.weight.data_ptr(), t.weight.stride(0)) for t in self.tables)
        for eg in self.egroups:  # pointer arrays cached on the local tables are stale too
            eg.table_set.refresh([t.weight.data for t in eg.tables])
        self.sync_peers()  # the copies above are complete on every rank before anyone reads a peer

    def plain_layout(self, k: int) -> ops.FeatureLayout:
        lay = self._plain_layouts.get(k)
        if lay is None:
            F = len(self.columns)
            lay = ops.FeatureLayout([dict(table=f, bag_len=1, neg_is_pad=True) for f in range(F)], self.dims[k], F)
            self._plain_layouts[k] = lay
        return lay

    def peer_buffers(self, C: int, device) -> dict:
        """Owner-side receive buffers in symmetric memory + the peers' addresses of theirs (collective on first use
        of a capacity)."""
        key = (C, device)
        b = self._peer_bufs.get(key)
        if b is None:
            import torch.distributed._symmetric_memory as symm
            group = self.group or dist.group.WORLD
            F, G = len(self.columns), self.world
            recv_ids = symm.empty(F * G * C, dtype=torch.int64, device=device)
            recv_g = symm.empty(G * F * C, self.slot_width, dtype=torch.float32, device=device)
            recv_b = symm.empty(F * G * C, dtype=torch.int32, device=device)  # push mode: sample index of each lookup
            recv_b.zero_()
            h_b = symm.rendezvous(recv_b, group)
            recv_ids.fill_(-1)  # ranks that have not pushed yet read "no lookup"; every later step rewrites the lists in full
            recv_g.zero_()
            h_ids, h_g = symm.rendezvous(recv_ids, group), symm.rendezvous(recv_g, group)
            b = {"recv_ids": recv_ids, "recv_g": recv_g, "recv_b": recv_b, "handles": (h_ids, h_g, h_b),
                 "slot_b": torch.full((G * F * C,), -1, dtype=torch.int32, device=device),  # local: sample behind each slot
                 "peer_b": torch.tensor([int(p) for p in h_b.buffer_ptrs], dtype=torch.int64).to(device),
                 "peer_ids": torch.tensor([int(p) for p in h_ids.buffer_ptrs], dtype=torch.int64).to(device),
                 "peer_g": torch.tensor([int(p) for p in h_g.buffer_ptrs], dtype=torch.int64).to(device)}
            self._peer_bufs[key] = b
            self.sync_peers()
        return b

    def scatter_grads(self, gs, ret_pos: Tensor, pb: dict, B: int, F: int, C: int) -> None:
        """Gradient rows of every width -> the owners' receive buffers (NVLink stores): in destination order (long
        contiguous runs; default) or in batch order (PTREC_ORDERED_SCATTER=0)."""
        if self.ordered_scatter:
            ops.a2a_scatter_rows_peer_ordered(gs, self.dims, self.col_of, pb["slot_b"], F, self.grad_scale, pb["peer_g"],
                                              self.slot_width, C, self.world, self.rank)
        else:
            ops.a2a_scatter_rows_peer_multi(gs, self.dims, self.col_of, ret_pos, B, F, self.grad_scale, pb["peer_g"],
                                            self.slot_width, C, self.world, self.rank)

    def push_outputs(self, B: int, device) -> dict:
        """This rank's lookup outputs, one [B, F * D_k] buffer per width in symmetric memory (the owners store rows
        into them), + the device arrays of every rank's buffers (collective on first use of a batch size)."""
        key = (B, device)
        po = self._push_out.get(key)
        if po is None:
            import torch.distributed._symmetric_memory as symm
            group = self.group or dist.group.WORLD
            F = len(self.columns)
            outs, peer_outs, handles = [], [], []
            for d in self.dims:
                o = symm.empty(B, F * d, dtype=torch.float32, device=device)
                o.zero_()
                h = symm.rendezvous(o, group)
                outs.append(o)
                handles.append(h)
                peer_outs.append(torch.tensor([int(p) for p in h.buffer_ptrs], dtype=torch.int64).to(device))
            po = self._push_out[key] = {"outs": outs, "peer_outs": peer_outs, "handles": handles}
            self.sync_peers()
        return po

    @property
    def emb_size(self) -> int:
        return self.dims[0]

    @property
    def weight(self) -> Tensor:  # see MultiTableEmbedding.weight: the container itself owns no parameter
        return torch.empty(0)

    @property
    def tables(self) -> List[EmbeddingTable]:
        return [t for g in self.groups for t in g]

    def buffers(self, C: int, device) -> dict:
        key = (C, device)
        b = self._bufs.get(key)
        if b is None:
            n = self.world * len(self.columns) * C
            mk = lambda: torch.zeros(n, self.slot_width, dtype=torch.float32, device=device)  # noqa: E731
            b = {"rows_out": mk(), "recv_rows": mk(), "send_g": mk(), "recv_g": mk()}
            b["slot_tables"] = [ops.TableSet().refresh([b["recv_rows"][:, c:c + d]]) for c, d in zip(self.col_of, self.dims)]
            self._bufs[key] = b
        return b

    def capacity(self, batch: int) -> int:
        """Slots per (owner, field) exchange list for a per-rank batch of ``batch`` samples."""
        return list_capacity(batch, self.world, self.capacity_factor, self.owner_share)

    def overflow_flag(self, device) -> Tensor:
        t = self._overflow.get(device)
        if t is None:
            t = torch.zeros(1, dtype=torch.int32, device=device)
            self._overflow[device] = t
        return t

    def publish_overflow(self, device) -> None:
        """Enqueue (or capture into the step graph) a copy of the sticky overflow word into pinned host memory, so
        that ``poll_errors`` can see it without synchronising."""
        pin = self._overflow_host.get(device)
        if pin is None:
            pin = self._overflow_host[device] = torch.zeros(1, dtype=torch.int32).pin_memory()
        pin.copy_(self.overflow_flag(device), non_blocking=True)

    def poll_errors(self) -> None:
        """Non-synchronising check, called at the top of every forward / train_step: raises once a pack launch of an
        EARLIER step is known to have overflowed a list (the word is sticky, so nothing is missed — the error
        surfaces one or two steps late instead of corrupting the run silently: a dropped lookup reads as a zero
        row in the forward and loses its gradient in the backward)."""
        for pin in self._overflow_host.values():
            v = int(pin[0])
            if v:
                raise RuntimeError(
                    f"a row-wise sharded lookup list overflowed its capacity (needed {v} slots): lookups were dropped "
                    "in an earlier step — results since then are invalid.  Raise capacity_factor (hot ids send all "
                    "their duplicates to one owner) or replicate low-cardinality columns")

    def owner_layout(self, C: int, k: int = 0) -> ops.FeatureLayout:
        """Owner-side view of the received lists for width k: feature (f, src) reads table f, batch = C, and writes
        its rows at slot ((src*F + f)*C + c), column col_of[k] — i.e. directly in the all-to-all return layout."""
        lay = self._owner_layouts.get((C, k))
        if lay is None:
            F, G, S = len(self.columns), self.world, self.slot_width
            specs = [dict(table=f, bag_len=1, neg_is_pad=True, out_col=(src * F + f) * C * S + self.col_of[k])
                     for f in range(F) for src in range(G)]
            lay = ops.FeatureLayout(specs, self.dims[k], F)
            self._owner_layouts[(C, k)] = lay
        return lay

    def slot_layout(self, k: int = 0) -> ops.FeatureLayout:
        lay = self._slot_layouts.get(k)
        if lay is None:
            lay = ops.FeatureLayout([dict(table=0, bag_len=1, neg_is_pad=True) for _ in self.columns], self.dims[k], 1)
            self._slot_layouts[k] = lay
        return lay

    def forward(self, batch: Dict[str, Tensor]):
        self.poll_errors()
        if not self.egroups:
            self.egroups = [EmbeddingGroup([t for t in g], d) for g, d in zip(self.groups, self.dims)]
        ids = torch.stack([c.get_feature_data(batch).reshape(-1) for c in self.columns])  # [F, B]
        weights = []
        for g in self.groups:
            for col, t in zip(self.columns, g):
                t._tag()
                if self.peer and getattr(t.weight, "_ptrec_alloc", None) is None:
                    t.weight._ptrec_alloc = self._symm_alloc((col.category_num + self.world - 1) // self.world)
                weights.append(t.weight)
        if self.exchange == "push":
            outs = _PushLookup.apply(self, ids, *weights)
        elif self.peer:
            self._ensure_peer(ids.device)
            outs = _PeerLookup.apply(self, ids, *weights)
        else:
            outs = _ShardedLookup.apply(self, ids, *weights)
        return outs[0] if self.single else list(outs)

    def check_index_errors(self) -> None:
        self.check_errors()

    def check_errors(self) -> None:
        """Synchronising check of the overflow / out-of-range flags."""
        for t in self._overflow.values():
            v = int(t.item())
            if v:
                t.zero_()
                for pin in self._overflow_host.values():
                    pin.zero_()
                raise RuntimeError(f"a row-wise sharded lookup list overflowed its capacity (needed {v} slots; lookups "
                                   "were dropped); raise capacity_factor or replicate low-cardinality columns")
        for eg in self.egroups:
            eg.check_index_errors()


class ShardedDeepFM(DeepFM):
    """DeepFM with the embedding (dim D) and first-order (dim 1) tables row-wise sharded behind ONE exchange, and the
    dense tower data-parallel.  ``sharded.groups[0][f]`` / ``sharded.groups[1][f]`` are the local shards."""

    def _init_weights(self):
        self.sharded = RowWiseShardedEmbedding(self.sparse_columns, [self.emb_size, 1], device=self.table_device)
        if self.dense_columns:
            self.dense_linear = nn.Linear(len(self.dense_columns), 1, bias=False)
        self.global_bias = nn.Parameter(torch.tensor(0.0))
        from ..model.layer import MLP, FMSecondOrder
        self.fm2 = FMSecondOrder()
        in_units = len(self.sparse_columns) * self.emb_size + len(self.dense_columns)
        self.mlp = MLP(input_units=in_units, hidden_units_list=self.layers, activation="relu", dropout=self.dropout)
        self.deep_out = nn.Linear(self.layers[-1], 1, bias=False)

    @property
    def embeddings(self):
        return self.sharded

    def forward(self, data: Dict[str, Tensor]):
        from ..model.ctr import _dense_matrix
        v, w = self.sharded(data)                       # [B, F, D], [B, F, 1]
        x = _dense_matrix(self.dense_columns, data)
        logit, deep_in = self._head(data, v, w, x, True)
        return logit + self._deep_logit(deep_in), self._target(data)

    # ---- N4: checkpoints interchangeable with the unsharded model -------------------------------------------------
    _GROUP_NAMES = ("embeddings", "first_order")

    @torch.no_grad()
    def full_state_dict(self) -> Dict[str, Tensor]:
        """Gather the row-wise shards (collective: call on every rank) into the ``state_dict`` of the UNSHARDED
        ``DeepFM`` (``embeddings.{f}.weight`` / ``first_order.{f}.weight`` with rows interleaved back as
        ``row = local_row * G + rank``), on the CPU; dense parameters are taken from this rank."""
        G = self.sharded.world
        out = {k: v.detach().cpu() for k, v in self.state_dict().items() if not k.startswith("sharded.")}
        for k, name in enumerate(self._GROUP_NAMES):
            for f, (col, table) in enumerate(zip(self.sparse_columns, self.sharded.groups[k])):
                n_own = shard_rows(col.category_num, G, self.sharded.rank)
                full = self._gather_rows(table.weight[:n_own], col.category_num)
                out[f"{name}.{f}.weight"] = full
        return out

    def _gather_rows(self, local: Tensor, category_num: int) -> Tensor:
        """[local_rows, ...] shards of every rank -> the full [category_num, ...] tensor (row = local_row * G + rank)."""
        G = self.sharded.world
        local = local.detach().contiguous()
        cap = (category_num + G - 1) // G
        padded = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[:min(local.shape[0], cap)] = local[:cap]
        parts = [torch.empty_like(padded) for _ in range(G)]
        dist.all_gather(parts, padded, group=self.sharded.group)
        return torch.stack(parts, dim=1).reshape((cap * G,) + tuple(local.shape[1:]))[:category_num].cpu()

    @torch.no_grad()
    def full_optimizer_state_dict(self) -> Dict:
        """N4 ("incl. optimizer state"; the reference saves weights only, IModel.py:79-81): the fused
        optimizer's state in UNSHARDED form (collective: call on every rank) —
        ``{"tables": {"embeddings.{f}.weight": {"sum": [category_num, D]}, ...}, "dense": <state_dict of the dense
        companion>, "step": n}`` — so that a run can restart on a different number of GPUs."""
        opt = self.compiled_optimizers
        if opt is None or not hasattr(opt, "table_state"):
            raise RuntimeError("full_optimizer_state_dict needs a compiled pytorchrec_b200.optim sparse optimizer")
        out = {"tables": {}, "step": int(opt._step_count_fused)}
        for k, name in enumerate(self._GROUP_NAMES):
            for f, (col, table) in enumerate(zip(self.sparse_columns, self.sharded.groups[k])):
                shard_n = shard_rows(col.category_num, self.sharded.world, self.sharded.rank)
                out["tables"][f"{name}.{f}.weight"] = {
                    sn: self._gather_rows(st[:shard_n], col.category_num) for sn, st in opt.table_state(table.weight).items()}
        opt._ensure_dense()
        out["dense"] = copy.deepcopy(opt._dense.state_dict()) if opt._dense is not None else None
        return out

    @torch.no_grad()
    def load_full_optimizer_state_dict(self, state: Dict) -> None:
        """Inverse of ``full_optimizer_state_dict`` (each rank keeps rows ``rank::G`` of every table's state)."""
        opt = self.compiled_optimizers
        G, rank = self.sharded.world, self.sharded.rank
        for k, name in enumerate(self._GROUP_NAMES):
            for f, table in enumerate(self.sharded.groups[k]):
                full = state["tables"][f"{name}.{f}.weight"]
                cur = opt.table_state(table.weight)
                for sn, v in full.items():
                    shard = v[rank::G].to(cur[sn].device)
                    cur[sn][:shard.shape[0]].copy_(shard)
        if state.get("dense") is not None:
            opt._ensure_dense()
            opt._dense.load_state_dict(state["dense"])
        opt._step_count_fused = int(state.get("step", 0))
        opt._ptr_cache.clear()
        if self.sharded.exchange != "a2a":
            self.sharded.sync_peers()

    @torch.no_grad()
    def load_full_state_dict(self, state_dict: Dict[str, Tensor]) -> None:
        """Inverse of ``full_state_dict``: take an unsharded ``DeepFM`` checkpoint and keep rows ``rank::G``."""
        G, rank = self.sharded.world, self.sharded.rank
        own = self.state_dict()
        for k, name in enumerate(self._GROUP_NAMES):
            for f, table in enumerate(self.sharded.groups[k]):
                shard = state_dict[f"{name}.{f}.weight"][rank::G]
                table.weight[:shard.shape[0]].copy_(shard)
        for key, v in own.items():
            if not key.startswith("sharded."):
                v.copy_(state_dict[key])
        if self.sharded.exchange != "a2a":
            self.sharded.sync_peers()  # peers read these rows directly / the next exchange must see them

    def _dense_params(self) -> List[Tensor]:
        table_ids = {id(t.weight) for t in self.sharded.tables}
        return [p for p in self.parameters() if id(p) not in table_ids]

    def _before_optimizer_step(self) -> None:
        opt = self.compiled_optimizers
        red = self.sharded.dense_reducer() if hasattr(opt, "_peer_reduce") else None
        if red is not None:
            opt._peer_reduce = red  # the optimizer's K7 launch sums the ranks' gradients itself (one barrier inside)
        else:
            allreduce_dense_grads(self._dense_params(), self.sharded.group)
