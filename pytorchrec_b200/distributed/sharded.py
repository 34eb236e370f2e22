"""Row-wise sharded embedding tables across the GPUs of one node (C1) + dense-tower gradient
allreduce (C2).  One process per GPU (torchrun); NCCL over NVLink / NVSwitch through
``torch.distributed``.  The reference is single-device (torchrec/task/Task.py:187-190): nothing here
has a reference counterpart, it widens the same hot path to the 8 x B200 box.

Plan:  owner(id) = id mod G,  local_row = id div G.
  forward   pack ids by owner (fixed-capacity lists, no host sync)  -> all_to_all(ids)
            -> owner-side fused gather straight into the return layout -> all_to_all(rows)
            -> local gather by slot  -> [B, F, D]
  backward  scatter gradient rows into the send layout -> all_to_all(grads)
            -> owner-side sort / dedup / segment-sum / fused optimizer update (no gradient returns)
  dense     one flat all_reduce(SUM) / G of the dense-tower gradients.
"""
import math
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


def list_capacity(batch: int, world: int, factor: float = 1.25) -> int:
    """Slots per (owner, field) list: expected batch/world plus slack, multiple of 16."""
    c = int(math.ceil(batch / world * factor)) + 64
    return min((c + 15) // 16 * 16, (batch + 15) // 16 * 16)


def allreduce_dense_grads(params: Sequence[Tensor], group=None) -> None:
    """Average ``.grad`` of the replicated dense parameters over the ranks with ONE flat all_reduce."""
    grads = [p.grad for p in params if p.grad is not None and not p.grad.is_sparse]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


class _ShardedLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: "RowWiseShardedEmbedding", ids: Tensor, *weights):
        F, B = ids.shape
        G, D, dev = mod.world, mod.emb_size, ids.device
        C = list_capacity(B, G, mod.capacity_factor)
        send_ids, ret_pos = ops.a2a_pack_by_owner(ids, F, B, G, C, mod.overflow_flag(dev))
        recv_ids = torch.empty_like(send_ids)
        dist.all_to_all_single(recv_ids, send_ids, group=mod.group)            # [G_src, F, C]
        own_ids = recv_ids.permute(1, 0, 2).contiguous().view(-1)              # [F, G_src, C]
        layout = mod.owner_layout(C)
        tables = mod.egroup.table_set.refresh([w.detach() for w in weights])
        bufs = mod.buffers(C, dev)  # persistent: fixed addresses (pointer arrays built once, graph-capturable)
        rows_out, recv_rows = bufs["rows_out"], bufs["recv_rows"]             # [G_src, F, C, D] / [G_owner, F, C, D]
        ops.gather_pool_fwd(tables, layout, own_ids, None, C, out=rows_out, err_flag=mod.egroup.err_flag(dev),
                            out_row_stride=D)
        dist.all_to_all_single(recv_rows, rows_out, group=mod.group)
        # local gather by slot: out[b, f] = recv_rows[ret_pos[f, b]]
        slot_tables = bufs["slot_tables"]
        out, _ = ops.gather_pool_fwd(slot_tables, mod.slot_layout(), ret_pos.view(-1).long(), None, B)
        ctx.mod, ctx.C, ctx.shape = mod, C, (F, B)
        ctx.save_for_backward(own_ids, ret_pos)
        return out.view(B, F, D)

    @staticmethod
    def backward(ctx, grad_out):
        mod, C = ctx.mod, ctx.C
        F, B = ctx.shape
        G, D = mod.world, mod.emb_size
        own_ids, ret_pos = ctx.saved_tensors
        grad_out = grad_out.reshape(B, F * D)
        if not grad_out.is_contiguous():
            grad_out = grad_out.contiguous()
        # slots that carry no lookup are -1 on the owner (masked in the sort): no need to clear them
        bufs = mod.buffers(C, grad_out.device)
        send_g, recv_g = bufs["send_g"], bufs["recv_g"]
        ops.a2a_scatter_rows(grad_out, ret_pos, B, F, D, mod.grad_scale, send_g)
        dist.all_to_all_single(recv_g, send_g, group=mod.group)                # [G_src, F, C, D]
        layout = mod.owner_layout(C)
        bind = mod.egroup.binding()
        if bind is None:
            raise RuntimeError("row-wise sharded tables need a pytorchrec_b200.optim sparse optimizer")
        optimizer, group = bind
        s1, s2, args = optimizer._fused_prepare(mod.egroup, group)  # may interleave weight | state: before pointers
        tables = mod.egroup.table_set.refresh([t.weight.data for t in mod.egroup.tables])
        srt = ops.sort_dedup(tables, layout, own_ids, None, C)
        ops.bwd_fused(tables, s1, s2, layout, C, srt, recv_g, None, args, grad_row_stride=D)
        return (None, None) + (None,) * len(mod.egroup.tables)


class RowWiseShardedEmbedding(nn.ModuleList):
    """``MultiTableEmbedding`` whose tables are sharded row-wise over the process group.  Children are the
    LOCAL shards (``EmbeddingTable(shard_rows(category_num), emb_size)``).  One-hot fields only
    (``[B]`` ids per column) — the Criteo-shaped configs; pooled bags stay on ``MultiTableEmbedding``."""

    def __init__(self, columns: Sequence[CategoricalColumn], emb_size: int, group=None,
                 capacity_factor: float = 1.25, device=None):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("RowWiseShardedEmbedding needs an initialised torch.distributed process group")
        self.columns = list(columns)
        self.emb_size = int(emb_size)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.capacity_factor = capacity_factor
        self.grad_scale = 1.0 / self.world  # mean loss over the GLOBAL batch
        if self.world * len(self.columns) > 256:
            raise ValueError("world_size * n_fields must be <= 256")
        for c in self.columns:
            self.append(EmbeddingTable(max(shard_rows(c.category_num, self.world, self.rank), 1), self.emb_size,
                                       device=device))
        self.egroup: Optional[EmbeddingGroup] = None
        self.slot_tables = ops.TableSet()
        self._owner_layouts: Dict[int, ops.FeatureLayout] = {}
        self._slot_layout: Optional[ops.FeatureLayout] = None
        self._overflow: Dict[torch.device, Tensor] = {}
        self._bufs: Dict[tuple, dict] = {}

    @property
    def weight(self) -> Tensor:  # see MultiTableEmbedding.weight
        return torch.empty(0)

    def buffers(self, C: int, device) -> dict:
        key = (C, device)
        b = self._bufs.get(key)
        if b is None:
            n = self.world * len(self.columns) * C
            mk = lambda: torch.empty(n, self.emb_size, dtype=torch.float32, device=device)  # noqa: E731
            b = {"rows_out": mk(), "recv_rows": mk(), "send_g": mk(), "recv_g": mk()}
            b["slot_tables"] = ops.TableSet().refresh([b["recv_rows"]])
            self._bufs[key] = b
        return b

    def overflow_flag(self, device) -> Tensor:
        t = self._overflow.get(device)
        if t is None:
            t = torch.zeros(1, dtype=torch.int32, device=device)
            self._overflow[device] = t
        return t

    def owner_layout(self, C: int) -> ops.FeatureLayout:
        """Owner-side view of the received lists: feature (f, src) reads table f, batch = C, and writes
        its rows at ((src*F + f)*C + c)*D — i.e. directly in the all-to-all return layout."""
        lay = self._owner_layouts.get(C)
        if lay is None:
            F, G, D = len(self.columns), self.world, self.emb_size
            specs = [dict(table=f, bag_len=1, neg_is_pad=True, out_col=(src * F + f) * C * D)
                     for f in range(F) for src in range(G)]
            lay = ops.FeatureLayout(specs, D, F)
            self._owner_layouts[C] = lay
        return lay

    def slot_layout(self) -> ops.FeatureLayout:
        if self._slot_layout is None:
            F, D = len(self.columns), self.emb_size
            self._slot_layout = ops.FeatureLayout([dict(table=0, bag_len=1, neg_is_pad=True) for _ in range(F)], D, 1)
        return self._slot_layout

    def forward(self, batch: Dict[str, Tensor]) -> Tensor:
        if self.egroup is None:
            self.egroup = EmbeddingGroup([m for m in self], self.emb_size)
        ids = torch.stack([c.get_feature_data(batch).reshape(-1) for c in self.columns])  # [F, B]
        for t in self:
            t._tag()
        return _ShardedLookup.apply(self, ids, *[t.weight for t in self])

    @property
    def tables(self) -> List[EmbeddingTable]:
        return [m for m in self]

    def check_index_errors(self) -> None:
        self.check_errors()

    def check_errors(self) -> None:
        """Synchronising check of the overflow / out-of-range flags."""
        for t in self._overflow.values():
            v = int(t.item())
            if v:
                t.zero_()
                raise RuntimeError(f"an all-to-all lookup list overflowed its capacity (needed {v}); raise capacity_factor")
        if self.egroup is not None:
            self.egroup.check_index_errors()


class ShardedDeepFM(DeepFM):
    """DeepFM with both table groups row-wise sharded and the dense tower data-parallel."""

    def _make_embedding(self, emb_size: int):
        return RowWiseShardedEmbedding(self.sparse_columns, emb_size, device=self.table_device)

    def _dense_params(self) -> List[Tensor]:
        table_ids = {id(t.weight) for m in (self.embeddings, self.first_order) for t in m}
        return [p for p in self.parameters() if id(p) not in table_ids]

    def _before_optimizer_step(self) -> None:
        allreduce_dense_grads(self._dense_params())
