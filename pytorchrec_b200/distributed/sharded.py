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
    """One exchange for every embedding width of the same fields: ids travel once, the rows of all widths share one
    all-to-all (interleaved per slot), the gradients likewise, and the owner sorts the received lookups once."""

    @staticmethod
    def forward(ctx, mod: "RowWiseShardedEmbedding", ids: Tensor, *weights):
        F, B = ids.shape
        G, dev = mod.world, ids.device
        C = list_capacity(B, G, mod.capacity_factor)
        send_ids, ret_pos = ops.a2a_pack_by_owner(ids, F, B, G, C, mod.overflow_flag(dev))
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


class RowWiseShardedEmbedding(nn.Module):
    """``MultiTableEmbedding`` whose tables are sharded row-wise over the process group.  ``emb_size`` may be a list
    (e.g. ``[16, 1]`` for DeepFM's embedding and first-order tables): every width gets its own table per column
    (``groups[k][f]`` = LOCAL shard ``EmbeddingTable(shard_rows(category_num), dim_k)``) and all widths share one
    id exchange, one row exchange and one gradient exchange per step.  One-hot fields only (``[B]`` ids per
    column) — the Criteo-shaped configs; pooled bags stay on ``MultiTableEmbedding``."""

    def __init__(self, columns: Sequence[CategoricalColumn], emb_size, group=None,
                 capacity_factor: float = 1.25, device=None):
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

    def overflow_flag(self, device) -> Tensor:
        t = self._overflow.get(device)
        if t is None:
            t = torch.zeros(1, dtype=torch.int32, device=device)
            self._overflow[device] = t
        return t

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
        if not self.egroups:
            self.egroups = [EmbeddingGroup([t for t in g], d) for g, d in zip(self.groups, self.dims)]
        ids = torch.stack([c.get_feature_data(batch).reshape(-1) for c in self.columns])  # [F, B]
        weights = []
        for g in self.groups:
            for t in g:
                t._tag()
                weights.append(t.weight)
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
                raise RuntimeError(f"an all-to-all lookup list overflowed its capacity (needed {v}); raise capacity_factor")
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
        logit = w.sum(dim=(1, 2)) + self.fm2(v) + self.global_bias
        if x is not None:
            logit = logit + self.dense_linear(x).squeeze(-1)
        flat = v.reshape(v.shape[0], -1)
        deep_in = torch.cat([flat, x], dim=1) if x is not None else flat
        logit = logit + self.deep_out(self.mlp(deep_in)).squeeze(-1)
        return logit, self._target(data)

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
                local = table.weight.detach().contiguous()
                cap = (col.category_num + G - 1) // G
                padded = torch.zeros(cap, local.shape[1], dtype=local.dtype, device=local.device)
                padded[:local.shape[0]] = local
                parts = [torch.empty_like(padded) for _ in range(G)]
                dist.all_gather(parts, padded, group=self.sharded.group)
                full = torch.stack(parts, dim=1).reshape(cap * G, local.shape[1])[:col.category_num]
                out[f"{name}.{f}.weight"] = full.cpu()
        return out

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

    def _dense_params(self) -> List[Tensor]:
        table_ids = {id(t.weight) for t in self.sharded.tables}
        return [p for p in self.parameters() if id(p) not in table_ids]

    def _before_optimizer_step(self) -> None:
        allreduce_dense_grads(self._dense_params())
