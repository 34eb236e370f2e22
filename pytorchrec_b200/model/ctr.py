"""CTR models on the fused embedding path: FM and DeepFM.

The reference ships neither; they are written in its idiom — column objects in the
constructor, layers created in ``_init_weights`` (torchrec/model/FunkSVD.py:27-41), embedding + bias
+ global-bias layout of SVDPP.py:36-42, ``concat -> MLP -> Linear(..., 1, bias=False)`` head of
NCF.py:44-51,68-74, point-wise target ``label.float()`` (FunkSVD.py:53-55).  Outputs are logits
(train with the ``bce`` loss).  ``oracle/ref_models.py`` holds the plain-torch twins with identical
parameter names and RNG draw order.

  FM:      y = w0 + sum_f w_f[id_f] + <w_dense, x> + 0.5 * sum_k((sum_f v_fk)^2 - sum_f v_fk^2)
  DeepFM:  y = FM + Linear(MLP(concat_f v_f || x))
"""
from typing import Dict, List, Optional

import torch
from torch import Tensor
from torch.nn import Linear, Parameter

from ..feature_column import CategoricalColumnWithIdentity, NumericColumn
from .IModel import IModel
from .layer import MLP, AttentionPooling, CrossNet, EmbeddingTable, FMSecondOrder, MultiTableEmbedding


def _dense_matrix(dense_columns: List[NumericColumn], data: Dict[str, Tensor]) -> Optional[Tensor]:
    if not dense_columns:
        return None
    return torch.stack([c.get_feature_data(data) for c in dense_columns], dim=1)


class FM(IModel):
    @classmethod
    def get_argument_descriptions(cls) -> list:
        from ..utils.argument import ArgumentDescription
        descriptions = super().get_argument_descriptions()
        descriptions.extend([
            ArgumentDescription(name="emb_size", type_=int, help_info="embedding dimension (1, 2 or a multiple of 4 up to 128)", default_value=16, lower_closed_bound=1, upper_closed_bound=128)
        ])
        return descriptions

    def __init__(self, sparse_columns: List[CategoricalColumnWithIdentity], dense_columns: List[NumericColumn],
                 label_column: CategoricalColumnWithIdentity, emb_size: int, table_device=None,
                 table_dtype: torch.dtype = torch.float32, **kwargs):
        # table_device: build the tables directly on that device (skips the CPU init of multi-GB
        # tables; seeded CPU-init parity with the reference then no longer applies)
        # table_dtype: torch.bfloat16 stores the embedding rows in bf16 (EmbeddingTable; fp32 arithmetic and state)
        self.table_device = table_device
        self.table_dtype = table_dtype
        self.sparse_columns = list(sparse_columns)
        self.dense_columns = list(dense_columns or [])
        self.label_column = label_column
        self.emb_size = emb_size
        super().__init__(**kwargs)

    def _make_embedding(self, emb_size: int):
        """Table-group factory (overridden by the row-wise sharded variant)."""
        return MultiTableEmbedding(self.sparse_columns, emb_size, device=self.table_device, dtype=self.table_dtype)

    def _init_weights(self):
        self.embeddings = self._make_embedding(self.emb_size)
        self.first_order = self._make_embedding(1)
        if self.dense_columns:
            self.dense_linear = Linear(len(self.dense_columns), 1, bias=False)
        self.global_bias = Parameter(torch.tensor(0.0))
        self.fm2 = FMSecondOrder()

    def _target(self, data):
        target = self.label_column.get_feature_data(data)
        return target.float() if target is not None else None

    def _head(self, data, v: Tensor, w: Tensor, x: Optional[Tensor], want_deep_in: bool):
        """(FM logit, tower input or None): one fused pass (K8) where the shape allows, K3 + library ops otherwise."""
        from .layer.interaction import fm_head
        mlp = getattr(self, "mlp", None)
        units = mlp.mlp[0].linear.out_features if (want_deep_in and mlp is not None) else 0
        fused = fm_head(v, w, x, self.dense_linear.weight if x is not None else None, self.global_bias, want_deep_in,
                        tower_units=units, tower=mlp if want_deep_in else None)  # deep_in goes to self.mlp only
        if fused is not None:
            return fused
        logit = w.sum(dim=(1, 2)) + self.fm2(v) + self.global_bias
        if x is not None:
            logit = logit + self.dense_linear(x).squeeze(-1)
        deep_in = None
        if want_deep_in:
            flat = v.reshape(v.shape[0], -1)
            deep_in = torch.cat([flat, x], dim=1) if x is not None else flat
        return logit, deep_in

    def _lookups(self, data):
        """(v [B, F, D], w [B, F, 1]): the first-order tables (4-byte rows, latency-bound) are looked up — and, in
        the backward, updated — on a second stream beside the embedding tables."""
        from .layer.embedding import aux_stream
        v = self.embeddings(data)
        with aux_stream(v.device if v.is_cuda else None) as aux:
            w = self.first_order(data)
        aux.join(w)
        return v, w

    def forward(self, data: Dict[str, Tensor]):
        v, w = self._lookups(data)  # [B, F, D], [B, F, 1]
        x = _dense_matrix(self.dense_columns, data)
        logit, _ = self._head(data, v, w, x, False)
        return logit, self._target(data)


class DeepFM(FM):
    @classmethod
    def get_argument_descriptions(cls) -> list:
        from ..utils.argument import ArgumentDescription
        descriptions = super().get_argument_descriptions()
        descriptions.extend([
            ArgumentDescription(name="dropout", type_=float, help_info="dropout of the dense tower", default_value=0.0, lower_closed_bound=0.0, upper_open_bound=1.0)
        ])
        return descriptions

    def __init__(self, sparse_columns, dense_columns, label_column, emb_size: int, layers: List[int],
                 dropout: float = 0.0, table_device=None, table_dtype: torch.dtype = torch.float32, **kwargs):
        self.layers = list(layers)
        self.dropout = dropout
        super().__init__(sparse_columns, dense_columns, label_column, emb_size, table_device=table_device,
                         table_dtype=table_dtype, **kwargs)

    def _init_weights(self):
        super()._init_weights()
        in_units = len(self.sparse_columns) * self.emb_size + len(self.dense_columns)
        self.mlp = MLP(input_units=in_units, hidden_units_list=self.layers, activation="relu", dropout=self.dropout)
        self.deep_out = Linear(self.layers[-1], 1, bias=False)

    def _deep_logit(self, deep_in: Tensor) -> Tensor:
        from .layer.interaction import row_dot
        h = self.mlp(deep_in)
        y = row_dot(h, self.deep_out.weight, tower_handoff=True)  # h has no other consumer
        return y if y is not None else self.deep_out(h).squeeze(-1)

    def forward(self, data: Dict[str, Tensor]):
        v, w = self._lookups(data)
        x = _dense_matrix(self.dense_columns, data)
        logit, deep_in = self._head(data, v, w, x, True)
        return logit + self._deep_logit(deep_in), self._target(data)


class DCN(IModel):
    """DCN-v2 (parallel structure): ``y = Linear(concat(CrossNet(x0), MLP(x0)))``, ``x0 = concat_f v_f || x_dense``.
    Cross layers run in bf16 on tensor cores (tolerance 1e-2 vs the fp32 oracle twin); the deep tower and the
    embedding path stay fp32."""

    @classmethod
    def get_argument_descriptions(cls) -> list:
        from ..utils.argument import ArgumentDescription
        descriptions = super().get_argument_descriptions()
        descriptions.extend([
            ArgumentDescription(name="emb_size", type_=int, help_info="embedding dimension (1, 2 or a multiple of 4 up to 128)", default_value=16, lower_closed_bound=1, upper_closed_bound=128),
            ArgumentDescription(name="cross_layers", type_=int, help_info="number of DCN-v2 cross layers", default_value=3, lower_closed_bound=1),
            ArgumentDescription(name="dropout", type_=float, help_info="dropout of the dense tower", default_value=0.0, lower_closed_bound=0.0, upper_open_bound=1.0)
        ])
        return descriptions

    def __init__(self, sparse_columns: List[CategoricalColumnWithIdentity], dense_columns: List[NumericColumn],
                 label_column: CategoricalColumnWithIdentity, emb_size: int, cross_layers: int, layers: List[int],
                 dropout: float = 0.0, table_device=None, **kwargs):
        self.table_device = table_device
        self.sparse_columns = list(sparse_columns)
        self.dense_columns = list(dense_columns or [])
        self.label_column = label_column
        self.emb_size = emb_size
        self.cross_layers = cross_layers
        self.layers = list(layers)
        self.dropout = dropout
        super().__init__(**kwargs)

    def _init_weights(self):
        self.embeddings = MultiTableEmbedding(self.sparse_columns, self.emb_size, device=self.table_device)
        d = len(self.sparse_columns) * self.emb_size + len(self.dense_columns)
        self.cross = CrossNet(d, self.cross_layers)
        self.mlp = MLP(input_units=d, hidden_units_list=self.layers, activation="relu", dropout=self.dropout)
        self.out = Linear(d + self.layers[-1], 1, bias=False)

    def forward(self, data: Dict[str, Tensor]):
        v = self.embeddings(data)
        x = _dense_matrix(self.dense_columns, data)
        flat = v.reshape(v.shape[0], -1)
        x0 = torch.cat([flat, x], dim=1) if x is not None else flat
        # Linear(concat(cross, deep)) = Linear_c(cross) + Linear_d(deep): the concatenation is never materialised, and
        # the deep half closes the tower with K8's row-dot, whose backward hands the tower its gradient as planes
        from .layer.interaction import row_dot
        d = x0.shape[1]
        w = self.out.weight                                  # [1, d + H]
        c, h = self.cross.forward_head(x0, w[0, :d]), self.mlp(x0)   # c [B]: the cross half of the closing Linear
        deep = row_dot(h, w[0, d:], tower_handoff=True)      # h has no other consumer
        if deep is None:
            deep = torch.nn.functional.linear(h, w[:, d:]).squeeze(-1)
        logit = c + deep
        target = self.label_column.get_feature_data(data)
        return logit, (target.float() if target is not None else None)


class DIN(IModel):
    """Deep Interest Network on the fused path.  Inputs follow the reference's history wire format
    (HistoryDataReader.py:55-69): a right-padded id matrix ``[B, L]`` (0 = PAD) plus a length column clipped to
    >= 1.  Item and category tables are shared between the candidate and the history; candidate and history ids
    go through ONE fused lookup per step (so each table sees one sort / dedup / update), the query is
    ``[item(cand) || cate(cand)]`` and the keys ``[item(h_l) || cate(h_l)]``.
        y = Linear(MLP([user || q || sum_{l<len} a_l k_l]))"""

    @classmethod
    def get_argument_descriptions(cls) -> list:
        from ..utils.argument import ArgumentDescription
        descriptions = super().get_argument_descriptions()
        descriptions.extend([
            ArgumentDescription(name="emb_size", type_=int, help_info="embedding dimension (1, 2 or a multiple of 4 up to 128)", default_value=16, lower_closed_bound=1, upper_closed_bound=128),
            ArgumentDescription(name="dropout", type_=float, help_info="dropout of the dense tower", default_value=0.0, lower_closed_bound=0.0, upper_open_bound=1.0)
        ])
        return descriptions

    def __init__(self, uid_column, iid_column, cid_column, his_iid_column, his_cid_column, his_len_column, label_column,
                 emb_size: int, layers: List[int], attention_hidden=(80, 40), dropout: float = 0.0, table_device=None,
                 **kwargs):
        self.uid_column, self.iid_column, self.cid_column = uid_column, iid_column, cid_column
        self.his_iid_column, self.his_cid_column, self.his_len_column = his_iid_column, his_cid_column, his_len_column
        self.label_column = label_column
        self.emb_size, self.layers, self.attention_hidden, self.dropout = emb_size, list(layers), tuple(attention_hidden), dropout
        self.table_device = table_device
        super().__init__(**kwargs)

    def _init_weights(self):
        D = self.emb_size
        self.seq_emb = MultiTableEmbedding([self.iid_column, self.cid_column], D, device=self.table_device)
        self.user_emb = EmbeddingTable(self.uid_column.category_num, D, device=self.table_device)
        self.attention = AttentionPooling(2 * D, self.attention_hidden)
        self.mlp = MLP(input_units=5 * D, hidden_units_list=self.layers, activation="relu", dropout=self.dropout)
        self.out = Linear(self.layers[-1], 1, bias=False)

    def forward(self, data: Dict[str, Tensor]):
        D = self.emb_size
        cand_i, cand_c = self.iid_column.get_feature_data(data), self.cid_column.get_feature_data(data)
        his_i, his_c = self.his_iid_column.get_feature_data(data), self.his_cid_column.get_feature_data(data)
        lens = self.his_len_column.get_feature_data(data)
        B, L = his_i.shape
        # one lookup for candidate + history ([B, 1+L, item||cate]: q = row 0, keys = rows 1..); on the tensor-core
        # builds K4 reads the history rows from the tables by id and only the candidate rows are gathered
        q, pooled = self.seq_emb.lookup_attention(
            {self.iid_column.feature_name: cand_i, self.cid_column.feature_name: cand_c},
            {self.iid_column.feature_name: his_i, self.cid_column.feature_name: his_c}, lens, self.attention)
        user = self.user_emb(self.uid_column.get_feature_data(data))
        logit = self.out(self.mlp(torch.cat([user, q, pooled], dim=1))).squeeze(-1)
        target = self.label_column.get_feature_data(data)
        return logit, (target.float() if target is not None else None)
