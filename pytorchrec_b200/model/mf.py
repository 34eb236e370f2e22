"""The reference's own embedding models re-hosted on the fused tables: FunkSVD, SVD++ and NCF.

Constructor arguments, attribute names (hence ``state_dict`` keys), forward shapes and outputs follow
torchrec/model/FunkSVD.py:12-67, SVDPP.py:12-91 and NCF.py:13-79; only the embedding modules differ
(``EmbeddingTable`` / ``MultiTableEmbedding`` instead of ``nn.Embedding``) and, for NCF, the tower's Linear layers
run on K6 when they are large enough (``model/layer/dense.py``).  They exist so that the
CUDA path can be checked against the UNMODIFIED reference models on identical weights and batches.
"""
from typing import Dict

import torch
from torch import Tensor
from torch.nn import Parameter

from ..feature_column import CategoricalColumnWithIdentity
from .IModel import IModel
from .layer import MLP, EmbeddingTable


def _candidate_target(prediction: Tensor) -> Tensor:
    target = torch.zeros_like(prediction, dtype=torch.float32)
    target[:, 0] = 1
    return target


class FunkSVD(IModel):
    @classmethod
    def get_argument_descriptions(cls) -> list:
        from ..utils.argument import ArgumentDescription
        descriptions = super().get_argument_descriptions()
        descriptions.extend([
            ArgumentDescription(name="emb_size", type_=int, help_info="embedding dimension", default_value=64, lower_closed_bound=1)
        ])
        return descriptions

    def __init__(self, uid_column: CategoricalColumnWithIdentity, iid_column: CategoricalColumnWithIdentity,
                 label_column: CategoricalColumnWithIdentity, emb_size: int, **kwargs):
        self.uid_column = uid_column
        self.iid_column = iid_column
        self.label_column = label_column
        self.emb_size = emb_size
        super().__init__(**kwargs)

    def _init_weights(self):
        self.u_embeddings = EmbeddingTable(self.uid_column.category_num, self.emb_size)
        self.i_embeddings = EmbeddingTable(self.iid_column.category_num, self.emb_size)

    def forward(self, data: Dict[str, Tensor]):
        u_ids = self.uid_column.get_feature_data(data)  # [B]
        i_ids = self.iid_column.get_feature_data(data)  # [B] or [B, N]
        u = self.u_embeddings(u_ids)
        i = self.i_embeddings(i_ids)
        if i_ids.dim() == 1:
            prediction = (u * i).sum(dim=-1)
            target = self.label_column.get_feature_data(data)
            if target is not None:
                target = target.float()
        else:
            prediction = (u.unsqueeze(1) * i).sum(dim=-1)  # [B, N]
            target = _candidate_target(prediction)
        return prediction, target


class SVDPP(IModel):
    @classmethod
    def get_argument_descriptions(cls) -> list:
        from ..utils.argument import ArgumentDescription
        descriptions = super().get_argument_descriptions()
        descriptions.extend([
            ArgumentDescription(name="emb_size", type_=int, help_info="embedding dimension", default_value=64, lower_closed_bound=1)
        ])
        return descriptions

    def __init__(self, random_seed: int, uid_column: CategoricalColumnWithIdentity,
                 iid_column: CategoricalColumnWithIdentity, iids_column: CategoricalColumnWithIdentity,
                 label_column: CategoricalColumnWithIdentity, emb_size: int):
        self.uid_column = uid_column
        self.iid_column = iid_column
        self.iids_column = iids_column
        self.label_column = label_column
        self.emb_size = emb_size
        super().__init__(random_seed)

    def _init_weights(self):
        self.u_embeddings = EmbeddingTable(self.uid_column.category_num, self.emb_size)
        self.i_embeddings = EmbeddingTable(self.iid_column.category_num, self.emb_size)
        self.implicit_i_embeddings = EmbeddingTable(self.iid_column.category_num, self.emb_size)
        self.u_bias = EmbeddingTable(self.uid_column.category_num, 1)
        self.i_bias = EmbeddingTable(self.iid_column.category_num, 1)
        self.global_bias = Parameter(torch.tensor(0.0))

    def forward(self, data: Dict[str, Tensor]):
        u_ids = self.uid_column.get_feature_data(data)
        i_ids = self.iid_column.get_feature_data(data)
        # implicit-feedback bag: masked sum / sqrt(count), SVDPP.py:49-55 -> pooling 'sqrtn', mask 'pad'
        implicit = self.implicit_i_embeddings.pooled(self.iids_column.get_feature_data(data), "sqrtn", "pad")
        u = self.u_embeddings(u_ids) + implicit
        i = self.i_embeddings(i_ids)
        u_bias = self.u_bias(u_ids).squeeze(-1)
        i_bias = self.i_bias(i_ids).squeeze(-1)
        if i_ids.dim() == 1:
            prediction = (u * i).sum(dim=-1) + u_bias + i_bias + self.global_bias
            target = self.label_column.get_feature_data(data)
            if target is not None:
                target = target.float()
        else:
            prediction = (u.unsqueeze(1) * i).sum(dim=-1) + u_bias.unsqueeze(1) + i_bias + self.global_bias
            target = _candidate_target(prediction)
        return prediction, target


class NCF(IModel):
    """Neural collaborative filtering (torchrec/model/NCF.py:38-79): GMF branch ``mf_u * mf_i`` beside
    ``MLP([mlp_u, mlp_i])``, ``Linear(emb + layers[-1], 1, bias=False)`` on their concatenation.  ``iid`` is
    ``[B, N]`` (candidates; N = 2 for pair-wise training) and the user is repeated over the candidates."""

    @classmethod
    def get_argument_descriptions(cls) -> list:
        from ..utils.argument import ArgumentDescription
        descriptions = super().get_argument_descriptions()
        descriptions.extend([
            ArgumentDescription(name="emb_size", type_=int, help_info="embedding dimension", default_value=64, lower_closed_bound=1),
            ArgumentDescription(name="dropout", type_=float, help_info="dropout of the dense tower", default_value=0.0, lower_closed_bound=0.0, upper_open_bound=1.0)
        ])
        return descriptions

    def __init__(self, random_seed: int, uid_column: CategoricalColumnWithIdentity,
                 iid_column: CategoricalColumnWithIdentity, label_column: CategoricalColumnWithIdentity,
                 emb_size: int, layers, dropout: float):
        self.uid_column = uid_column
        self.iid_column = iid_column
        self.label_column = label_column
        self.emb_size = emb_size
        self.layers = list(layers)
        self.dropout = dropout
        super().__init__(random_seed)

    def _init_weights(self):
        n_u, n_i = self.uid_column.category_num, self.iid_column.category_num
        self.mf_u_embeddings = EmbeddingTable(n_u, self.emb_size)
        self.mf_i_embeddings = EmbeddingTable(n_i, self.emb_size)
        self.mlp_u_embeddings = EmbeddingTable(n_u, self.emb_size)
        self.mlp_i_embeddings = EmbeddingTable(n_i, self.emb_size)
        self.mlp = MLP(input_units=2 * self.emb_size, hidden_units_list=self.layers, activation="relu",
                       dropout=self.dropout)
        self.prediction = torch.nn.Linear(self.emb_size + self.layers[-1], 1, bias=False)

    def forward(self, data: Dict[str, Tensor]):
        u_ids = self.uid_column.get_feature_data(data)            # [B]
        i_ids = self.iid_column.get_feature_data(data)            # [B, N]
        n = i_ids.shape[1]
        u_ids = u_ids.unsqueeze(-1).repeat(1, n).reshape(-1)      # [B * N]
        i_ids = i_ids.reshape(-1)
        mf = self.mf_u_embeddings(u_ids) * self.mf_i_embeddings(i_ids)
        deep = self.mlp(torch.cat([self.mlp_u_embeddings(u_ids), self.mlp_i_embeddings(i_ids)], dim=-1))
        prediction = self.prediction(torch.cat([mf, deep], dim=-1)).reshape(-1, n)
        return prediction, _candidate_target(prediction)
