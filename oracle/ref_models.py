"""Model-level CPU restatements in the reference's idiom (test oracle; see oracle/__init__.py).

Every model here is built from stock ``torch.nn.Embedding`` / ``Linear`` modules, takes dense
autograd gradients (``aten::embedding_dense_backward``) and is stepped by a dense ``torch.optim``
optimizer — i.e. exactly what the reference does on this path.  Column objects are duck-typed
(``get_feature_data(batch)``, ``category_num``) so this file imports nothing from the product.
"""
from typing import Dict, List, Optional

import torch
from torch import Tensor, nn
from torch.nn import Embedding, Linear, Parameter


def set_seed_ref(seed: int) -> None:
    """torchrec/utils/global_utils.py:7-16 (CPU part; the cudnn flags do not affect CPU draws)."""
    torch.manual_seed(seed)


class IModelRef(nn.Module):
    """Lifecycle + step of torchrec/model/IModel.py restated:
    seed -> ``_init_weights`` -> re-draw N(0, 0.01) over Linear/Embedding weights and biases in
    ``Module.apply`` order (:37-71); two param groups, biases undecayed (:83-92); ``train_step`` =
    forward, loss, zero_grad, backward, step (:116-125)."""

    def __init__(self, random_seed: int):
        set_seed_ref(random_seed)
        super().__init__()
        self._init_weights()
        self.apply(self._reset)

    def _init_weights(self):
        raise NotImplementedError

    @staticmethod
    def _reset(m):
        tname = str(type(m))
        if 'Linear' in tname:
            nn.init.normal_(m.weight, mean=0.0, std=0.01)
            if m.bias is not None:
                nn.init.normal_(m.bias, mean=0.0, std=0.01)
        elif 'Embedding' in tname:
            nn.init.normal_(m.weight, mean=0.0, std=0.01)

    def get_parameters(self):
        weights = [p for n, p in self.named_parameters() if p.requires_grad and 'bias' not in n]
        biases = [p for n, p in self.named_parameters() if p.requires_grad and 'bias' in n]
        return [{'params': weights}, {'params': biases, 'weight_decay': 0.0}]

    def compile(self, optimizer, loss):
        self.opt, self.loss_fn = optimizer, loss

    def fp64(self) -> "IModelRef":
        """The same model in double precision — the "exact" result both fp32 implementations (this oracle's and the
        CUDA path's) are measured against where fp32 summation order decides the outcome (Adagrad's
        g / (sqrt(sum g^2) + eps) is discontinuous where duplicate gradients cancel).  Call before ``compile``."""
        return self.double()

    def train_step(self, data: Dict[str, Tensor]):
        self.train()
        prediction, target = self(data)
        loss = self.loss_fn(prediction, target.to(prediction.dtype))
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return {"loss": loss}


def _one_hot_target(prediction: Tensor) -> Tensor:
    t = torch.zeros_like(prediction, dtype=torch.float32)
    t[:, 0] = 1
    return t


class FunkSVDRef(IModelRef):
    """torchrec/model/FunkSVD.py:27-67: two tables, dot product; ``[B, N]`` candidates broadcast the user."""

    def __init__(self, random_seed, uid_column, iid_column, label_column, emb_size):
        self.uid_column, self.iid_column, self.label_column, self.emb_size = uid_column, iid_column, label_column, emb_size
        super().__init__(random_seed)

    def _init_weights(self):
        self.u_embeddings = Embedding(self.uid_column.category_num, self.emb_size)
        self.i_embeddings = Embedding(self.iid_column.category_num, self.emb_size)

    def forward(self, data):
        u = self.u_embeddings(self.uid_column.get_feature_data(data))
        i_ids = self.iid_column.get_feature_data(data)
        i = self.i_embeddings(i_ids)
        if i_ids.dim() == 1:
            target = self.label_column.get_feature_data(data)
            return (u * i).sum(-1), (target.float() if target is not None else None)
        pred = (u.unsqueeze(1) * i).sum(-1)
        return pred, _one_hot_target(pred)


class SVDPPRef(IModelRef):
    """torchrec/model/SVDPP.py:36-91: user + sqrt-n pooled implicit items, item, two bias tables, global bias."""

    def __init__(self, random_seed, uid_column, iid_column, iids_column, label_column, emb_size):
        self.uid_column, self.iid_column, self.iids_column = uid_column, iid_column, iids_column
        self.label_column, self.emb_size = label_column, emb_size
        super().__init__(random_seed)

    def _init_weights(self):
        n_u, n_i = self.uid_column.category_num, self.iid_column.category_num
        self.u_embeddings = Embedding(n_u, self.emb_size)
        self.i_embeddings = Embedding(n_i, self.emb_size)
        self.implicit_i_embeddings = Embedding(n_i, self.emb_size)
        self.u_bias = Embedding(n_u, 1)
        self.i_bias = Embedding(n_i, 1)
        self.global_bias = Parameter(torch.tensor(0.0))

    def forward(self, data):
        u_ids = self.uid_column.get_feature_data(data)
        i_ids = self.iid_column.get_feature_data(data)
        his = self.iids_column.get_feature_data(data)
        m = his.gt(0).float()                                            # :49
        pooled = (self.implicit_i_embeddings(his) * m.unsqueeze(-1)).sum(1)  # :50-52
        pooled = pooled / m.sum(-1).sqrt().unsqueeze(-1)                 # :53-55
        u = self.u_embeddings(u_ids) + pooled
        i = self.i_embeddings(i_ids)
        bu = self.u_bias(u_ids).squeeze(-1)
        bi = self.i_bias(i_ids).squeeze(-1)
        if i_ids.dim() == 1:
            target = self.label_column.get_feature_data(data)
            return (u * i).sum(-1) + bu + bi + self.global_bias, (target.float() if target is not None else None)
        pred = (u.unsqueeze(1) * i).sum(-1) + bu.unsqueeze(1) + bi + self.global_bias
        return pred, _one_hot_target(pred)


# ------------------------------------------------------------------------------------------------
# FM / DeepFM (NOT in the reference: parity unpinned; restated in its idiom)
# ------------------------------------------------------------------------------------------------
class DenseRef(nn.Module):
    """torchrec/model/layer/Dense.py:4-24: Linear -> ReLU -> Dropout (activation always ReLU)."""

    def __init__(self, input_units: int, output_units: int, dropout: float):
        super().__init__()
        self.linear = Linear(input_units, output_units)
        self.activation = nn.ReLU()
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        return self.dropout(self.activation(self.linear(x)))


class MLPRef(nn.Module):
    """torchrec/model/layer/MLP.py:8-23: ``Sequential`` of ``dense_{i}`` blocks under ``mlp``."""

    def __init__(self, input_units: int, hidden: List[int], dropout: float):
        super().__init__()
        self.mlp = nn.Sequential()
        units = input_units
        for k, h in enumerate(hidden):
            self.mlp.add_module(f"dense_{k}", DenseRef(units, h, dropout))
            units = h

    def forward(self, x):
        return self.mlp(x)


class FMRef(IModelRef):
    """y = w0 + sum_f w_f[id_f] + <w_dense, x> + 0.5 sum_k((sum_f v_fk)^2 - sum_f v_fk^2).
    One ``Embedding`` per column for v and for w (SVDPP.py:36-42 pattern), ``Linear(n_dense, 1, bias=False)``."""

    def __init__(self, random_seed, sparse_columns, dense_columns, label_column, emb_size):
        self.sparse_columns, self.dense_columns = list(sparse_columns), list(dense_columns or [])
        self.label_column, self.emb_size = label_column, emb_size
        super().__init__(random_seed)

    def _init_weights(self):
        self.embeddings = nn.ModuleList([Embedding(c.category_num, self.emb_size) for c in self.sparse_columns])
        self.first_order = nn.ModuleList([Embedding(c.category_num, 1) for c in self.sparse_columns])
        if self.dense_columns:
            self.dense_linear = Linear(len(self.dense_columns), 1, bias=False)
        self.global_bias = Parameter(torch.tensor(0.0))

    def _parts(self, data):
        ids = [c.get_feature_data(data) for c in self.sparse_columns]
        v = torch.stack([e(i) for e, i in zip(self.embeddings, ids)], dim=1)        # [B, F, D]
        w = torch.stack([e(i) for e, i in zip(self.first_order, ids)], dim=1)       # [B, F, 1]
        x = torch.stack([c.get_feature_data(data) for c in self.dense_columns], dim=1) if self.dense_columns else None
        x = x.to(v.dtype) if x is not None else None  # a no-op in fp32; widens the inputs of an fp64() twin
        s = v.sum(dim=1)
        fm2 = 0.5 * (s * s - (v * v).sum(dim=1)).sum(dim=-1)
        logit = w.sum(dim=(1, 2)) + fm2 + self.global_bias
        if x is not None:
            logit = logit + self.dense_linear(x).squeeze(-1)
        return v, x, logit

    def _target(self, data):
        t = self.label_column.get_feature_data(data)
        return t.float() if t is not None else None

    def forward(self, data):
        _, _, logit = self._parts(data)
        return logit, self._target(data)


class DeepFMRef(FMRef):
    """FM + ``Linear(MLP(concat_f v_f || x), 1, bias=False)`` (head pattern of NCF.py:44-51,68-74)."""

    def __init__(self, random_seed, sparse_columns, dense_columns, label_column, emb_size, layers, dropout=0.0):
        self.layers, self.dropout = list(layers), dropout
        super().__init__(random_seed, sparse_columns, dense_columns, label_column, emb_size)

    def _init_weights(self):
        super()._init_weights()
        in_units = len(self.sparse_columns) * self.emb_size + len(self.dense_columns)
        self.mlp = MLPRef(in_units, self.layers, self.dropout)
        self.deep_out = Linear(self.layers[-1], 1, bias=False)

    def forward(self, data):
        v, x, logit = self._parts(data)
        flat = v.reshape(v.shape[0], -1)
        deep_in = torch.cat([flat, x], dim=1) if x is not None else flat
        return logit + self.deep_out(self.mlp(deep_in)).squeeze(-1), self._target(data)


class NCFRef(IModelRef):
    """torchrec/model/NCF.py:38-79 restated: GMF (element-wise product of one table pair) beside an MLP over the
    concatenation of a second table pair, ``Linear(emb + layers[-1], 1, bias=False)`` on their concatenation; the user
    is repeated over the ``[B, N]`` candidate items; target = one-hot on candidate 0.  Pinned by
    tests/golden/reference_ncf.npz (reference run) — which also pins ``MLPRef`` / ``DenseRef``."""

    def __init__(self, random_seed, uid_column, iid_column, label_column, emb_size, layers, dropout=0.0):
        self.uid_column, self.iid_column, self.label_column = uid_column, iid_column, label_column
        self.emb_size, self.layers, self.dropout = emb_size, list(layers), dropout
        super().__init__(random_seed)

    def _init_weights(self):
        n_u, n_i = self.uid_column.category_num, self.iid_column.category_num
        self.mf_u_embeddings = Embedding(n_u, self.emb_size)
        self.mf_i_embeddings = Embedding(n_i, self.emb_size)
        self.mlp_u_embeddings = Embedding(n_u, self.emb_size)
        self.mlp_i_embeddings = Embedding(n_i, self.emb_size)
        self.mlp = MLPRef(2 * self.emb_size, self.layers, self.dropout)
        self.prediction = Linear(self.emb_size + self.layers[-1], 1, bias=False)

    def forward(self, data):
        u_ids = self.uid_column.get_feature_data(data)
        i_ids = self.iid_column.get_feature_data(data)
        n = i_ids.shape[1]
        u_ids = u_ids.unsqueeze(-1).repeat(1, n).reshape(-1)
        i_ids = i_ids.reshape(-1)
        mf = self.mf_u_embeddings(u_ids) * self.mf_i_embeddings(i_ids)
        deep = self.mlp(torch.cat([self.mlp_u_embeddings(u_ids), self.mlp_i_embeddings(i_ids)], dim=-1))
        pred = self.prediction(torch.cat([mf, deep], dim=-1)).reshape(-1, n)
        return pred, _one_hot_target(pred)


class DCNRef(IModelRef):
    """DCN-v2, parallel structure (NOT in the reference: parity unpinned).  Cross layer
    ``x_{l+1} = x0 * (x_l W_l^T + b_l) + x_l`` with ``nn.Linear(d, d)`` per layer, fp32 throughout; head and
    tower in the reference's idiom (MLP.py:8-23, ``Linear(..., 1, bias=False)`` as NCF.py:51)."""

    def __init__(self, random_seed, sparse_columns, dense_columns, label_column, emb_size, cross_layers, layers, dropout=0.0):
        self.sparse_columns, self.dense_columns = list(sparse_columns), list(dense_columns or [])
        self.label_column, self.emb_size = label_column, emb_size
        self.cross_layers, self.layers, self.dropout = cross_layers, list(layers), dropout
        super().__init__(random_seed)

    def _init_weights(self):
        self.embeddings = nn.ModuleList([Embedding(c.category_num, self.emb_size) for c in self.sparse_columns])
        d = len(self.sparse_columns) * self.emb_size + len(self.dense_columns)
        self.cross = nn.Module()
        self.cross.layers = nn.ModuleList([Linear(d, d) for _ in range(self.cross_layers)])
        self.mlp = MLPRef(d, self.layers, self.dropout)
        self.out = Linear(d + self.layers[-1], 1, bias=False)

    def forward(self, data):
        ids = [c.get_feature_data(data) for c in self.sparse_columns]
        flat = torch.cat([e(i) for e, i in zip(self.embeddings, ids)], dim=1)
        x0 = flat
        if self.dense_columns:
            x0 = torch.cat([flat, torch.stack([c.get_feature_data(data) for c in self.dense_columns], dim=1).to(flat.dtype)],
                           dim=1)
        x = x0
        for lin in self.cross.layers:
            x = x0 * lin(x) + x
        logit = self.out(torch.cat([x, self.mlp(x0)], dim=1)).squeeze(-1)
        t = self.label_column.get_feature_data(data)
        return logit, (t.float() if t is not None else None)


def cross_net_ref(x0: Tensor, weights, biases) -> Tensor:
    x = x0
    for W, b in zip(weights, biases):
        x = x0 * (x @ W.t() + b) + x
    return x


def din_attention_ref(q: Tensor, keys: Tensor, lens: Optional[Tensor], fc1, fc2, fc3) -> Tensor:
    """DIN activation unit + masked weighted sum in plain torch: materialises [B, L, 4*DQ] and the hidden layers."""
    B, L, DQ = keys.shape
    qe = q.unsqueeze(1).expand(B, L, DQ)
    z = torch.cat([qe, keys, qe - keys, qe * keys], dim=-1)
    a = fc3(torch.relu(fc2(torch.relu(fc1(z))))).squeeze(-1)
    if lens is not None:
        a = a * (torch.arange(L).unsqueeze(0) < lens.reshape(B, 1)).to(a.dtype)
    return (a.unsqueeze(-1) * keys).sum(dim=1)


class _AttentionRef(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = Linear(4 * dim, hidden[0]), Linear(hidden[0], hidden[1]), Linear(hidden[1], 1)


class DINRef(IModelRef):
    """Deep Interest Network (NOT in the reference: parity unpinned).  History conventions are the reference's:
    right-padded ``[B, L]`` ids with 0 = PAD and a length column (HistoryDataReader.py:55-69); shared item /
    category tables as SASRec shares ``i_embeddings`` between candidates and history (SASRec.py:85-86)."""

    def __init__(self, random_seed, uid_column, iid_column, cid_column, his_iid_column, his_cid_column, his_len_column,
                 label_column, emb_size, layers, attention_hidden=(80, 40), dropout=0.0):
        self.cols = (uid_column, iid_column, cid_column, his_iid_column, his_cid_column, his_len_column, label_column)
        self.emb_size, self.layers, self.attention_hidden, self.dropout = emb_size, list(layers), attention_hidden, dropout
        super().__init__(random_seed)

    def _init_weights(self):
        uid, iid, cid = self.cols[0], self.cols[1], self.cols[2]
        D = self.emb_size
        self.seq_emb = nn.ModuleList([Embedding(iid.category_num, D), Embedding(cid.category_num, D)])
        self.user_emb = Embedding(uid.category_num, D)
        self.attention = _AttentionRef(2 * D, self.attention_hidden)
        self.mlp = MLPRef(5 * D, self.layers, self.dropout)
        self.out = Linear(self.layers[-1], 1, bias=False)

    def forward(self, data):
        uid, iid, cid, hi, hc, hl, lab = self.cols
        q = torch.cat([self.seq_emb[0](iid.get_feature_data(data)), self.seq_emb[1](cid.get_feature_data(data))], dim=-1)
        keys = torch.cat([self.seq_emb[0](hi.get_feature_data(data)), self.seq_emb[1](hc.get_feature_data(data))], dim=-1)
        pooled = din_attention_ref(q, keys, hl.get_feature_data(data), self.attention.fc1, self.attention.fc2,
                                   self.attention.fc3)
        user = self.user_emb(uid.get_feature_data(data))
        logit = self.out(self.mlp(torch.cat([user, q, pooled], dim=1))).squeeze(-1)
        t = lab.get_feature_data(data)
        return logit, (t.float() if t is not None else None)
