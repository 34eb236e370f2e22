"""Feature-interaction layers (fused CUDA forward + backward)."""
import torch
from torch import Tensor, nn

from ... import ops


class FMSecondOrder(nn.Module):
    """``y[b] = 0.5 * sum_k ((sum_f v[b,f,k])^2 - sum_f v[b,f,k]^2)`` for ``v [B, F, D]``.
    The F = 2 case is the reference's matrix-factorisation dot ``(u * i).sum(-1)``
    (torchrec/model/FunkSVD.py:51,62)."""

    def forward(self, v: Tensor) -> Tensor:
        return ops.fm2(v)


class CrossNet(nn.Module):
    """DCN-v2 cross network: ``x_{l+1} = x0 * (x_l W_l^T + b_l) + x_l`` for ``l < num_layers``.
    Parameters are ``layers.{l}.weight [d, d]`` / ``layers.{l}.bias [d]`` (``nn.Linear`` modules, so the
    reference's ``_reset_weights_fn`` initialises them).  The contraction runs in bf16 on the tcgen05 tensor
    cores with fp32 accumulation and a fused ``x0 * u + x`` epilogue; fp32 master weights."""

    def __init__(self, input_units: int, num_layers: int):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(input_units, input_units) for _ in range(num_layers)])

    def forward(self, x0: Tensor) -> Tensor:
        return ops.cross_net(x0, [m.weight for m in self.layers], [m.bias for m in self.layers])


class AttentionPooling(nn.Module):
    """DIN attention pooling: ``pooled = sum_{l < len} a_l k_l`` with the activation unit
    ``a_l = fc3(relu(fc2(relu(fc1([q, k_l, q - k_l, q * k_l])))))`` (no softmax).  Parameters are three
    ``nn.Linear`` modules (``fc1: 4*dim -> hidden[0]``, ``fc2``, ``fc3: hidden[1] -> 1``) so the reference's init
    applies; forward and backward are one fused CUDA kernel each."""

    def __init__(self, dim: int, hidden=(80, 40)):
        super().__init__()
        self.dim = dim
        self.fc1 = nn.Linear(4 * dim, hidden[0])
        self.fc2 = nn.Linear(hidden[0], hidden[1])
        self.fc3 = nn.Linear(hidden[1], 1)

    def forward(self, q: Tensor, keys: Tensor, lens: Tensor = None) -> Tensor:
        return ops.din_attn_pool(q, keys, lens, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                                 self.fc3.weight, self.fc3.bias)
