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
