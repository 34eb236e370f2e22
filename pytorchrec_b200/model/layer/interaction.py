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
