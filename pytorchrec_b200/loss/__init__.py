"""Losses with the reference's registry convention (torchrec/loss/losses.py:8-21, BPRLoss.py:15-23,
Top1Loss.py:14-22).  ``bce`` (``BCEWithLogitsLoss``) is added for the CTR models; it is a ``_Loss`` so
``IModel.compile``'s isinstance check accepts it."""
from typing import Dict, Type

import torch
import torch.nn.functional as F  # noqa
from torch.nn.modules.loss import BCEWithLogitsLoss, MSELoss, _Loss  # noqa


class _PairLoss(_Loss):
    def __init__(self, reduction='mean'):
        super().__init__(None, None, reduction)

    def _reduce(self, x):
        if self.reduction == 'none':
            return x
        return x.mean() if self.reduction == 'mean' else x.sum()

    @staticmethod
    def _split(input):
        assert input.dim() == 2 and input.shape[1] == 2, input.shape
        return input[:, 0], input[:, 1]


class BPRLoss(_PairLoss):
    """softplus(-(pos - neg)) on ``[B, 2]`` scores."""

    def forward(self, input: torch.Tensor, target: torch.Tensor):
        pos, neg = self._split(input)
        return self._reduce(F.softplus(neg - pos))


class Top1Loss(_PairLoss):
    """sigmoid(neg - pos) + sigmoid(neg^2) on ``[B, 2]`` scores."""

    def forward(self, input: torch.Tensor, target: torch.Tensor):
        pos, neg = self._split(input)
        return self._reduce(torch.sigmoid(neg - pos) + torch.sigmoid(neg * neg))


_loss_classes: Dict[str, Type[_Loss]] = {"bpr": BPRLoss, "top1": Top1Loss, "mse": MSELoss, "bce": BCEWithLogitsLoss}
loss_name_list = _loss_classes.keys()


def get_loss(loss_name: str) -> Type[_Loss]:
    if (not isinstance(loss_name, str)) or (loss_name not in _loss_classes):
        raise ValueError(f"invalid loss_name: {loss_name}")
    return _loss_classes[loss_name]
