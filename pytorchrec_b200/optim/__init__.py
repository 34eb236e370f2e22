"""Optimizer registry with the reference's calling convention (torchrec/optim/optimizers.py:7-20):
``get_optimizer(name)`` returns a class instantiated as ``cls(params=model.get_parameters(), **kw)``.
The reference's entries (``sgd``, ``adam``, ``adamw``) are kept; ``adagrad`` and the fused
``sparse_*`` optimizers are new."""
from typing import Dict, Type

from torch.optim import SGD, Adagrad, Adam
from torch.optim.optimizer import Optimizer

from .AdamW import AdamW
from .sparse import SparseAdagrad, SparseAdam, SparseRowWiseAdagrad, SparseSGD

_optimizer_classes: Dict[str, Type[Optimizer]] = {
    "sgd": SGD,
    "adam": Adam,
    "adamw": AdamW,
    "adagrad": Adagrad,
    "sparse_sgd": SparseSGD,
    "sparse_adagrad": SparseAdagrad,
    "sparse_rowwise_adagrad": SparseRowWiseAdagrad,
    "sparse_adam": SparseAdam,
}

optimizer_name_list = _optimizer_classes.keys()


def get_optimizer(optimizer_name: str) -> Type[Optimizer]:
    if (not isinstance(optimizer_name, str)) or (optimizer_name not in _optimizer_classes):
        raise ValueError(f"invalid optimizer_name: {optimizer_name}")
    return _optimizer_classes[optimizer_name]


__all__ = ["AdamW", "SparseSGD", "SparseAdagrad", "SparseRowWiseAdagrad", "SparseAdam", "get_optimizer",
           "optimizer_name_list"]
