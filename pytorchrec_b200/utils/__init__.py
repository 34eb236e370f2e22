"""Step-path utilities (reference: torchrec/utils/global_utils.py:7-16, data_structure.py:10-52)."""
from typing import Callable

import numpy as np
import torch
from torch import Tensor


def set_torch_seed(seed: int) -> None:
    """Seed torch (CPU + all CUDA devices) and pin cudnn to deterministic mode — same knobs as the
    reference's ``set_torch_seed`` so that init draws are identical."""
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def map_structure(func: Callable, structure):
    """Apply ``func`` to every leaf of nested lists / dicts."""
    if not callable(func):
        raise TypeError("func must be callable, got: %s" % func)
    if isinstance(structure, list):
        return [map_structure(func, item) for item in structure]
    if isinstance(structure, dict):
        return {key: map_structure(func, value) for key, value in structure.items()}
    return func(structure)


def tensor_to_device(structure, device: torch.device):
    """Move every tensor leaf to ``device``.  Pinned host tensors are copied asynchronously (the
    reference issues one blocking pageable copy per key, data_structure.py:44-52)."""

    def _move(t):
        if isinstance(t, Tensor):
            if t.device == device:
                return t
            return t.to(device=device, non_blocking=t.is_pinned() if t.device.type == "cpu" else False)
        return t

    return map_structure(_move, structure)


def tensor_to_numpy_or_python_type(structure):
    def _conv(t):
        if isinstance(t, Tensor):
            x = t.detach().cpu().numpy()
            return x.item() if np.ndim(x) == 0 else x
        return t

    return map_structure(_conv, structure)
