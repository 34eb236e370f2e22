"""CPU restatement of the row-wise sharding plan (test oracle; see oracle/__init__.py).
No reference counterpart exists (the reference is single-device, torchrec/task/Task.py:187-190): this
file defines the integer artefacts the CUDA pack kernels must reproduce bit-exactly."""
from typing import Tuple

import torch
from torch import Tensor


def pack_by_owner_ref(ids: Tensor, G: int, C: int) -> Tuple[Tensor, Tensor, int]:
    """ids [F, B] -> (send_ids [G, F, C] with -1 padding, ret_pos [F, B] int32, max list length).
    owner = id mod G, local row = id div G, slots assigned in batch order (stable)."""
    F, B = ids.shape
    send = torch.full((G, F, C), -1, dtype=torch.int64)
    ret = torch.full((F, B), -1, dtype=torch.int32)
    longest = 0
    for f in range(F):
        fill = [0] * G
        for b in range(B):
            i = int(ids[f, b])
            if i < 0:
                continue
            d = i % G
            s = fill[d]
            fill[d] += 1
            if s < C:
                send[d, f, s] = i // G
                ret[f, b] = (d * F + f) * C + s
        longest = max(longest, max(fill))
    return send, ret, longest


def owner_lookup_ref(recv_ids: Tensor, local_tables) -> Tensor:
    """recv_ids [G_src, F, C] local rows (-1 = empty) -> rows [G_src, F, C, D] (zeros for empty slots)."""
    G, F, C = recv_ids.shape
    D = local_tables[0].shape[1]
    out = torch.zeros(G, F, C, D, dtype=local_tables[0].dtype)
    for f in range(F):
        idx = recv_ids[:, f, :]
        ok = idx >= 0
        out[:, f][ok] = local_tables[f][idx[ok]]
    return out


def unpack_ref(recv_rows: Tensor, ret_pos: Tensor) -> Tensor:
    """recv_rows [G, F, C, D] flattened by slot, ret_pos [F, B] -> [B, F, D]."""
    F, B = ret_pos.shape
    D = recv_rows.shape[-1]
    flat = recv_rows.reshape(-1, D)
    out = torch.zeros(B, F, D, dtype=recv_rows.dtype)
    for f in range(F):
        ok = ret_pos[f] >= 0
        out[ok, f] = flat[ret_pos[f][ok].long()]
    return out
