"""Synthetic Criteo-shaped batches in the reference's wire format: ``Dict[str, Tensor]`` keyed by
feature name (torchrec/utils/const.py:79-98), one ``[B]`` id tensor per sparse field, one ``[B]`` float
per dense field, a ``label`` column.  Ids are drawn in ``[1, rows)`` (0 is PAD in the reference)."""
from typing import Dict, List, Tuple

import numpy as np
import torch

from ..feature_column import CategoricalColumnWithIdentity, NumericColumn


def criteo_columns(n_sparse: int = 26, n_dense: int = 13, rows: int = 1_000_000):
    sparse = [CategoricalColumnWithIdentity(rows, f"C{i + 1}") for i in range(n_sparse)]
    dense = [NumericColumn(f"I{i + 1}", 0.0, 1.0, 0.5, 0.2887) for i in range(n_dense)]
    label = CategoricalColumnWithIdentity(2, "label")
    return sparse, dense, label


def criteo_batch(batch: int, n_sparse: int, n_dense: int, rows: int, seed: int, dist: str = "uniform",
                 pin: bool = False, id_dtype=torch.int64) -> Dict[str, torch.Tensor]:
    """Host batch.  ``dist``: 'uniform' (worst case for caches / dedup) or 'zipf' (alpha = 1.05)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, torch.Tensor] = {}
    for i in range(n_sparse):
        if dist == "zipf":
            x = 1 + (rng.zipf(1.05, size=batch) - 1) % (rows - 1)
        else:
            x = rng.integers(1, rows, size=batch)
        out[f"C{i + 1}"] = torch.from_numpy(x.astype(np.int64)).to(id_dtype)
    for i in range(n_dense):
        out[f"I{i + 1}"] = torch.from_numpy(rng.random(batch, dtype=np.float32))
    out["label"] = torch.from_numpy((rng.random(batch) < 0.25).astype(np.int64))
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


# Amazon-Books-shaped statistics used for the DIN config (BASELINE.json configs[3]; SURVEY.md 8d: the published
# DIN / DIEN dataset sizes — the reference ships no Books loader)
AMAZON_BOOKS = dict(users=603_668, items=367_982, cates=1_600)


def amazon_columns(L: int = 100, users: int = AMAZON_BOOKS["users"], items: int = AMAZON_BOOKS["items"],
                   cates: int = AMAZON_BOOKS["cates"]):
    """(uid, iid, cid, his_iid, his_cid, his_len, label) columns of the DIN config: right-padded ``[B, L]`` histories
    with 0 = PAD and a length column clipped to >= 1 (HistoryDataReader.py:55-69)."""
    C = CategoricalColumnWithIdentity
    return (C(users, "uid"), C(items, "iid"), C(cates, "cid"), C(items, "his_iid"), C(cates, "his_cid"),
            C(L + 1, "his_len"), C(2, "label"))


def amazon_batch(batch: int, L: int = 100, seed: int = 0, users: int = AMAZON_BOOKS["users"],
                 items: int = AMAZON_BOOKS["items"], cates: int = AMAZON_BOOKS["cates"],
                 pin: bool = False) -> Dict[str, torch.Tensor]:
    """Host batch of the DIN config: history lengths ~ U{1..L}, ids uniform in [1, n), padded tail = 0."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, L + 1, size=batch)
    pad = np.arange(L)[None, :] >= lens[:, None]
    hi = rng.integers(1, items, size=(batch, L))
    hi[pad] = 0
    hc = rng.integers(1, cates, size=(batch, L))
    hc[pad] = 0
    out = {"uid": rng.integers(1, users, size=batch), "iid": rng.integers(1, items, size=batch),
           "cid": rng.integers(1, cates, size=batch), "his_iid": hi, "his_cid": hc, "his_len": lens,
           "label": (rng.random(batch) < 0.25).astype(np.int64)}
    out = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in out.items()}
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out
