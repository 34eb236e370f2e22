"""Row-wise restatement of the reference's batch assembly and pair-wise negative sampler (N2 oracle).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned by ``tests/golden/reference_reader.npz``, which
``oracle/make_golden_reader.py`` writes by executing the unmodified reference readers.
"""
from typing import Dict, Mapping, Set

import numpy as np
from torch.utils.data import default_collate


def train_neg_sample_ref(rng: np.random.Generator, uids: np.ndarray, pos: Mapping[int, Set[int]],
                         low: int, high: int) -> np.ndarray:
    """torchrec/data/SimpleDataReader.py:283-299: one vectorised int32 draw, then per row (in order) scalar redraws
    from the same generator while the candidate is one of the user's positives."""
    neg = rng.integers(low=low, high=high, size=len(uids), dtype=np.int32)
    for index, uid in enumerate(uids):
        inter = pos[uid]
        while neg[index] in inter:
            neg[index] = rng.integers(low=low, high=high, dtype=np.int32)
    return neg


def assemble_batch_ref(frame: Dict[str, np.ndarray], index: np.ndarray, items: Dict[str, np.ndarray] = None,
                       cand: np.ndarray = None) -> Dict:
    """torchrec/data/SimpleDataReader.py:323-331 + ``default_collate``: one dict per sample (row values, ``index``,
    item columns gathered at ``cand[i] - 1``), stacked key by key."""
    rows = []
    for i in index:
        d = {k: v[i] for k, v in frame.items()}
        d["index"] = int(i)
        if cand is not None:
            for k, v in items.items():
                d[k] = v[cand[i] - 1]
        rows.append(d)
    return default_collate(rows)
