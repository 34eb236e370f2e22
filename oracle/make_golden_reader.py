"""Generate tests/golden/reference_reader.npz by EXECUTING THE UNMODIFIED REFERENCE READERS.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run here (CPU container), never on the GPU box:
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_reader.py

Pins N2: the batches ``DataLoader(TrainDataset(reader), shuffle=True)`` collates from
``SimpleDataReader.get_train_dataset_item`` (torchrec/data/SimpleDataReader.py:323-331), the dev batches with
their top-k candidate lists (:333-341), ``SVDPPDataReader``'s per-user ``iids`` (torchrec/data/SVDPPDataReader.py:97-104),
``HistoryDataReader``'s list columns (torchrec/data/HistoryDataReader.py:54-62), and the pair-wise negative
sampler ``train_neg_sample`` (SimpleDataReader.py:280-300) over several epochs of one ``Generator``.

The readers' constructors read a dataset directory (feather / npy / pkl files produced by the offline
preprocessing, out of scope); the instances here are created with ``object.__new__`` and given the state
``_load_dataset`` would leave (frames, split, candidate arrays, positive sets), after which every method
called is the reference's own, unmodified.
"""
import contextlib
import io
import os
import sys

import numpy as np
import pandas as pd
import torch
from numpy.random import default_rng
from torch.utils.data import DataLoader

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden", "reference_reader.npz")
REF = "/root/reference"

sys.dont_write_bytecode = True


def _import_reference():
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):  # utils/const.py prints the hostname at import
        from torchrec.data.SimpleDataReader import SimpleDataReader
        from torchrec.data.SVDPPDataReader import SVDPPDataReader
        from torchrec.data.HistoryDataReader import HistoryDataReader
        from torchrec.data.adapter.TrainDataset import TrainDataset
        from torchrec.data.adapter.DevDataset import DevDataset
        from torchrec.data.adapter.TestDataset import TestDataset
        from torchrec.data.dataset import SplitMode
        from torchrec.task import TrainMode
    return dict(Simple=SimpleDataReader, SVDPP=SVDPPDataReader, History=HistoryDataReader, Train=TrainDataset,
                Dev=DevDataset, Test=TestDataset, SplitMode=SplitMode, TrainMode=TrainMode)


def make_tables(seed, n, n_users, n_items, with_float=False, with_negatives=False):
    rng = np.random.default_rng(seed)
    df = pd.DataFrame({
        "uid": rng.integers(1, n_users + 1, n).astype(np.int32),
        "iid": rng.integers(1, n_items + 1, n).astype(np.int32),
        "rate": rng.integers(1, 6, n).astype(np.int32),
        "label": (rng.random(n) < 0.8).astype(np.int32) if with_negatives else np.ones(n, dtype=np.int32),
        "time": np.arange(n, dtype=np.int32),
        "c_c_hour": rng.integers(1, 25, n).astype(np.int32),
    })
    if with_float:
        df["c_n_price"] = rng.random(n).astype(np.float32)
    items = pd.DataFrame({
        "iid": np.arange(1, n_items + 1, dtype=np.int32),
        "i_c_cat": rng.integers(1, 6, n_items).astype(np.int32),
        "i_c_brand": rng.integers(1, 9, n_items).astype(np.int32),
    })
    return df, items


def positives(df):
    pos = {}
    for u, i, l in zip(df.uid.values, df.iid.values, df.label.values):
        s = pos.setdefault(int(u), set())
        if l == 1:
            s.add(int(i))
    return pos


def save_frame(rec, prefix, df):
    for c in df.columns:
        v = df[c].values
        rec[f"{prefix}/{c}"] = np.stack(v) if v.dtype == object else v.copy()


def save_sets(rec, prefix, pos):
    us = sorted(pos)
    rec[f"{prefix}/uids"] = np.array(us, dtype=np.int64)
    rec[f"{prefix}/offsets"] = np.cumsum([0] + [len(pos[u]) for u in us]).astype(np.int64)
    rec[f"{prefix}/iids"] = np.array([i for u in us for i in sorted(pos[u])], dtype=np.int64)


def save_batches(rec, prefix, loader):
    nb = 0
    for b, batch in enumerate(loader):
        for k, v in batch.items():
            rec[f"{prefix}/batch{b}/{k}"] = v.numpy().copy()
        nb += 1
    rec[f"{prefix}/n_batches"] = np.int64(nb)


def bare(cls, R, df, items, train_mode, split_mode, n_train, n_dev, seed):
    """An instance in the state ``_load_dataset`` leaves (SimpleDataReader.py:150-160), without the disk layout."""
    r = object.__new__(cls)
    r.dataset = "golden"
    r.train_mode, r.split_mode = train_mode, split_mode
    r.random_seed = seed
    r.rng = default_rng(seed)
    r.load_feature, r.append_id = True, True
    r.interaction_df, r.item_df = df, items
    r.train_df = df.iloc[:n_train]
    r.dev_df = df.iloc[n_train:n_train + n_dev]
    r.test_df = df.iloc[n_train + n_dev:]
    r.feature_column_dict = {}
    return r


def main():
    R = _import_reference()
    TM, SM = R["TrainMode"], R["SplitMode"]
    rec = {"torch_version": np.array(torch.__version__), "numpy_version": np.array(np.__version__),
           "pandas_version": np.array(pd.__version__)}

    # ---- case 1: pair-wise SimpleDataReader, leave-k-out candidates, dense positives (many collisions)
    n, n_train, n_dev, n_users, n_items, n_neg = 120, 90, 15, 9, 24, 5
    df, items = make_tables(11, n, n_users, n_items)
    r = bare(R["Simple"], R, df, items, TM.PAIR_WISE, SM.LEAVE_K_OUT, n_train, n_dev, 2020)
    rng = np.random.default_rng(5)
    r.dev_iid_topk_array = np.hstack((r.dev_df.iid.values.reshape(-1, 1),
                                      rng.integers(1, n_items + 1, (n_dev, n_neg)).astype(np.int32)))
    r.test_iid_topk_array = np.hstack((r.test_df.iid.values.reshape(-1, 1),
                                       rng.integers(1, n_items + 1, (len(r.test_df), n_neg)).astype(np.int32)))
    # what _prepare_train_neg_sample sets up (SimpleDataReader.py:253-278) minus the pickle read
    r.min_iid_array_index, r.max_iid_array_index = 1, int(items.iid.max()) + 1
    r.train_df = r.train_df[r.train_df.label == 1]
    r.user_pos_his_set_dict = positives(df)
    r.train_iid_pair_array = np.hstack((r.train_df.iid.values.reshape(-1, 1),
                                        np.empty_like(r.train_df.iid.values).reshape(-1, 1)))
    tag = "pair"
    save_frame(rec, f"{tag}/train", r.train_df)
    save_frame(rec, f"{tag}/dev", r.dev_df)
    save_frame(rec, f"{tag}/test", r.test_df)
    save_frame(rec, f"{tag}/items", items)
    save_sets(rec, f"{tag}/pos", r.user_pos_his_set_dict)
    rec[f"{tag}/dev_topk"], rec[f"{tag}/test_topk"] = r.dev_iid_topk_array, r.test_iid_topk_array
    rec[f"{tag}/seed"] = np.int64(2020)
    for epoch in range(3):
        with contextlib.redirect_stderr(io.StringIO()):  # tqdm
            r.train_neg_sample()
        rec[f"{tag}/epoch{epoch}/pairs"] = r.train_iid_pair_array.copy()
        torch.manual_seed(100 + epoch)
        save_batches(rec, f"{tag}/epoch{epoch}", DataLoader(R["Train"](r), batch_size=16, shuffle=True))
        rec[f"{tag}/epoch{epoch}/rng_after"] = torch.rand(4).numpy()  # the loader's draws from the global generator
    torch.manual_seed(1)
    save_batches(rec, f"{tag}/dev", DataLoader(R["Dev"](r), batch_size=4))
    save_batches(rec, f"{tag}/test", DataLoader(R["Test"](r), batch_size=7))
    torch.manual_seed(2)
    save_batches(rec, f"{tag}/droplast", DataLoader(R["Train"](r), batch_size=16, shuffle=True, drop_last=True))

    # ---- case 2: point-wise, sequential split, mixed int / float frame (row-wise iloc upcasts to float64)
    df, items = make_tables(12, 60, 7, 15, with_float=True, with_negatives=True)
    r = bare(R["Simple"], R, df, items, TM.POINT_WISE, SM.SEQUENTIAL_SPLIT, 40, 10, 2020)
    tag = "point"
    save_frame(rec, f"{tag}/train", r.train_df)
    save_frame(rec, f"{tag}/dev", r.dev_df)
    save_frame(rec, f"{tag}/items", items)
    torch.manual_seed(3)
    save_batches(rec, f"{tag}/train", DataLoader(R["Train"](r), batch_size=12, shuffle=True))
    save_batches(rec, f"{tag}/dev", DataLoader(R["Dev"](r), batch_size=12))

    # ---- case 3: SVDPPDataReader (per-user padded / cut history), point-wise
    df, items = make_tables(13, 80, 6, 30)
    r = bare(R["SVDPP"], R, df, items, TM.POINT_WISE, SM.SEQUENTIAL_SPLIT, 60, 10, 2020)
    r.limit, r.train_all_his_dict = 8, {}
    r.dev_df = r.dev_df[r.dev_df.uid.isin(r.train_df.uid)]  # cold users have no train history (the split guarantees it)
    r._create_user_all_history()
    tag = "svdpp"
    save_frame(rec, f"{tag}/train", r.train_df)
    save_frame(rec, f"{tag}/dev", r.dev_df)
    rec[f"{tag}/limit"] = np.int64(8)
    us = sorted(r.train_all_his_dict)
    rec[f"{tag}/his_uids"] = np.array(us, dtype=np.int64)
    rec[f"{tag}/his"] = np.stack([r.train_all_his_dict[u] for u in us])
    torch.manual_seed(4)
    save_batches(rec, f"{tag}/train", DataLoader(R["Train"](r), batch_size=16, shuffle=True))
    save_batches(rec, f"{tag}/dev", DataLoader(R["Dev"](r), batch_size=16))

    # ---- case 4: HistoryDataReader list columns (what _load_history adds, HistoryDataReader.py:54-62)
    df, items = make_tables(14, 50, 5, 12)
    L = 6
    rng = np.random.default_rng(6)
    lens = rng.integers(0, L + 1, len(df))
    mix = np.zeros((len(df), L + 1), dtype=np.int32)
    mix[:, 0] = lens
    for i, l in enumerate(lens):
        mix[i, 1:1 + l] = rng.integers(1, 13, l)
    df["pos_his_len"] = mix[:, 0].clip(min=1)
    df["pos_his"] = list(mix[:, 1:])
    r = bare(R["History"], R, df, items, TM.POINT_WISE, SM.SEQUENTIAL_SPLIT, 40, 5, 2020)
    tag = "history"
    save_frame(rec, f"{tag}/train", r.train_df)
    torch.manual_seed(5)
    save_batches(rec, f"{tag}/train", DataLoader(R["Train"](r), batch_size=16, shuffle=True))

    np.savez_compressed(OUT, **rec)
    print(f"wrote {OUT}: {len(rec)} arrays, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
