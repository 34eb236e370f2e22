"""N2 — tensor-native data reader: whole batches assembled by index, no per-row Python.

The reference builds a batch one sample at a time: ``DataLoader`` calls ``TrainDataset.__getitem__``
(torchrec/data/adapter/TrainDataset.py:18-22) -> ``SimpleDataReader.get_train_dataset_item``
(torchrec/data/SimpleDataReader.py:323-331), i.e. ``dict(train_df.iloc[index])`` plus, pair-wise, a
``item_df.iloc[pair - 1]`` lookup per sample, then ``default_collate`` stacks B dicts.  Its pair-wise
negative sampler (``train_neg_sample``, SimpleDataReader.py:280-300) is a Python loop over every
training row.  Measured here (8 host cores, 2e5-row frame, batch 4096): 3.1 k samples/s of batch assembly and
0.84 M rows/s of negative sampling — the step itself runs at 2e7 samples/s.

``TensorDataReader`` holds the same state the reference reader holds after ``_load_dataset``
(train / dev / test frames, the item table, the top-k candidate arrays, the pair array, the users'
positive-item sets, the numpy ``Generator``) as tensors, on the host (pinned) or resident in HBM:

* every split is ONE row-major ``[N, W]`` matrix per dtype family (ints / floats): a batch is one
  ``index_select`` per family and the dict values are column views of the result;
* item-side columns (``item_df.iloc[iids - 1]``) are one gather of ``[B, k]`` candidate ids into the
  item matrix; per-user columns (SVDPP's ``iids``, torchrec/data/SVDPPDataReader.py:100-104) one gather
  by ``uid``;
* ``batches()`` reproduces the reference ``DataLoader``'s sample ORDER bit for bit under the same torch
  seed (base-seed draw, then the ``RandomSampler`` seed draw, then ``randperm``), so a run that swaps
  the loader sees the same batches;
* ``train_neg_sample()`` is bit-exact with the reference's draws for the same ``Generator`` state: the
  first vectorised draw is the reference's own, collisions with a user's positives are found with one
  sorted-key search, and only the colliding rows (in row order, as the reference's loop meets them)
  take the scalar redraws.  ``sampler="device"`` draws and rejects on the reader's device instead
  (torch RNG: same distribution, not the same stream).

Values and shapes equal the reference's batches key for key.  Dtypes are the columns' own; the
reference's row-wise ``iloc`` upcasts a mixed int / float frame to float64 before ``default_collate``,
which ``get_feature_data`` undoes with ``.long()`` / ``.float()`` — the tensors a model sees are equal.
"""
from typing import Any, Dict, Iterator, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

INDEX, UID, IID, LABEL = "index", "uid", "iid", "label"  # torchrec/utils/const.py:79-84
PAIR_WISE = "pair_wise"
LEAVE_K_OUT = "leave_k_out"


def _enum_value(x) -> Any:
    return getattr(x, "value", x)


def _as_2d(a: np.ndarray) -> np.ndarray:
    """A frame column as ``[N]`` or ``[N, L]``: object columns of equal-length arrays (the reference stores
    histories that way, torchrec/data/HistoryDataReader.py:62) are stacked."""
    a = np.asarray(a)
    if a.dtype == object:
        a = np.stack([np.asarray(x) for x in a]) if len(a) else np.zeros((0, 0), dtype=np.int32)
    return a


class _ColumnStore:
    """Columns of one table packed into one row-major matrix per dtype family."""

    def __init__(self, columns: Mapping[str, np.ndarray], device: torch.device, pin: bool):
        self.n = 0
        self.keys: List[str] = list(columns.keys())
        self.slots: Dict[str, Tuple[str, int, int, bool, torch.dtype]] = {}  # key -> (family, col0, width, is_2d, dtype)
        fam_cols: Dict[str, List[np.ndarray]] = {"i": [], "f": []}
        fam_w = {"i": 0, "f": 0}
        arrays = {k: _as_2d(v) for k, v in columns.items()}
        if arrays:
            lens = {len(a) for a in arrays.values()}
            if len(lens) != 1:
                raise ValueError(f"columns of unequal length: { {k: len(a) for k, a in arrays.items()} }")
            self.n = lens.pop()
        for k, a in arrays.items():
            if a.ndim not in (1, 2):
                raise ValueError(f"column {k!r}: expected [N] or [N, L], got shape {a.shape}")
            fam = "f" if a.dtype.kind == "f" else "i"
            if a.dtype.kind not in "iufb":
                raise ValueError(f"column {k!r}: unsupported dtype {a.dtype}")
            w = 1 if a.ndim == 1 else a.shape[1]
            self.slots[k] = (fam, fam_w[fam], w, a.ndim == 2, torch.from_numpy(np.zeros(0, dtype=a.dtype)).dtype)
            fam_cols[fam].append(a.reshape(self.n, w))
            fam_w[fam] += w
        self.mats: Dict[str, Tensor] = {}
        for fam, cols in fam_cols.items():
            if not cols:
                continue
            # the family matrix is as wide as its widest member dtype (int32 ids + an int64 column -> int64)
            dt = np.result_type(*[c.dtype for c in cols])
            m = torch.from_numpy(np.ascontiguousarray(np.concatenate([c.astype(dt, copy=False) for c in cols], axis=1)))
            if device.type == "cuda":
                m = m.to(device)
            elif pin and torch.cuda.is_available():
                m = m.pin_memory()
            self.mats[fam] = m

    def gather(self, index: Tensor) -> Dict[str, Tensor]:
        """``{key: column[index]}``; ``index`` may be ``[B]`` or ``[B, k]`` (then values are ``[B, k]`` / ``[B, k, L]``)."""
        flat = index.reshape(-1)
        rows = {fam: m.index_select(0, flat) for fam, m in self.mats.items()}
        out: Dict[str, Tensor] = {}
        for k in self.keys:
            fam, c0, w, is_2d, dt = self.slots[k]
            v = rows[fam][:, c0:c0 + w] if is_2d else rows[fam][:, c0]
            if v.dtype != dt:
                v = v.to(dt)
            out[k] = v.reshape(tuple(index.shape) + ((w,) if is_2d else ()))
        return out


class _PositiveSets:
    """Users' positive-item sets (``user_pos_his_set_dict``, torchrec/data/process/vt_negative_sample.py:24-43)
    as one sorted array of ``uid * stride + iid`` keys: membership of a whole candidate vector is one search."""

    def __init__(self, sets: Mapping[int, Any], stride: int, device: torch.device):
        uids, iids = [], []
        for u, s in sets.items():
            s = np.fromiter(s, dtype=np.int64, count=len(s))
            uids.append(np.full(len(s), int(u), dtype=np.int64))
            iids.append(s)
        u = np.concatenate(uids) if uids else np.zeros(0, np.int64)
        i = np.concatenate(iids) if iids else np.zeros(0, np.int64)
        self.stride = int(max(stride, int(i.max()) + 1 if len(i) else 1))
        self.keys_np = np.unique(u * self.stride + i)
        self.keys_cpu = torch.from_numpy(self.keys_np)
        self.keys = self.keys_cpu.to(device)

    def contains_np(self, uid: np.ndarray, iid: np.ndarray) -> np.ndarray:
        """Host arrays in, host mask out (torch's multi-threaded search: ~7x numpy's at 1e6 queries)."""
        return self.contains(torch.from_numpy(np.ascontiguousarray(uid)), torch.from_numpy(np.ascontiguousarray(iid)),
                             self.keys_cpu).numpy()

    def contains(self, uid: Tensor, iid: Tensor, keys: Optional[Tensor] = None) -> Tensor:
        keys = self.keys if keys is None else keys
        if keys.numel() == 0:
            return torch.zeros_like(uid, dtype=torch.bool)
        q = uid.to(torch.int64) * self.stride + iid.to(torch.int64)
        pos = torch.searchsorted(keys, q).clamp_(max=keys.numel() - 1)
        return keys[pos] == q


class SplitDataset:
    """One split of a ``TensorDataReader`` with the reference adapters' surface (``__len__`` / ``__getitem__``,
    ``train_neg_sample`` on the train split — torchrec/data/adapter/{Train,Dev,Test}Dataset.py) plus ``batches``,
    the vectorised loader ``IModel.fit`` / ``evaluate`` / ``predict`` use when they find it."""

    def __init__(self, reader: "TensorDataReader", split: str):
        self.reader, self.split = reader, split

    def __len__(self) -> int:
        return self.reader.size(self.split)

    def __getitem__(self, item: int) -> Dict[str, Any]:
        b = self.reader.get_batch(self.split, torch.tensor([int(item)], dtype=torch.int64))
        return {k: (v[0].item() if v[0].dim() == 0 else v[0].cpu().numpy()) for k, v in b.items()}

    def train_neg_sample(self) -> None:
        self.reader.train_neg_sample()

    def batches(self, batch_size: int, shuffle: bool = False, drop_last: bool = False, rank: int = 0,
                world_size: int = 1) -> Iterator[Dict[str, Tensor]]:
        return self.reader.batches(self.split, batch_size, shuffle=shuffle, drop_last=drop_last, rank=rank,
                                   world_size=world_size)


class TensorDataReader:
    """See the module docstring.  Frames are ``Mapping[str, ndarray]`` (``[N]`` or ``[N, L]`` columns); ``items`` is
    the item table in ``iid`` order (row ``i`` describes item ``i + 1``; 0 is PAD — SimpleDataReader.py:327)."""

    def __init__(self,
                 train: Mapping[str, np.ndarray],
                 dev: Optional[Mapping[str, np.ndarray]] = None,
                 test: Optional[Mapping[str, np.ndarray]] = None,
                 items: Optional[Mapping[str, np.ndarray]] = None,
                 *,
                 train_mode: Any = "point_wise",
                 split_mode: Any = "sequential_split",
                 dev_iid_topk: Optional[np.ndarray] = None,
                 test_iid_topk: Optional[np.ndarray] = None,
                 user_pos_his_set_dict: Optional[Mapping[int, Any]] = None,
                 per_user: Optional[Mapping[str, np.ndarray]] = None,
                 feature_column_dict: Optional[Dict[str, Any]] = None,
                 rng: Optional[np.random.Generator] = None,
                 random_seed: int = 2020,
                 device: Optional[torch.device] = None,
                 sampler: str = "reference",
                 pin_memory: bool = True):
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.train_mode = _enum_value(train_mode)
        self.split_mode = _enum_value(split_mode)
        if sampler not in ("reference", "device"):
            raise ValueError(f"invalid sampler: {sampler}")
        self.sampler = sampler
        self.rng = rng if rng is not None else np.random.default_rng(random_seed)
        self.random_seed = random_seed
        self.feature_column_dict = feature_column_dict if feature_column_dict is not None else {}

        train = dict(train)
        if self.train_mode == PAIR_WISE:
            # the reference drops the negatives of the train frame first (SimpleDataReader.py:262)
            if LABEL in train:
                keep = np.asarray(train[LABEL]) == 1
                if not keep.all():
                    train = {k: _as_2d(v)[keep] for k, v in train.items()}
        self._stores: Dict[str, _ColumnStore] = {"train": _ColumnStore(train, self.device, pin_memory)}
        for name, frame in (("dev", dev), ("test", test)):
            if frame is not None:
                self._stores[name] = _ColumnStore(frame, self.device, pin_memory)
        self._items = _ColumnStore(items, self.device, pin_memory) if items is not None else None
        self._per_user = _ColumnStore(per_user, self.device, pin_memory) if per_user else None

        self._topk: Dict[str, Tensor] = {}
        if self.split_mode == LEAVE_K_OUT:
            for name, arr in (("dev", dev_iid_topk), ("test", test_iid_topk)):
                if name in self._stores:
                    if arr is None:
                        raise ValueError(f"split_mode leave_k_out needs {name}_iid_topk")
                    if len(arr) != self._stores[name].n:
                        raise ValueError(f"{name}_iid_topk has {len(arr)} rows, the {name} frame {self._stores[name].n}")
                    self._topk[name] = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)
            if len({t.shape[1] for t in self._topk.values()}) > 1:
                raise ValueError("dev and test candidate lists differ in width")  # SimpleDataReader.py:248

        self._pair: Optional[Tensor] = None
        self._pos: Optional[_PositiveSets] = None
        if self.train_mode == PAIR_WISE:
            if self._items is None or IID not in self._items.slots:
                raise ValueError("pair-wise training needs the item table (with its iid column)")
            if user_pos_his_set_dict is None:
                raise ValueError("pair-wise training needs user_pos_his_set_dict")
            iid = np.asarray(train[IID])
            self._train_uid_np = np.asarray(train[UID])
            self._train_uid = torch.from_numpy(np.ascontiguousarray(self._train_uid_np)).to(self.device)
            self.min_iid_array_index = 1                                  # 0 is PAD (SimpleDataReader.py:256)
            self.max_iid_array_index = int(np.asarray(items[IID]).max()) + 1
            self._pos = _PositiveSets(user_pos_his_set_dict, self.max_iid_array_index, self.device)
            self._pair_np = np.stack([iid, np.zeros_like(iid)], axis=1)  # column 1 is filled by train_neg_sample
            self._pair = torch.from_numpy(self._pair_np.copy()).to(self.device)
            self._neg_sampled = False

    # ------------------------------------------------------------------ adoption of a reference-style reader
    @classmethod
    def from_reference_reader(cls, reader, device: Optional[torch.device] = None, sampler: str = "reference",
                              pin_memory: bool = True) -> "TensorDataReader":
        """Adopt the state of a constructed reference reader (``SimpleDataReader`` / ``HistoryDataReader`` /
        ``SVDPPDataReader`` — duck-typed on the attributes its ``_load_dataset`` leaves behind).  The numpy
        ``Generator`` is shared, so negative sampling continues the reader's own random stream."""

        def frame(df):
            return None if df is None else {c: df[c].values for c in df.columns}

        per_user = None
        his = getattr(reader, "train_all_his_dict", None)
        if his:  # SVDPPDataReader: uid -> padded / cut history (SVDPPDataReader.py:88-95)
            limit = len(next(iter(his.values())))
            table = np.zeros((max(int(u) for u in his) + 1, limit), dtype=next(iter(his.values())).dtype)
            for u, a in his.items():
                table[int(u)] = a
            per_user = {"iids": table}
        train = frame(reader.train_df)
        pair = getattr(reader, "train_iid_pair_array", None)
        out = cls(train, frame(getattr(reader, "dev_df", None)), frame(getattr(reader, "test_df", None)),
                  frame(getattr(reader, "item_df", None)),
                  train_mode=reader.train_mode, split_mode=reader.split_mode,
                  dev_iid_topk=getattr(reader, "dev_iid_topk_array", None),
                  test_iid_topk=getattr(reader, "test_iid_topk_array", None),
                  user_pos_his_set_dict=getattr(reader, "user_pos_his_set_dict", None),
                  per_user=per_user, feature_column_dict=reader.get_feature_column_dict()
                  if hasattr(reader, "get_feature_column_dict") else None,
                  rng=reader.rng, random_seed=getattr(reader, "random_seed", 2020), device=device, sampler=sampler,
                  pin_memory=pin_memory)
        if pair is not None and out._pair is not None and len(pair) == len(out._pair_np):
            out._pair_np[:, 1] = pair[:, 1]   # whatever the reader had sampled so far
            out._pair = torch.from_numpy(out._pair_np.copy()).to(out.device)
        return out

    # ------------------------------------------------------------------ IDataReader surface (torchrec/data/IDataReader.py)
    def get_feature_column_dict(self) -> Dict[str, Any]:
        return self.feature_column_dict

    def size(self, split: str) -> int:
        return self._stores[split].n

    def get_train_dataset_size(self) -> int:
        return self.size("train")

    def get_dev_dataset_size(self) -> int:
        return self.size("dev")

    def get_test_dataset_size(self) -> int:
        return self.size("test")

    def get_train_dataset_item(self, index: int) -> Dict[str, Any]:
        return SplitDataset(self, "train")[index]

    def get_dev_dataset_item(self, index: int) -> Dict[str, Any]:
        return SplitDataset(self, "dev")[index]

    def get_test_dataset_item(self, index: int) -> Dict[str, Any]:
        return SplitDataset(self, "test")[index]

    def train_dataset(self) -> SplitDataset:
        return SplitDataset(self, "train")

    def dev_dataset(self) -> SplitDataset:
        return SplitDataset(self, "dev")

    def test_dataset(self) -> SplitDataset:
        return SplitDataset(self, "test")

    # ------------------------------------------------------------------ negative sampling
    def train_neg_sample(self) -> None:
        """Fill column 1 of the pair array with one negative per training row that the row's user has not
        interacted with (SimpleDataReader.py:280-300)."""
        if self.train_mode != PAIR_WISE:
            raise AssertionError("train_neg_sample needs train_mode pair_wise")  # the reference asserts (:282)
        n = len(self._pair_np)
        lo, hi = self.min_iid_array_index, self.max_iid_array_index
        if self.sampler == "reference":
            neg = self.rng.integers(low=lo, high=hi, size=n, dtype=np.int32)
            hit = np.flatnonzero(self._pos.contains_np(self._train_uid_np, neg))
            for i in hit:  # ascending row order = the order in which the reference's loop reaches them
                u = self._train_uid_np[i:i + 1]
                while True:
                    neg[i] = self.rng.integers(low=lo, high=hi, dtype=np.int32)
                    if not self._pos.contains_np(u, neg[i:i + 1])[0]:
                        break
            self._pair_np[:, 1] = neg
            self._pair = torch.from_numpy(self._pair_np.copy()).to(self.device)
        else:
            gen = self._device_generator()
            neg = torch.randint(lo, hi, (n,), generator=gen, device=self.device, dtype=torch.int64)
            todo = self._pos.contains(self._train_uid, neg).nonzero().flatten()
            while todo.numel():  # one host sync per rejection round; rounds ~ log(n) / log(1 / p_collision)
                neg[todo] = torch.randint(lo, hi, (todo.numel(),), generator=gen, device=self.device, dtype=torch.int64)
                todo = todo[self._pos.contains(self._train_uid[todo], neg[todo])]
            self._pair[:, 1] = neg.to(self._pair.dtype)
        self._neg_sampled = True

    def _device_generator(self) -> torch.Generator:
        g = getattr(self, "_gen", None)
        if g is None:
            g = self._gen = torch.Generator(device=self.device)
            g.manual_seed(int(self.random_seed))
        return g

    @property
    def train_iid_pair_array(self) -> Optional[np.ndarray]:
        return None if self._pair is None else self._pair.cpu().numpy()

    # ------------------------------------------------------------------ batch assembly
    def get_batch(self, split: str, index: Tensor) -> Dict[str, Tensor]:
        """The batch the reference's ``DataLoader`` would collate from rows ``index`` of ``split``."""
        store = self._stores[split]
        index = index.to(device=self.device, dtype=torch.int64)
        batch = store.gather(index)
        batch[INDEX] = index
        cand = None
        if split == "train":
            if self.train_mode == PAIR_WISE:
                cand = self._pair.index_select(0, index)
        elif self.split_mode == LEAVE_K_OUT:
            cand = self._topk[split].index_select(0, index)
        if cand is not None:
            if self._items is None:
                raise RuntimeError("candidate lists need the item table")
            batch.update(self._items.gather(cand.to(torch.int64) - 1))  # iid -> item row; replaces the scalar iid
        if self._per_user is not None:
            batch.update(self._per_user.gather(batch[UID].to(torch.int64)))
        return batch

    def batches(self, split: str, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                generator: Optional[torch.Generator] = None, rank: int = 0, world_size: int = 1
                ) -> Iterator[Dict[str, Tensor]]:
        """Iterate ``split`` in the order of ``DataLoader(dataset, batch_size, shuffle=shuffle, drop_last=drop_last)``:
        the loader's iterator draws its base seed from the global CPU generator, ``RandomSampler`` then draws the
        seed of a private generator and takes ``randperm(n)`` from it — the same draws are made here, so the same
        torch seed gives the same batches (and leaves the global generator in the same state).

        ``world_size > 1`` (data-parallel ranks of the row-wise-sharded models, one process per GPU): every rank makes
        the same draws (same seed on every rank, as ``set_torch_seed`` arranges) and takes the strided slice
        ``order[rank::world_size]`` of the epoch's order, trimmed so that all ranks see the same number of samples —
        the union over ranks of step k's batches is the reference loader's batch k of size ``world_size * batch_size``,
        so a sharded run consumes the reference's sample stream."""
        n = self.size(split)
        if not (0 <= rank < world_size):
            raise ValueError(f"rank {rank} outside world of {world_size}")
        torch.empty((), dtype=torch.int64).random_(generator=generator)  # _BaseDataLoaderIter's base seed
        if shuffle:
            g = torch.Generator()
            g.manual_seed(int(torch.empty((), dtype=torch.int64).random_(generator=generator).item()))
            order = torch.randperm(n, generator=g)
        else:
            order = torch.arange(n)
        if world_size > 1:
            order = order[:n - n % world_size][rank::world_size]
            n = order.numel()
        order = order.to(self.device)
        stop = n - n % batch_size if drop_last else n
        for s in range(0, stop, batch_size):
            yield self.get_batch(split, order[s:min(s + batch_size, stop)])
