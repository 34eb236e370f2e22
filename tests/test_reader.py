"""N2 — tensor-native reader against the unmodified reference readers (tests/golden/reference_reader.npz, written
by oracle/make_golden_reader.py) and against the row-wise restatement (oracle/ref_reader.py) at larger sizes.
Everything here is integer / index work: comparisons are bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import ref_reader
from pytorchrec_b200.data import SplitDataset, TensorDataReader


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(GOLDEN, "reference_reader.npz"))


def frame(G, prefix):
    return {k[len(prefix) + 1:]: np.array(G[k]) for k in G.files if k.startswith(prefix + "/") and "/batch" not in k
            and k.count("/") == prefix.count("/") + 1 and not k.endswith("n_batches")}


def sets(G, prefix):
    us, off, ii = G[f"{prefix}/uids"], G[f"{prefix}/offsets"], G[f"{prefix}/iids"]
    return {int(u): set(int(x) for x in ii[off[n]:off[n + 1]]) for n, u in enumerate(us)}


def check_batches(G, prefix, it, exact_dtype=True):
    nb = int(G[f"{prefix}/n_batches"])
    got = list(it)
    assert len(got) == nb
    for b, batch in enumerate(got):
        keys = [k.rsplit("/", 1)[1] for k in G.files if k.startswith(f"{prefix}/batch{b}/")]
        assert sorted(batch.keys()) == sorted(keys)
        for k in keys:
            want = G[f"{prefix}/batch{b}/{k}"]
            have = batch[k].cpu().numpy()
            assert have.shape == want.shape, (b, k)
            if exact_dtype:
                assert have.dtype == want.dtype, (b, k, have.dtype, want.dtype)
            # mixed int / float frames: the reference's row-wise iloc upcasts to float64; the VALUES are equal
            np.testing.assert_array_equal(have.astype(np.float64), want.astype(np.float64), err_msg=f"{b}/{k}")


def pair_reader(G, device=None, sampler="reference"):
    return TensorDataReader(frame(G, "pair/train"), frame(G, "pair/dev"), frame(G, "pair/test"), frame(G, "pair/items"),
                            train_mode="pair_wise", split_mode="leave_k_out", dev_iid_topk=G["pair/dev_topk"],
                            test_iid_topk=G["pair/test_topk"], user_pos_his_set_dict=sets(G, "pair/pos"),
                            rng=np.random.default_rng(int(G["pair/seed"])), device=device, sampler=sampler)


def run_pair_case(G, device=None):
    r = pair_reader(G, device)
    for epoch in range(3):
        r.train_neg_sample()
        np.testing.assert_array_equal(r.train_iid_pair_array, G[f"pair/epoch{epoch}/pairs"])
        torch.manual_seed(100 + epoch)
        check_batches(G, f"pair/epoch{epoch}", r.train_dataset().batches(16, shuffle=True))
        # the loader consumed exactly the reference DataLoader's draws from the global generator
        np.testing.assert_array_equal(torch.rand(4).numpy(), G[f"pair/epoch{epoch}/rng_after"])
    torch.manual_seed(1)
    check_batches(G, "pair/dev", r.dev_dataset().batches(4))
    check_batches(G, "pair/test", r.test_dataset().batches(7))
    torch.manual_seed(2)
    check_batches(G, "pair/droplast", r.train_dataset().batches(16, shuffle=True, drop_last=True))


def test_pairwise_reader_matches_reference_run(G):
    run_pair_case(G)


def test_pointwise_mixed_frame_matches_reference_run(G):
    r = TensorDataReader(frame(G, "point/train"), frame(G, "point/dev"), None, frame(G, "point/items"))
    torch.manual_seed(3)
    check_batches(G, "point/train", r.batches("train", 12, shuffle=True), exact_dtype=False)
    check_batches(G, "point/dev", r.batches("dev", 12), exact_dtype=False)
    b = next(iter(r.batches("train", 12)))
    assert b["uid"].dtype == torch.int32 and b["c_n_price"].dtype == torch.float32  # the columns' own dtypes


def test_svdpp_per_user_history_matches_reference_run(G):
    us, his = G["svdpp/his_uids"], G["svdpp/his"]
    table = np.zeros((int(us.max()) + 1, int(G["svdpp/limit"])), dtype=his.dtype)
    table[us] = his
    r = TensorDataReader(frame(G, "svdpp/train"), frame(G, "svdpp/dev"), per_user={"iids": table})
    torch.manual_seed(4)
    check_batches(G, "svdpp/train", r.batches("train", 16, shuffle=True))
    check_batches(G, "svdpp/dev", r.batches("dev", 16))


def test_history_list_columns_match_reference_run(G):
    r = TensorDataReader(frame(G, "history/train"))
    torch.manual_seed(5)
    # pos_his is an object column of int32 arrays in the reference frame: default_collate gives [B, L] int32
    check_batches(G, "history/train", r.batches("train", 16, shuffle=True))


def test_item_access_and_adapter_surface(G):
    r = pair_reader(G)
    r.train_neg_sample()
    ds = r.train_dataset()
    assert isinstance(ds, SplitDataset) and len(ds) == r.get_train_dataset_size() == len(G["pair/train/uid"])
    item = r.get_train_dataset_item(3)
    assert item["index"] == 3 and item["uid"] == int(G["pair/train/uid"][3])
    np.testing.assert_array_equal(item["iid"], G["pair/epoch0/pairs"][3])
    assert r.get_dev_dataset_item(0)["iid"].shape == G["pair/dev_topk"][0].shape
    with pytest.raises(AssertionError):
        TensorDataReader(frame(G, "point/train")).train_neg_sample()
    with pytest.raises(ValueError):
        TensorDataReader({"uid": np.arange(3), "iid": np.arange(4)})
    with pytest.raises(ValueError):
        TensorDataReader(frame(G, "pair/train"), train_mode="pair_wise")  # no item table / positive sets


def synthetic_pairs(n=200_000, n_users=5000, n_items=300, seed=0):
    rng = np.random.default_rng(seed)
    uid = rng.integers(1, n_users + 1, n).astype(np.int32)
    iid = rng.integers(1, n_items + 1, n).astype(np.int32)
    pos = {}
    for u, i in zip(uid.tolist(), iid.tolist()):
        pos.setdefault(u, set()).add(i)
    train = {"uid": uid, "iid": iid, "label": np.ones(n, dtype=np.int32)}
    items = {"iid": np.arange(1, n_items + 1, dtype=np.int32)}
    return train, items, pos


def test_negative_sampler_bit_exact_with_rowwise_loop_at_scale():
    """2e5 rows, ~12 % first-draw collisions: same negatives as the reference's loop, two epochs of one generator."""
    train, items, pos = synthetic_pairs()
    r = TensorDataReader(train, items=items, train_mode="pair_wise", user_pos_his_set_dict=pos,
                         rng=np.random.default_rng(7))
    ref_rng = np.random.default_rng(7)
    for _ in range(2):
        r.train_neg_sample()
        want = ref_reader.train_neg_sample_ref(ref_rng, train["uid"].tolist(), pos, 1, 301)
        np.testing.assert_array_equal(r.train_iid_pair_array[:, 1], want)
        np.testing.assert_array_equal(r.train_iid_pair_array[:, 0], train["iid"])


def check_device_sampler(device):
    train, items, pos = synthetic_pairs(n=100_000, n_users=50, n_items=40, seed=1)  # users hold ~40 of 40 items?
    # keep every user short of the full catalogue so a negative exists
    for u in pos:
        while len(pos[u]) > 30:
            pos[u].pop()
    keep = np.array([i in pos[u] for u, i in zip(train["uid"].tolist(), train["iid"].tolist())])
    train = {k: v[keep] for k, v in train.items()}
    r = TensorDataReader(train, items=items, train_mode="pair_wise", user_pos_his_set_dict=pos, device=device,
                         sampler="device", random_seed=3)
    r.train_neg_sample()
    pairs = r.train_iid_pair_array
    np.testing.assert_array_equal(pairs[:, 0], train["iid"])
    neg = pairs[:, 1]
    assert neg.min() >= 1 and neg.max() <= 40
    assert not any(int(n) in pos[int(u)] for u, n in zip(train["uid"], neg))
    # uniform over each user's complement: every allowed item of user 1 shows up about equally often
    u = int(train["uid"][0])
    mine = neg[train["uid"] == u]
    allowed = sorted(set(range(1, 41)) - pos[u])
    counts = np.array([(mine == a).sum() for a in allowed])
    assert counts.min() > 0.5 * counts.mean() and counts.max() < 1.5 * counts.mean()
    first = neg.copy()
    r.train_neg_sample()
    assert (r.train_iid_pair_array[:, 1] != first).mean() > 0.5  # a fresh draw every epoch


def test_device_sampler_properties_cpu():
    check_device_sampler(torch.device("cpu"))


def test_rowwise_restatement_matches_reference_run(G):
    """The oracle's row-wise assembly (used for larger property checks) against the reference-run batches."""
    fr, items = frame(G, "pair/train"), frame(G, "pair/items")
    torch.manual_seed(100)
    torch.empty((), dtype=torch.int64).random_()
    g = torch.Generator()
    g.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
    order = torch.randperm(len(fr["uid"]), generator=g).numpy()
    b = ref_reader.assemble_batch_ref(fr, order[:16], items, G["pair/epoch0/pairs"])
    for k, v in b.items():
        np.testing.assert_array_equal(v.numpy(), G[f"pair/epoch0/batch0/{k}"])
    rng = np.random.default_rng(int(G["pair/seed"]))
    neg = ref_reader.train_neg_sample_ref(rng, fr["uid"].tolist(), sets(G, "pair/pos"), 1, int(items["iid"].max()) + 1)
    np.testing.assert_array_equal(neg, G["pair/epoch0/pairs"][:, 1])


def test_batches_equal_rowwise_assembly_at_scale():
    rng = np.random.default_rng(2)
    n, n_items, k = 50_000, 1000, 20
    fr = {"uid": rng.integers(1, 999, n).astype(np.int32), "iid": rng.integers(1, n_items + 1, n).astype(np.int32),
          "his": rng.integers(0, n_items + 1, (n, 10)).astype(np.int32), "x": rng.random(n).astype(np.float32)}
    items = {"iid": np.arange(1, n_items + 1, dtype=np.int32), "cat": rng.integers(1, 30, n_items).astype(np.int32)}
    topk = rng.integers(1, n_items + 1, (n, k)).astype(np.int32)
    r = TensorDataReader({"uid": fr["uid"][:1]}, dev=fr, items=items, split_mode="leave_k_out", dev_iid_topk=topk)
    idx = torch.from_numpy(rng.integers(0, n, 512))
    got = r.get_batch("dev", idx)
    want = ref_reader.assemble_batch_ref(fr, idx.numpy(), items, topk)
    assert list(got.keys()) == list(want.keys())
    for key in want:
        assert got[key].dtype == want[key].dtype and got[key].shape == want[key].shape, key
        assert torch.equal(got[key], want[key]), key


def test_data_parallel_ranks_partition_the_reference_stream(G):
    """world_size ranks x batch b: the union of the ranks' step-k batches is the reference loader's batch k of size
    world_size * b (same epoch order on every rank, strided slices)."""
    r = pair_reader(G)
    r.train_neg_sample()
    world, b = 2, 8
    torch.manual_seed(100)
    whole = list(r.batches("train", world * b, shuffle=True, drop_last=True))
    per_rank = []
    for rank in range(world):
        torch.manual_seed(100)
        per_rank.append(list(r.batches("train", b, shuffle=True, drop_last=True, rank=rank, world_size=world)))
    assert len(per_rank[0]) == len(per_rank[1]) == len(whole)
    for k, ref in enumerate(whole):
        for key in ref:
            merged = torch.stack([per_rank[rank][k][key] for rank in range(world)], dim=1).reshape(ref[key].shape)
            assert torch.equal(merged, ref[key]), (k, key)
    seen = torch.cat([x["index"] for rank in range(world) for x in r.batches("train", b, rank=rank, world_size=world)])
    assert seen.numel() == len(set(seen.tolist())) == r.size("train") - r.size("train") % world
    with pytest.raises(ValueError):
        next(r.batches("train", b, rank=2, world_size=2))


class _FakeFrame:
    def __init__(self, cols):
        self._c = cols
        self.columns = list(cols)

    def __getitem__(self, k):
        class V:
            values = self._c[k]
        return V


def test_adopting_a_reference_style_reader(G):
    class Fake:
        pass
    f = Fake()
    f.train_df, f.dev_df, f.test_df = (_FakeFrame(frame(G, f"pair/{s}")) for s in ("train", "dev", "test"))
    f.item_df = _FakeFrame(frame(G, "pair/items"))
    f.train_mode, f.split_mode = "pair_wise", "leave_k_out"
    f.dev_iid_topk_array, f.test_iid_topk_array = G["pair/dev_topk"], G["pair/test_topk"]
    f.user_pos_his_set_dict = sets(G, "pair/pos")
    f.rng = np.random.default_rng(int(G["pair/seed"]))
    f.random_seed = 2020
    f.train_iid_pair_array = None
    r = TensorDataReader.from_reference_reader(f)
    r.train_neg_sample()
    np.testing.assert_array_equal(r.train_iid_pair_array, G["pair/epoch0/pairs"])
    assert r.rng is f.rng  # one random stream: the adopted reader continues the reference reader's draws


def test_fit_is_transparent_to_the_loader_swap():
    """IModel.fit over a SplitDataset sees the batches DataLoader would have produced from the same rows."""
    from torch.utils.data import Dataset
    from pytorchrec_b200.metric import LogLoss
    from pytorchrec_b200.model import IModel

    rng = np.random.default_rng(0)
    n = 200
    fr = {"x": rng.random((n, 4)).astype(np.float32), "label": rng.integers(0, 2, n).astype(np.int32)}

    class Tiny(IModel):
        def _init_weights(self):
            self.lin = torch.nn.Linear(4, 1)

        def forward(self, data):
            return self.lin(data["x"]).flatten(), data["label"].float()

    class Rows(Dataset):
        def __len__(self):
            return n

        def __getitem__(self, i):
            return {"x": fr["x"][i], "label": fr["label"][i], "index": i}

    losses = []
    for ds in (Rows(), TensorDataReader(fr).train_dataset()):
        m = Tiny(random_seed=1)
        m.compile(torch.optim.SGD(m.get_parameters(), lr=0.1), torch.nn.BCEWithLogitsLoss(), [LogLoss()],
                  torch.device("cpu"))
        torch.manual_seed(9)
        h = m.fit(ds, batch_size=32, epochs=3, dev_dataset=ds)
        losses.append((h.history["loss"], h.history["logloss"] if "logloss" in h.history else None,
                       m.lin.weight.detach().clone()))
    assert losses[0][0] == losses[1][0]
    assert torch.equal(losses[0][2], losses[1][2])


@pytest.mark.gpu
def test_reader_resident_in_hbm_matches_reference_run(G):
    run_pair_case(G, torch.device("cuda:0"))
    r = pair_reader(G, torch.device("cuda:0"))
    r.train_neg_sample()
    b = next(iter(r.train_dataset().batches(16)))
    assert all(v.device.type == "cuda" for v in b.values())


@pytest.mark.gpu
def test_device_sampler_properties_gpu():
    check_device_sampler(torch.device("cuda:0"))
