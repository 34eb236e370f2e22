"""CPU-side checks: the C-ABI library loads and exports every declared symbol, host-side layout /
registry / lifecycle logic, seed parity of product models against the reference run, and that the
product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, state_from
from pytorchrec_b200 import _lib, ops
from pytorchrec_b200.feature_column import CategoricalColumnWithIdentity as Col
from pytorchrec_b200.feature_column import NumericColumn, NormalizationMode
from pytorchrec_b200.model import SVDPP, DeepFM, FunkSVD, get_model_type
from pytorchrec_b200.model.layer import EmbeddingTable, MultiTableEmbedding
from pytorchrec_b200.optim import SparseAdagrad, SparseAdam, SparseSGD, get_optimizer


def _declared_symbols():
    syms = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            text = open(os.path.join(inc, fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            syms.update(re.findall(r"\b(ptrec_[a-z0-9_]+)\s*\(", text))
    return syms


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert lib.ptrec_abi_version() == _lib.ABI_VERSION
    declared = _declared_symbols()
    assert len(declared) >= 15
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(raw, s), f"{s} declared in include/ but not exported"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.FeatureDesc) == 40 and ctypes.sizeof(_lib.OptimArgs) == 32
    assert _lib.FeatureDesc.id_base.offset == 24 and _lib.FeatureDesc.out_col.offset == 32


def test_argument_errors_come_back_as_codes_not_crashes():
    lib = _lib.load()
    rc = lib.ptrec_fm2_fwd(None, 0, 4, 2, 8, None, None)
    assert rc == -1 and b"fm2_fwd" in lib.ptrec_last_error()
    rc = lib.ptrec_index_prep(None, None, 4, 0, 0, None, None, None, 0, None)
    assert rc < 0
    # K6, fp16 x 2 format: null operands / missing scale words / missing workspace are refused before any launch
    assert lib.ptrec_tc_split2h(None, 8, 4, 8, None, 0, None, 8, None, 8, None, None, None, None, 0, None) < 0
    assert b"tc_split" in lib.ptrec_last_error()
    assert lib.ptrec_tc_gemm_split2h(None, None, 4, 8, None, None, 4, 8, 8, None, 0, None, 4, None, 1, None, 0, None) < 0
    assert lib.ptrec_tc_split2h_workspace_bytes(16384, 400) >= lib.ptrec_tc_split3_workspace_bytes(16384, 400) + 4096
    lib.ptrec_tc_set_bn(128)
    assert lib.ptrec_tc_get_bn() == 128
    lib.ptrec_tc_set_bn(7)          # anything but 128 selects the 256-wide tiles
    assert lib.ptrec_tc_get_bn() == 256


def test_k6_operand_mode_switch():
    assert ops.tc_mode() in ("fp16x2", "bf16x3")
    old = ops.tc_mode()
    ops.set_tc_mode("bf16x3")
    assert ops.tc_mode() == "bf16x3"
    with pytest.raises(ValueError):
        ops.set_tc_mode("tf32")
    ops.set_tc_mode(old)


def test_no_cpu_fallback():
    v = torch.randn(4, 3, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.fm2(v)
    t = EmbeddingTable(10, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        t(torch.tensor([1, 2, 3]))


def test_feature_layout_ordering_and_shared_tables():
    cols = [Col(10, "a"), Col(20, "b"), Col(10, "a_his"), Col(30, "c")]
    emb = MultiTableEmbedding(cols, 8, pooling={"a_his": "mean"}, mask={"a_his": "pad"}, share={"a_his": "a"})
    assert len(emb) == 3  # a_his shares a's table
    assert emb._table_of == [0, 1, 0, 2]
    lay = emb._layout_for((1, 1, 5, 1))
    tables = [lay.host[i].table for i in range(4)]
    assert tables == sorted(tables) == [0, 0, 1, 2]
    # internal order a, a_his, b, c ; output columns follow the user's order
    assert [lay.host[i].out_col for i in range(4)] == [0, 16, 8, 24]
    assert [lay.host[i].id_base for i in range(4)] == [0, 1, 6, 7]
    assert lay.total_bag_len == 8 and lay.slots(3) == 24
    assert lay.host[1].pooling == _lib.POOL_MEAN and lay.host[1].mask_mode == _lib.MASK_PAD


def test_seed_parity_with_reference_models(golden_mf):
    """Product models consume the RNG exactly like the reference's nn.Embedding models."""
    n_u, n_i, D, B, L, steps = (int(x) for x in golden_mf["dims"])
    uid, iid, iids, label = Col(n_u, "uid"), Col(n_i, "iid"), Col(n_i, "iids"), Col(2, "label")
    m = SVDPP(2020, uid, iid, iids, label, D)
    init = state_from(golden_mf, "svdpp_point_sgd/init")
    assert set(m.state_dict()) == set(init)
    for k, v in m.state_dict().items():
        assert torch.equal(v, init[k]), k
    f = FunkSVD(uid, iid, label, D, random_seed=2020)
    init = state_from(golden_mf, "funksvd_point_sgd/init")
    for k, v in f.state_dict().items():
        assert torch.equal(v, init[k]), k
    # a reference state_dict loads straight into the product model
    m.load_state_dict(state_from(golden_mf, "svdpp_point_sgd/final"))


def test_ncf_seed_parity_and_checkpoint_keys_with_reference_run(golden_ncf):
    """The product NCF (fused tables + K6-capable tower) has the reference NCF's ``state_dict`` keys and, under the same
    seed, its initial weights bit for bit — a reference checkpoint loads unchanged."""
    from conftest import state_from
    from pytorchrec_b200.model import NCF, get_model_type
    n_u, n_i, D, B, steps, *layers = (int(x) for x in golden_ncf["dims"])
    model = NCF(2020, Col(n_u, "uid"), Col(n_i, "iid"), Col(2, "label"), D, layers, 0.0)
    init = state_from(golden_ncf, "ncf_n2_sgd/init")
    assert list(model.state_dict().keys()) == list(init.keys())
    for k, v in model.state_dict().items():
        assert torch.equal(v, init[k]), k
    model.load_state_dict(state_from(golden_ncf, "ncf_n2_sgd/final"))
    assert get_model_type("ncf") is NCF
    groups = model.get_parameters()
    assert sum(p.numel() for p in groups[1]["params"]) == sum(layers)   # the tower's biases, undecayed


def test_deepfm_seed_parity_with_oracle_twin(golden_ctr):
    F, nd, D, B = (int(x) for x in golden_ctr["dims"])
    rows = [int(r) for r in golden_ctr["rows"]]
    scols = [Col(rows[f], f"C{f}") for f in range(F)]
    dcols = [NumericColumn(f"I{j}", 0.0, 1.0, 0.5, 0.25) for j in range(nd)]
    m = DeepFM(scols, dcols, Col(2, "label"), D, [16, 8], random_seed=2020)
    init = state_from(golden_ctr, "deepfm_adagrad/init")
    assert set(m.state_dict()) == set(init)
    for k, v in m.state_dict().items():
        assert torch.equal(v, init[k]), k


def test_param_groups_and_registries():
    cols = [Col(10, "a"), Col(20, "b")]
    m = DeepFM(cols, [NumericColumn("x", 0, 1, .5, .2)], Col(2, "label"), 4, [8], random_seed=1)
    groups = m.get_parameters()
    assert groups[1]["weight_decay"] == 0.0
    names = {id(p): n for n, p in m.named_parameters()}
    assert all("bias" in names[id(p)] for p in groups[1]["params"])
    assert any("global_bias" == names[id(p)] for p in groups[1]["params"])
    for name in ["sgd", "adam", "adamw", "adagrad", "sparse_sgd", "sparse_adagrad", "sparse_rowwise_adagrad", "sparse_adam"]:
        opt = get_optimizer(name)(params=m.get_parameters(), lr=0.01)
        assert isinstance(opt, torch.optim.Optimizer)
    with pytest.raises(ValueError):
        get_optimizer("nope")
    assert get_model_type("deepfm") is DeepFM
    opt = SparseAdagrad(m.get_parameters(), lr=0.1)
    assert len(opt._fused) == 4  # 2 embedding + 2 first-order tables
    fused_ids = {id(p) for p in opt._fused}
    opt._ensure_dense()
    for g in opt._dense.param_groups:
        assert all(id(p) not in fused_ids for p in g["params"])


def test_compile_validates_like_the_reference():
    from pytorchrec_b200.metric import LogLoss
    m = FunkSVD(Col(5, "uid"), Col(5, "iid"), Col(2, "label"), 4, random_seed=1)
    opt = SparseSGD(m.get_parameters(), lr=0.1)
    with pytest.raises(ValueError):
        m.compile("sgd", torch.nn.MSELoss(), [LogLoss()], torch.device("cpu"))
    with pytest.raises(ValueError):
        m.compile(opt, "mse", [LogLoss()], torch.device("cpu"))
    with pytest.raises(ValueError):
        m.compile(opt, torch.nn.MSELoss(), [], "cpu")
    with pytest.raises(RuntimeError):
        m.fit(None, 1, 1)
    m.compile(opt, torch.nn.MSELoss(), [LogLoss()], torch.device("cpu"))


def test_sparse_optimizer_dense_companion_matches_torch():
    torch.manual_seed(0)
    lin_a, lin_b = torch.nn.Linear(4, 3), torch.nn.Linear(4, 3)
    lin_b.load_state_dict(lin_a.state_dict())
    oa = SparseAdam([{"params": list(lin_a.parameters())}], lr=0.01)
    ob = torch.optim.Adam(lin_b.parameters(), lr=0.01)
    x = torch.randn(5, 4)
    for _ in range(3):
        for lin, o in ((lin_a, oa), (lin_b, ob)):
            o.zero_grad()
            lin(x).pow(2).sum().backward()
            o.step()
    for pa, pb in zip(lin_a.parameters(), lin_b.parameters()):
        assert torch.equal(pa, pb)
    sd = oa.state_dict()
    assert sd["step"] == 3 and sd["dense"] is not None


def test_numeric_column_normalisation():
    c = NumericColumn("x", 1.0, 3.0, 2.0, 0.5)
    b = {"x": torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)}
    assert c.get_feature_data(b).dtype == torch.float32
    np.testing.assert_allclose(c.get_feature_data(b, NormalizationMode.MAX_MIN).numpy(), [0, .5, 1])
    np.testing.assert_allclose(c.get_feature_data(b, NormalizationMode.Z_SCORE).numpy(), [-2, 0, 2])


def test_metrics_and_losses():
    from pytorchrec_b200.loss import BPRLoss, Top1Loss, get_loss
    from pytorchrec_b200.metric import Hit, MetricList, NDCG
    pred = np.array([[0.9, 0.1, 0.2], [0.1, 0.9, 0.3], [0.5, 0.6, 0.1]])
    out = MetricList([Hit(3, 1), NDCG(3, 2)])(pred, None)
    np.testing.assert_allclose(out["hit@1"], 1 / 3)
    np.testing.assert_allclose(out["ndcg@2"], (1.0 + 0.0 + 1 / np.log2(3)) / 3)
    x = torch.tensor([[2.0, 1.0], [0.0, 3.0]])
    np.testing.assert_allclose(BPRLoss()(x, None).item(), torch.nn.functional.softplus(torch.tensor([-1.0, 3.0])).mean().item())
    assert Top1Loss(reduction="none")(x, None).shape == (2,)
    assert get_loss("bce") is torch.nn.BCEWithLogitsLoss


def test_device_rank_equals_host_rank():
    """N3: the torch rank (usable on the GPU) agrees with the numpy argsort rank, ties included."""
    from pytorchrec_b200.metric import Hit, MetricList, NDCG, get_pos_rank, get_pos_rank_torch
    g = torch.Generator().manual_seed(0)
    pred = torch.randint(0, 6, (400, 100), generator=g).float()  # many ties
    np.testing.assert_array_equal(get_pos_rank_torch(pred, 100).numpy(), get_pos_rank(pred.numpy(), 100))
    ml = MetricList([Hit(100, 10), NDCG(100, 10)])
    a, b = ml(pred, None), ml(pred.numpy(), None)
    assert a.keys() == b.keys() and all(abs(a[k] - b[k]) < 1e-12 for k in a)


def test_fit_drives_the_reference_callback_lifecycle_and_save_weights_stores_weights_only(tmp_path):
    """IModel.fit calls set_model / on_train_begin / on_epoch_begin / on_epoch_end / on_train_end like the reference's
    CallbackList (IModel.py:160-208) — an EarlyStopping-style callback can stop training and restore the best weights in
    on_train_end; save_weights writes contiguous weight tensors, not the storage a strided table view lives in."""
    from pytorchrec_b200.metric import LogLoss
    from pytorchrec_b200.model.IModel import IModel

    class Tiny(IModel):
        def _init_weights(self):
            self.lin = torch.nn.Linear(3, 1)

        def forward(self, data):
            return self.lin(data["x"]).squeeze(-1), data["y"]

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return 8

        def __getitem__(self, i):
            g = torch.Generator().manual_seed(i)
            return {"x": torch.randn(3, generator=g), "y": torch.tensor(float(i % 2))}

    calls = []

    class Stopper:  # reference-style callback: keeps the model it was given, stops after epoch 1, restores at the end
        def set_model(self, model):
            self.model = model

        def set_params(self, params):
            calls.append(("params", params["epochs"], params["batches"]))

        def on_train_begin(self, logs=None):
            calls.append("train_begin")

        def on_epoch_begin(self, epoch, logs=None):
            calls.append(("epoch_begin", epoch))

        def on_epoch_end(self, epoch, logs=None):
            calls.append(("epoch_end", epoch, "loss" in logs))
            if epoch == 0:
                self.model.save_best_weights()
            if epoch == 1:
                self.model.stop_training = True

        def on_train_end(self, logs=None):
            calls.append("train_end")
            self.model.load_best_weights()

    m = Tiny(random_seed=3)
    m.compile(torch.optim.SGD(m.parameters(), lr=0.1), torch.nn.MSELoss(), [LogLoss()], torch.device("cpu"))
    m.fit(DS(), batch_size=4, epochs=5, callbacks=[Stopper()])
    assert calls == [("params", 5, 2), "train_begin", ("epoch_begin", 0), ("epoch_end", 0, True), ("epoch_begin", 1),
                     ("epoch_end", 1, True), "train_end"]
    for k, v in m.state_dict().items():
        assert torch.equal(v, m.best_state_dict[k])

    # a table re-housed as a strided view of a [rows, 2*D] weight | state buffer (what SparseAdagrad does on the GPU)
    emb = FunkSVD(Col(50, "uid"), Col(40, "iid"), Col(2, "label"), 8, random_seed=1)
    w = emb.uid_embeddings.weight if hasattr(emb, "uid_embeddings") else next(emb.parameters())
    buf = torch.zeros(w.shape[0], 4 * w.shape[1])
    buf[:, :w.shape[1]] = w.data
    w.data = buf[:, :w.shape[1]]
    path = str(tmp_path / "w.pt")
    emb.save_weights(path)
    loaded = torch.load(path, weights_only=False)
    for k, v in emb.state_dict().items():
        assert torch.equal(loaded[k], v) and loaded[k].is_contiguous()
        assert loaded[k].untyped_storage().nbytes() == v.numel() * v.element_size(), k


def test_side_stream_bookkeeping_is_inert_off_the_gpu_and_outside_a_train_step():
    """Host logic of the backward's side streams (model/layer/dense.py, embedding.py): registrations are per device,
    deduplicated, dropped by ``reset_join_streams`` (a step that never reached its join) and consumed by
    ``join_aux_streams``; ``side_reductions`` does nothing — and does not touch the library's reduce stream — unless an
    ``IModel`` train step on a CUDA device is running."""
    from pytorchrec_b200.model.layer import dense, embedding

    class FakeStream:
        pass

    dev = torch.device("cuda", 0)
    a, b = FakeStream(), FakeStream()
    embedding.reset_join_streams(dev)
    embedding.register_join_stream(dev, a)
    embedding.register_join_stream(dev, a)
    embedding.register_join_stream("cuda:0", b)
    assert embedding._JOIN_STREAMS[dev] == [a, b]
    embedding.reset_join_streams(dev)
    assert dev not in embedding._JOIN_STREAMS
    embedding.join_aux_streams(None)          # CPU model: nothing to join

    assert dense._DEFER_JOIN == [False]
    with dense.side_reductions(None) as sr:   # CPU tensors
        assert not sr.on
        sr.adopt(torch.zeros(1), None)
    with dense.side_reductions(dev) as sr:    # CUDA device but no train step around: stays on the producer's stream
        assert not sr.on
    dense._DEFER_JOIN[0] = True
    try:
        with dense.side_reductions(None) as sr:
            assert not sr.on
    finally:
        dense._DEFER_JOIN[0] = False
