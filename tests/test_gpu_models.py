"""Model-level parity on the GPU: the product models (fused CUDA embedding path + fused sparse
optimizers) against (a) golden vectors from the unmodified reference run and (b) the CPU oracle
twins, through the reference's own five-line ``train_step``."""
import copy
import os

import numpy as np
import pytest
import torch

from conftest import adagrad_sums, assert_as_exact_as_the_oracle, batch_from, state_from
from oracle import ref_models
from pytorchrec_b200.feature_column import CategoricalColumnWithIdentity as Col
from pytorchrec_b200.feature_column import NumericColumn
from pytorchrec_b200.loss import BPRLoss
from pytorchrec_b200.metric import LogLoss
from pytorchrec_b200.model import SVDPP, DeepFM, FM, FunkSVD
from pytorchrec_b200.optim import SparseAdagrad, SparseAdam, SparseRowWiseAdagrad, SparseSGD

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _mf_cols(g):
    n_u, n_i, D, B, L, steps = (int(x) for x in g["dims"])
    return dict(uid=Col(n_u, "uid"), iid=Col(n_i, "iid"), iids=Col(n_i, "iids"), label=Col(2, "label")), D, steps


@pytest.mark.parametrize("case", ["svdpp_point_sgd", "svdpp_pair_sgd", "funksvd_point_sgd"])
def test_reference_models_on_fused_tables_match_reference_run(golden_mf, case):
    """SVD++ / FunkSVD on EmbeddingTable + SparseSGD reproduce the reference's own train_step results
    (dense SGD == sparse SGD on touched rows; untouched rows do not move in either)."""
    cols, D, steps = _mf_cols(golden_mf)
    if case.startswith("svdpp"):
        model = SVDPP(2020, cols["uid"], cols["iid"], cols["iids"], cols["label"], D)
    else:
        model = FunkSVD(cols["uid"], cols["iid"], cols["label"], D, random_seed=2020)
    init = state_from(golden_mf, f"{case}/init")
    for k, v in model.state_dict().items():
        assert torch.equal(v, init[k]), k
    opt = SparseSGD(params=model.get_parameters(), lr=0.5)
    loss = BPRLoss() if "_pair_" in case else torch.nn.MSELoss()
    model.compile(opt, loss, [LogLoss()], DEV)
    for s in range(steps):
        batch = batch_from(golden_mf, f"{case}/batch{s}")
        pred, target = model.test_step(batch)
        np.testing.assert_allclose(pred.detach().cpu().numpy(), golden_mf[f"{case}/pred{s}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_array_equal(target.cpu().numpy(), golden_mf[f"{case}/target{s}"])
        logs = model.train_step(batch)
        np.testing.assert_allclose(logs["loss"].item(), golden_mf[f"{case}/loss{s}"], rtol=1e-5)
    final = state_from(golden_mf, f"{case}/final")
    for k, v in model.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), final[k].numpy(), rtol=1e-5, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("case", ["ncf_n2_sgd", "ncf_n5_sgd"])
def test_reference_ncf_on_fused_tables_and_k6_tower_matches_reference_run(golden_ncf, case):
    """NCF on EmbeddingTable + SparseSGD with its MLP on the tensor-core Linear path reproduces the reference's own
    train_step results: the Dense / MLP row of the scope table (a8) checked against reference-run numbers."""
    from pytorchrec_b200.model import NCF
    n_u, n_i, D, B, steps, *layers = (int(x) for x in golden_ncf["dims"])
    model = NCF(2020, Col(n_u, "uid"), Col(n_i, "iid"), Col(2, "label"), D, layers, 0.0)
    init = state_from(golden_ncf, f"{case}/init")
    for k, v in model.state_dict().items():
        assert torch.equal(v, init[k]), k
    loss = BPRLoss() if "_n2_" in case else torch.nn.BCEWithLogitsLoss()
    model.compile(SparseSGD(params=model.get_parameters(), lr=0.5), loss, [LogLoss()], DEV)
    for s in range(steps):
        batch = batch_from(golden_ncf, f"{case}/batch{s}")
        pred, target = model.test_step(batch)
        np.testing.assert_allclose(pred.detach().cpu().numpy(), golden_ncf[f"{case}/pred{s}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_array_equal(target.cpu().numpy(), golden_ncf[f"{case}/target{s}"])
        np.testing.assert_allclose(model.train_step(batch)["loss"].item(), golden_ncf[f"{case}/loss{s}"], rtol=1e-5)
    final = state_from(golden_ncf, f"{case}/final")
    for k, v in model.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), final[k].numpy(), rtol=1e-5, atol=1e-7, err_msg=k)


@pytest.fixture(autouse=True)
def _small_layers_on_k6():
    """The test models are tiny; push their Dense layers through the tensor-core path (K6) anyway — the size
    heuristic would leave them on cuBLAS (covered by test_dense_layer_size_heuristic)."""
    from pytorchrec_b200.model.layer import dense
    old, dense.TC_MIN_MACS = dense.TC_MIN_MACS, 0
    yield
    dense.TC_MIN_MACS = old


def _ctr_setup(F=6, nd=3, D=16, rows=None, seed=11):
    rows = rows or [50 + 13 * f for f in range(F)]
    scols = [Col(rows[f], f"C{f}") for f in range(F)]
    dcols = [NumericColumn(f"I{j}", 0.0, 1.0, 0.5, 0.25) for j in range(nd)]
    return scols, dcols, Col(2, "label"), rows


def _ctr_batch(rows, nd, B, seed, zipf=False):
    rng = np.random.default_rng(seed)
    b = {}
    for f, r in enumerate(rows):
        x = (rng.zipf(1.2, size=B) % r) if zipf else rng.integers(0, r, size=B)
        b[f"C{f}"] = torch.from_numpy(x.astype(np.int32))
    for j in range(nd):
        b[f"I{j}"] = torch.from_numpy(rng.random(B).astype(np.float32))
    b["label"] = torch.from_numpy(rng.integers(0, 2, size=B).astype(np.int32))
    return b


@pytest.mark.parametrize("model_name", ["fm", "deepfm"])
@pytest.mark.parametrize("opt_name", ["sgd", "adagrad"])
@pytest.mark.parametrize("zipf", [False, True])
def test_ctr_models_match_oracle_twins(model_name, opt_name, zipf):
    """Same seed -> bit-identical init; N train steps: logits, loss and every parameter within 1e-5 rel."""
    scols, dcols, lab, rows = _ctr_setup()
    D, B, nd = 16, 512, len(dcols)
    if model_name == "fm":
        prod, ref = FM(scols, dcols, lab, D, random_seed=7), ref_models.FMRef(7, scols, dcols, lab, D)
    else:
        prod = DeepFM(scols, dcols, lab, D, [32, 16], random_seed=7)
        ref = ref_models.DeepFMRef(7, scols, dcols, lab, D, [32, 16])
    for (k, v), (k2, v2) in zip(prod.state_dict().items(), ref.state_dict().items()):
        assert k == k2 and torch.equal(v, v2), k
    if opt_name == "sgd":
        popt, ropt = SparseSGD(prod.get_parameters(), lr=0.3), torch.optim.SGD(ref.get_parameters(), lr=0.3)
    else:
        popt, ropt = SparseAdagrad(prod.get_parameters(), lr=0.05), torch.optim.Adagrad(ref.get_parameters(), lr=0.05)
    ref64 = copy.deepcopy(ref).fp64()
    ref64.compile(type(ropt)(ref64.get_parameters(), lr=ropt.defaults["lr"]), torch.nn.BCEWithLogitsLoss())
    prod.compile(popt, torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
    ref.compile(ropt, torch.nn.BCEWithLogitsLoss())
    for s in range(4):
        batch = _ctr_batch(rows, nd, B, seed=100 + s, zipf=zipf)
        ref64.train_step(batch)
        pl, _ = prod.test_step(batch)
        ref.eval()
        with torch.no_grad():
            rl, _ = ref({k: v for k, v in batch.items()})
        # 1e-5 relative to the logit scale: a logit is a cancelling sum of ~100 terms, and Adagrad's
        # g / sqrt(sum g^2) normalisation turns reduction-order noise in g into O(lr * 1e-6) weight noise
        scale = max(1.0, rl.abs().max().item())
        np.testing.assert_allclose(pl.detach().cpu().numpy(), rl.numpy(), rtol=1e-5, atol=1e-5 * scale)
        lp = prod.train_step(batch)["loss"].item()
        lr_ = ref.train_step(batch)["loss"].item()
        np.testing.assert_allclose(lp, lr_, rtol=1e-5)
    sd64 = ref64.state_dict()
    sums = adagrad_sums(ref64) if opt_name == "adagrad" else {}
    for (k, v), (_, v2) in zip(prod.state_dict().items(), ref.state_dict().items()):
        # |dw| per step is O(lr): tolerance 1e-5 * (|w| + steps * lr).  Adagrad's g / (sqrt(sum g^2) + 1e-10) is
        # discontinuous where a row's duplicate gradients cancel to ~0; there the fp64 twin referees (conftest): the
        # error must stay within what a 1e-5 gradient error can cause through Adagrad (no fraction of elements exempted)
        assert_as_exact_as_the_oracle(k, v.cpu().numpy(), v2.numpy(), sd64[k].numpy(), rtol=1e-5, atol=1e-5 * 4 * 0.3,
                                      adagrad=(sums[k].numpy(), 0.05, 4) if k in sums else None)
    prod.embeddings.check_index_errors()


def test_golden_ctr_regression_on_gpu(golden_ctr):
    F, nd, D, B = (int(x) for x in golden_ctr["dims"])
    rows = [int(r) for r in golden_ctr["rows"]]
    scols = [Col(rows[f], f"C{f}") for f in range(F)]
    dcols = [NumericColumn(f"I{j}", 0.0, 1.0, 0.5, 0.25) for j in range(nd)]
    m = DeepFM(scols, dcols, Col(2, "label"), D, [16, 8], random_seed=2020)
    m.compile(SparseAdagrad(m.get_parameters(), lr=0.1), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
    tag = "deepfm_adagrad"
    for s in range(3):
        b = batch_from(golden_ctr, f"{tag}/batch{s}")
        pred, _ = m.test_step(b)
        np.testing.assert_allclose(pred.detach().cpu().numpy(), golden_ctr[f"{tag}/pred{s}"], rtol=1e-5, atol=1e-6)
        m.train_step(b)
    final = state_from(golden_ctr, f"{tag}/final")
    for k, v in m.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), final[k].numpy(), rtol=2e-5, atol=2e-6, err_msg=k)


def test_lazy_adam_and_rowwise_run_and_state_dict_roundtrip(tmp_path):
    scols, dcols, lab, rows = _ctr_setup(F=4)
    for Opt in (SparseAdam, SparseRowWiseAdagrad):
        m = DeepFM(scols, dcols, lab, 8, [16], random_seed=3)
        m.compile(Opt(m.get_parameters(), lr=0.01), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
        losses = [m.train_step(_ctr_batch(rows, len(dcols), 256, seed=5))["loss"].item() for _ in range(20)]
        assert losses[-1] < losses[0], "training on a fixed batch must reduce the loss"
        path = str(tmp_path / "w.pt")
        m.save_weights(path)
        m2 = DeepFM(scols, dcols, lab, 8, [16], random_seed=4)
        m2.load_weights(path, DEV)
        for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
            assert torch.equal(a, b), k
        m.save_best_weights()
        m.load_best_weights()
        sd = m.compiled_optimizers.state_dict()
        assert sd["step"] == 20


@pytest.mark.parametrize("sparse", [False, True])
def test_stock_optimizer_mode_gives_nn_embedding_style_grads(sparse):
    """Without a fused optimizer the tables still train: weight.grad equals the dense reference gradient — a dense
    tensor by default (nn.Embedding's sparse=False, the reference's models), a coalesced COO tensor with sparse=True."""
    scols, dcols, lab, rows = _ctr_setup(F=3, nd=0)
    m = FM(scols, [], lab, 8, random_seed=5).to(DEV)
    for t in list(m.embeddings) + list(m.first_order):
        t.sparse = sparse
    ref = ref_models.FMRef(5, scols, [], lab, 8)
    batch = _ctr_batch(rows, 0, 128, seed=1, zipf=True)
    dbatch = {k: v.to(DEV) for k, v in batch.items()}
    pl, t = m(dbatch)
    torch.nn.functional.binary_cross_entropy_with_logits(pl, t).backward()
    rl, rt = ref(batch)
    torch.nn.functional.binary_cross_entropy_with_logits(rl, rt).backward()
    for f in range(3):
        g = m.embeddings[f].weight.grad
        assert g.is_sparse == sparse
        g = g.coalesce().to_dense() if sparse else g
        np.testing.assert_allclose(g.cpu().numpy(), ref.embeddings[f].weight.grad.numpy(), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("name,kw", [("sgd", dict(lr=0.1)), ("sgd", dict(lr=0.1, weight_decay=1e-3)),
                                     ("adam", dict(lr=0.01)), ("adamw", dict(lr=0.01, weight_decay=1e-2)),
                                     ("adagrad", dict(lr=0.05)), ("sparse_sgd", dict(lr=0.1)),
                                     ("sparse_adagrad", dict(lr=0.05)), ("sparse_rowwise_adagrad", dict(lr=0.05)),
                                     ("sparse_adam", dict(lr=0.01))])
def test_every_registry_optimizer_trains_a_model(name, kw):
    """get_optimizer(name)(params=model.get_parameters(), **kw) (RepeatTask.py:96) drives train_step for every entry
    of the registry; the reference's own entries (sgd / adam / adamw: dense nn.Embedding gradients) reproduce the
    oracle twin trained with the same stock optimizer on the CPU."""
    from pytorchrec_b200.optim import get_optimizer
    scols, dcols, lab, rows = _ctr_setup(F=3, nd=2)
    m = FM(scols, dcols, lab, 8, random_seed=5)
    m.compile(get_optimizer(name)(params=m.get_parameters(), **kw), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
    ref = None
    if not name.startswith("sparse_") :
        ref = ref_models.FMRef(5, scols, dcols, lab, 8)
        ref.compile(get_optimizer(name)(params=ref.get_parameters(), **kw), torch.nn.BCEWithLogitsLoss())
    before = {k: v.clone() for k, v in m.state_dict().items()}
    for s in range(3):
        batch = _ctr_batch(rows, 2, 64, seed=20 + s, zipf=True)
        loss = m.train_step(batch)["loss"].item()
        assert np.isfinite(loss)
        if ref is not None:
            np.testing.assert_allclose(loss, ref.train_step(batch)["loss"].item(), rtol=1e-5)
    after = m.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before if k.startswith("embeddings"))
    if ref is not None:
        for (k, a), (_, b) in zip(after.items(), ref.state_dict().items()):
            np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=2e-4, atol=2e-6, err_msg=k)


def test_out_of_range_id_is_reported():
    scols, dcols, lab, rows = _ctr_setup(F=2, nd=0)
    m = FM(scols, [], lab, 8, random_seed=5).to(DEV)
    batch = {k: v.to(DEV) for k, v in _ctr_batch(rows, 0, 16, seed=1).items()}
    batch["C0"][3] = rows[0] + 5
    with torch.no_grad():
        m(batch)
    with pytest.raises(IndexError):
        m.embeddings.check_index_errors()


def test_cuda_graph_train_step_equals_eager():
    """The whole-step CUDA graph replays exactly the eager step (same kernels, same order)."""
    scols, dcols, lab, rows = _ctr_setup()
    models = []
    for graphed in (False, True):
        m = DeepFM(scols, dcols, lab, 16, [32, 16], random_seed=9)
        m.compile(SparseAdagrad(m.get_parameters(), lr=0.05), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
        if graphed:
            m.enable_cuda_graph(True, warmup=2)
        losses = [m.train_step(_ctr_batch(rows, len(dcols), 256, seed=40 + s, zipf=True))["loss"].item() for s in range(6)]
        models.append((m, losses))
    (me, le), (mg, lg) = models
    np.testing.assert_allclose(lg, le, rtol=1e-6)
    for (k, a), (_, b) in zip(me.state_dict().items(), mg.state_dict().items()):
        assert torch.equal(a, b), k
    assert mg.compiled_optimizers._step_count_fused == 6
    with pytest.raises(RuntimeError):
        m = DeepFM(scols, dcols, lab, 16, [32, 16], random_seed=9)
        m.compile(SparseAdam(m.get_parameters(), lr=0.05), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
        m.enable_cuda_graph(True, warmup=0)
        m.train_step(_ctr_batch(rows, len(dcols), 256, seed=1))


@pytest.mark.parametrize("model_name", ["deepfm", "dcn"])
@pytest.mark.parametrize("graphed", [False, True], ids=["eager", "graph"])
def test_side_stream_gradients_train_identically(monkeypatch, model_name, graphed):
    """The fused tower's weight-gradient GEMMs (PTREC_WGRAD_STREAM) and the gradient-finalising reductions
    (PTREC_REDUCE_STREAM) run on a side stream beside the input-gradient chain inside an IModel train step: every
    kernel is deterministic and the join precedes the optimizer step, so many steps with the side streams on equal the
    same steps with everything on the main stream bit for bit — a missing dependency shows up as a difference."""
    from pytorchrec_b200.model import DCN
    from pytorchrec_b200.model.layer import dense
    monkeypatch.setattr(dense, "TC_MIN_MACS", 0)       # small layers on the fused tower too
    scols, dcols, lab, rows = _ctr_setup()

    def run(side: str):
        monkeypatch.setenv("PTREC_WGRAD_STREAM", side)
        monkeypatch.setenv("PTREC_REDUCE_STREAM", side)
        if model_name == "deepfm":
            m = DeepFM(scols, dcols, lab, 16, [64, 48, 32], random_seed=9)
        else:
            m = DCN(scols, dcols, lab, 16, 2, [64, 32], random_seed=9)
        m.compile(SparseAdagrad(m.get_parameters(), lr=0.05), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
        if graphed:
            m.enable_cuda_graph(True, warmup=3)
        losses = [m.train_step(_ctr_batch(rows, len(dcols), 2048, seed=70 + s, zipf=True))["loss"].item() for s in range(25)]
        return losses, {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}

    l1, sd1 = run("1")
    l0, sd0 = run("0")
    assert l1 == l0
    for k in sd1:
        assert torch.equal(sd1[k], sd0[k]), k


@pytest.mark.parametrize("peer", ["push", "1", "0"], ids=["push", "pull_peer_memory", "all_to_all"])
def test_sharded_deepfm_matches_unsharded_on_two_gpus(peer):
    """Row-wise sharded tables (owners push rows over NVLink; requesters pull from peer memory; NCCL all-to-all) + the
    dense-gradient mean == single-GPU model on the concatenated batch; low-cardinality columns, a forced list overflow
    (must be fatal) and the sharded checkpoint incl. optimizer state ride along (tests/dist_sharded_worker.py)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", {"push": "29532", "1": "29533", "0": "29534"}[peer],
           os.path.join(root, "tests", "dist_sharded_worker.py")]
    env = dict(os.environ, PTREC_EXCHANGE="push") if peer == "push" else dict(os.environ, PTREC_PEER_GATHER=peer)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "DIST_SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_a2a_pack_kernels_match_cpu_plan():
    from oracle import ref_sharding
    from pytorchrec_b200 import ops
    for (F, B, G) in ((3, 100, 2), (26, 4099, 8), (5, 2048, 4)):
        rows = 1000
        ids = torch.stack([torch.randint(-1, rows, (B,), generator=torch.Generator().manual_seed(f)) for f in range(F)])
        for C in (B, max(16, B // G)):  # ample capacity, and one that overflows
            ovf = torch.zeros(1, dtype=torch.int32, device=DEV)
            send_ids, ret_pos = ops.a2a_pack_by_owner(ids.to(DEV), F, B, G, C, ovf)
            rs, rp, longest = ref_sharding.pack_by_owner_ref(ids, G, C)
            assert torch.equal(send_ids.cpu(), rs) and torch.equal(ret_pos.cpu(), rp)
            assert ovf.item() == (longest if longest > C else 0)
        D = 16
        src = torch.randn(B, F * D)
        dst = torch.zeros(G * F * C, D, device=DEV)
        ops.a2a_scatter_rows(src.to(DEV), ret_pos, B, F, D, 0.5, dst)
        want = torch.zeros(G * F * C, D)
        for f in range(F):
            ok = rp[f] >= 0
            want[rp[f][ok].long()] = 0.5 * src.view(B, F, D)[ok, f]
        assert torch.equal(dst.cpu(), want)


def test_peer_dispatch_kernels_equal_all_to_all_plan():
    """The peer-memory pack / scatter (stores straight into the owners' buffers) leave exactly what the all-to-all of
    the CPU plan would deliver; G virtual ranks on one device, bit-exact."""
    from oracle import ref_sharding
    from pytorchrec_b200 import ops
    for (F, B, G, D) in ((3, 100, 2, 16), (26, 1031, 8, 16), (5, 2048, 4, 1)):
        rows, S = 997, 20
        C = max(16, B // G + 40)
        own_ids = [torch.full((F * G * C,), -1, dtype=torch.int64, device=DEV) for _ in range(G)]
        own_g = [torch.zeros(G * F * C, S, device=DEV) for _ in range(G)]
        peer_ids = torch.tensor([t.data_ptr() for t in own_ids], dtype=torch.int64).to(DEV)
        peer_g = torch.tensor([t.data_ptr() for t in own_g], dtype=torch.int64).to(DEV)
        want_ids = torch.full((G, F, G, C), -1, dtype=torch.int64)      # [owner][f, src, c]
        want_g = torch.zeros(G, G * F * C, S)                            # [owner][(src*F + f)*C + c]
        for r in range(G):
            gen = torch.Generator().manual_seed(100 + r)
            ids = torch.stack([torch.randint(-1, rows, (B,), generator=gen) for _ in range(F)])
            src = torch.randn(B, F * D, generator=gen)
            ovf = torch.zeros(1, dtype=torch.int32, device=DEV)
            ret_pos = ops.a2a_pack_by_owner_peer(ids.to(DEV), F, B, G, C, r, peer_ids, ovf)
            ops.a2a_scatter_rows_peer(src.to(DEV), ret_pos, B, F, D, 0.25, peer_g, S, 4, C, G, r)
            rs, rp, longest = ref_sharding.pack_by_owner_ref(ids, G, C)
            assert torch.equal(ret_pos.cpu(), rp) and ovf.item() == (longest if longest > C else 0)
            want_ids[:, :, r, :] = rs                                   # all-to-all: owner o receives rs[o] from r
            for f in range(F):
                ok = rp[f] >= 0
                pos = rp[f][ok].long()
                o, rem = pos // (F * C), pos % (F * C)
                want_g[o, r * F * C + rem, 4:4 + D] = 0.25 * src.view(B, F, D)[ok, f]
        for o in range(G):
            assert torch.equal(own_ids[o].cpu().view(F, G, C), want_ids[o])
            assert torch.equal(own_g[o].cpu(), want_g[o])
        # every width in one launch (DeepFM: D and 1) leaves the same bytes as one launch per width
        multi = [torch.zeros(G * F * C, S, device=DEV) for _ in range(G)]
        single = [torch.zeros(G * F * C, S, device=DEV) for _ in range(G)]
        ordered = [torch.zeros(G * F * C, S, device=DEV) for _ in range(G)]
        pm = torch.tensor([t.data_ptr() for t in multi], dtype=torch.int64).to(DEV)
        ps = torch.tensor([t.data_ptr() for t in single], dtype=torch.int64).to(DEV)
        po = torch.tensor([t.data_ptr() for t in ordered], dtype=torch.int64).to(DEV)
        for r in range(G):
            gen = torch.Generator().manual_seed(200 + r)
            ids = torch.stack([torch.randint(-1, rows, (B,), generator=gen) for _ in range(F)]).to(DEV)
            g0, g1 = torch.randn(B, F * D, generator=gen).to(DEV), torch.randn(B, F, generator=gen).to(DEV)
            scratch = torch.tensor([t.data_ptr() for t in own_ids], dtype=torch.int64).to(DEV)
            slot_b = torch.full((G * F * C,), 12345, dtype=torch.int32, device=DEV)
            ret_pos = ops.a2a_pack_by_owner_peer(ids, F, B, G, C, r, scratch, torch.zeros(1, dtype=torch.int32, device=DEV),
                                                 slot_b=slot_b)
            # slot_b is the inverse of ret_pos, -1 behind the last lookup of every list
            rp, sb = ret_pos.cpu(), slot_b.cpu()
            want_sb = torch.full_like(sb, -1)
            for f in range(F):
                ok = rp[f] >= 0
                want_sb[rp[f][ok].long()] = torch.arange(B, dtype=torch.int32)[ok]
            assert torch.equal(sb, want_sb)
            ops.a2a_scatter_rows_peer_multi([g0, g1], [D, 1], [0, 16], ret_pos, B, F, 0.5, pm, S, C, G, r)
            ops.a2a_scatter_rows_peer(g0, ret_pos, B, F, D, 0.5, ps, S, 0, C, G, r)
            ops.a2a_scatter_rows_peer(g1, ret_pos, B, F, 1, 0.5, ps, S, 16, C, G, r)
            if D == 16:  # the ordered kernel wants the widths packed at 16-byte boundaries in slot order
                ops.a2a_scatter_rows_peer_ordered([g0, g1], [D, 1], [0, 16], slot_b, F, 0.5, po, S, C, G, r)
        for o in range(G):
            assert torch.equal(multi[o], single[o])
            if D == 16:
                assert torch.equal(ordered[o], single[o])  # destination order, whole slots (pad columns stay zero)


@pytest.mark.parametrize("F,B,G,dims", [(3, 100, 2, [16, 1]), (26, 1031, 8, [16, 1]), (5, 2048, 4, [64, 1]), (4, 300, 3, [8])])
def test_push_exchange_kernels_deliver_the_gathered_rows(F, B, G, dims):
    """Push-mode forward on G virtual ranks of one device: pack (ids + sample index into the owners' lists) then every
    owner's gather_push leave in each requester's outputs exactly table[id] (zeros for negative ids) — bit-exact, the
    values are copied."""
    from pytorchrec_b200 import ops
    rows = [997, 5, 4099, 64, 31][:F] + [257] * max(0, F - 5)
    C = (B + 15) // 16 * 16  # cannot overflow: a 5-row table sends everything to few owners
    gen = torch.Generator().manual_seed(7)
    full = [[torch.randn(rows[f], d, generator=gen) for f in range(F)] for d in dims]
    # owner o holds rows o::G of every table, in a row-strided buffer (as the interleaved optimizer layout)
    shards = [[[torch.zeros(max((rows[f] - o + G - 1) // G, 1), 2 * d + 4, device=DEV) for f in range(F)] for d in dims]
              for o in range(G)]
    for o in range(G):
        for k, d in enumerate(dims):
            for f in range(F):
                part = full[k][f][o::G]
                shards[o][k][f][:part.shape[0], :d] = part.to(DEV)
    own_ids = [torch.full((F * G * C,), -1, dtype=torch.int64, device=DEV) for _ in range(G)]
    own_b = [torch.zeros(F * G * C, dtype=torch.int32, device=DEV) for _ in range(G)]
    peer_ids = torch.tensor([t.data_ptr() for t in own_ids], dtype=torch.int64).to(DEV)
    peer_b = torch.tensor([t.data_ptr() for t in own_b], dtype=torch.int64).to(DEV)
    outs = [[torch.full((B, F * d), 7.0, device=DEV) for d in dims] for _ in range(G)]
    peer_outs = [torch.tensor([outs[r][k].data_ptr() for r in range(G)], dtype=torch.int64).to(DEV) for k in range(len(dims))]
    all_ids = []
    for r in range(G):
        ids = torch.stack([torch.randint(-1, rows[f], (B,), generator=gen) for f in range(F)])
        all_ids.append(ids)
        ovf = torch.zeros(1, dtype=torch.int32, device=DEV)
        ops.a2a_pack_by_owner_push(ids.to(DEV), F, B, G, C, r, peer_ids, peer_b, outs[r], dims, ovf)
        assert ovf.item() == 0
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    for o in range(G):
        tabs = [torch.tensor([shards[o][k][f].data_ptr() for f in range(F)], dtype=torch.int64).to(DEV) for k in range(len(dims))]
        shard_rows = torch.tensor([(rows[f] - o + G - 1) // G if rows[f] > o else 0 for f in range(F)], dtype=torch.int64).to(DEV)
        ops.gather_push(tabs, peer_outs, [2 * d + 4 for d in dims], [F * d for d in dims], dims, own_ids[o], own_b[o],
                        shard_rows, F, G, C, err_flag=err)
    assert err.item() == 0
    for r in range(G):
        for k, d in enumerate(dims):
            got = outs[r][k].cpu().view(B, F, d)
            for f in range(F):
                idf = all_ids[r][f]
                want = torch.zeros(B, d)
                want[idf >= 0] = full[k][f][idf[idf >= 0]]
                assert torch.equal(got[:, f], want), (r, k, f)


@pytest.mark.parametrize("G,D,stride_mult", [(2, 16, 1), (8, 16, 2), (4, 1, 1), (3, 64, 2)])
def test_sharded_gather_reads_row_wise_shards_bit_exact(G, D, stride_mult):
    from pytorchrec_b200 import ops
    rows = [1000, 7, 4099, 1]
    B, T = 777, 4
    gen = torch.Generator().manual_seed(G * 100 + D)
    full = [torch.randn(r, D, generator=gen) for r in rows]
    keep, ptrs = [], []
    for t in range(T):
        row_ptrs = []
        for g in range(G):
            part = full[t][g::G]
            buf = torch.zeros(max(part.shape[0], 1), stride_mult * D, device=DEV)  # weight | (state) interleaved
            buf[:part.shape[0], :D] = part.to(DEV)
            keep.append(buf)
            row_ptrs.append(buf.data_ptr())
        ptrs.append(row_ptrs)
    shard_ptrs = torch.tensor(ptrs, dtype=torch.int64).to(DEV)
    rows_t = torch.tensor(rows, dtype=torch.int64).to(DEV)
    lay = ops.FeatureLayout([dict(table=t, bag_len=1, neg_is_pad=True) for t in range(T)], D, T)
    ids = torch.stack([torch.randint(0, rows[t], (B,), generator=gen) for t in range(T)])
    ids[0, 5] = -1                                  # "no lookup": zero row, no error
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = ops.gather_fwd_sharded(shard_ptrs, rows_t, G, stride_mult * D, lay, ids.to(DEV).view(-1), B, err_flag=err)
    want = torch.stack([full[t][ids[t].clamp(min=0)] for t in range(T)], dim=1)
    want[5, 0] = 0
    assert torch.equal(out.cpu().view(B, T, D), want) and err.item() == 0
    ids[2, 9] = rows[2]                             # out of range against the GLOBAL height
    ops.gather_fwd_sharded(shard_ptrs, rows_t, G, stride_mult * D, lay, ids.to(DEV).view(-1), B, err_flag=err)
    assert err.item() == 1


def test_dcn_matches_oracle_twin_within_bf16_tolerance():
    """DCN-v2: bf16 tensor-core cross layers vs the fp32 oracle twin (north_star tolerance 1e-2 for bf16 variants)."""
    from pytorchrec_b200.model import DCN
    scols, dcols, lab, rows = _ctr_setup(F=6, nd=3)
    D, B = 16, 512
    prod = DCN(scols, dcols, lab, D, 3, [64, 32], random_seed=5)
    ref = ref_models.DCNRef(5, scols, dcols, lab, D, 3, [64, 32])
    for (k, v), (k2, v2) in zip(prod.state_dict().items(), ref.state_dict().items()):
        assert k == k2 and torch.equal(v, v2), (k, k2)
    prod.compile(SparseAdagrad(prod.get_parameters(), lr=0.02), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
    ref.compile(torch.optim.Adagrad(ref.get_parameters(), lr=0.02), torch.nn.BCEWithLogitsLoss())
    for s in range(3):
        batch = _ctr_batch(rows, len(dcols), B, seed=300 + s)
        with torch.no_grad():
            pl, _ = prod.test_step(batch)
            ref.eval()
            rl, _ = ref(batch)
        scale = max(1.0, rl.abs().max().item())
        np.testing.assert_allclose(pl.cpu().numpy(), rl.numpy(), rtol=1e-2, atol=1e-2 * scale)
        lp, lr_ = prod.train_step(batch)["loss"].item(), ref.train_step(batch)["loss"].item()
        np.testing.assert_allclose(lp, lr_, rtol=1e-2)


def _din_setup():
    cols = dict(uid=Col(300, "uid"), iid=Col(500, "iid"), cid=Col(40, "cid"), his_iid=Col(500, "his_iid"),
                his_cid=Col(40, "his_cid"), his_len=Col(101, "his_len"), label=Col(2, "label"))
    return cols


def _din_batch(cols, B, L, seed):
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, L + 1, size=B)
    pad = np.arange(L)[None, :] >= lens[:, None]
    hi = rng.integers(1, 500, size=(B, L)); hi[pad] = 0
    hc = rng.integers(1, 40, size=(B, L)); hc[pad] = 0
    return {"uid": torch.from_numpy(rng.integers(1, 300, size=B)), "iid": torch.from_numpy(rng.integers(1, 500, size=B)),
            "cid": torch.from_numpy(rng.integers(1, 40, size=B)), "his_iid": torch.from_numpy(hi),
            "his_cid": torch.from_numpy(hc), "his_len": torch.from_numpy(lens),
            "label": torch.from_numpy(rng.integers(0, 2, size=B))}


def test_din_fused_key_gather_trains_identically(monkeypatch):
    """DIN with K4 reading the history rows from the tables by id (default) == DIN with the materialised [B, 1+L, 2D]
    lookup (PTREC_DIN_FUSED_GATHER=0): same losses and bit-identical weights after Adagrad steps, and the fused path
    launches no gather over the B * (1 + L) slots."""
    from pytorchrec_b200.model import DIN
    c = _din_setup()
    args = (c["uid"], c["iid"], c["cid"], c["his_iid"], c["his_cid"], c["his_len"], c["label"])

    def run(fused):
        monkeypatch.setenv("PTREC_DIN_FUSED_GATHER", "1" if fused else "0")
        m = DIN(*args, emb_size=16, layers=[64, 32], random_seed=11)
        m.compile(SparseAdagrad(m.get_parameters(), lr=0.02, eps=1e-6), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
        losses = [m.train_step(_din_batch(c, 200, 100, seed=40 + s))["loss"].item() for s in range(4)]
        return losses, {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}

    l1, sd1 = run(True)
    l0, sd0 = run(False)
    assert l1 == l0, (l1, l0)
    for k in sd1:
        assert torch.equal(sd1[k], sd0[k]), k


@pytest.mark.parametrize("opt_name", ["sgd", "adagrad"])
def test_din_matches_oracle_twin(opt_name):
    from pytorchrec_b200.model import DIN
    c = _din_setup()
    args = (c["uid"], c["iid"], c["cid"], c["his_iid"], c["his_cid"], c["his_len"], c["label"])
    prod = DIN(*args, emb_size=16, layers=[64, 32], random_seed=11)
    ref = ref_models.DINRef(11, *args, 16, [64, 32])
    for (k, v), (k2, v2) in zip(prod.state_dict().items(), ref.state_dict().items()):
        assert k == k2 and torch.equal(v, v2), (k, k2)
    if opt_name == "sgd":
        popt, ropt = SparseSGD(prod.get_parameters(), lr=0.2), torch.optim.SGD(ref.get_parameters(), lr=0.2)
    else:
        # eps well above the fp32 noise floor of the (cancelling) gradient sums: with the default 1e-10 an element
        # whose true gradient is ~0 moves by lr * noise / (|noise| + eps), i.e. by an arbitrary fraction of lr
        popt = SparseAdagrad(prod.get_parameters(), lr=0.02, eps=1e-6)
        ropt = torch.optim.Adagrad(ref.get_parameters(), lr=0.02, eps=1e-6)
    ref64 = copy.deepcopy(ref).fp64()
    ref64.compile(type(ropt)(ref64.get_parameters(), **{k: ropt.defaults[k] for k in ("lr", "eps") if k in ropt.defaults}),
                  torch.nn.BCEWithLogitsLoss())
    prod.compile(popt, torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
    ref.compile(ropt, torch.nn.BCEWithLogitsLoss())
    for s in range(3):
        batch = _din_batch(c, 96, 100, seed=500 + s)
        ref64.train_step(batch)
        with torch.no_grad():
            pl, _ = prod.test_step(batch)
            ref.eval()
            rl, _ = ref(batch)
        scale = max(1.0, rl.abs().max().item())
        np.testing.assert_allclose(pl.cpu().numpy(), rl.numpy(), rtol=1e-5, atol=1e-5 * scale)
        lp, lr_ = prod.train_step(batch)["loss"].item(), ref.train_step(batch)["loss"].item()
        np.testing.assert_allclose(lp, lr_, rtol=1e-5)
    sd64 = ref64.state_dict()
    sums = adagrad_sums(ref64) if opt_name == "adagrad" else {}
    for (k, v), (_, v2) in zip(prod.state_dict().items(), ref.state_dict().items()):
        # the unit's weight gradients are cancelling sums over B*L positions and Adagrad's g / (|g| + eps) step turns
        # their summation-order noise into visible differences: refereed by the fp64 twin (conftest)
        assert_as_exact_as_the_oracle(k, v.cpu().numpy(), v2.numpy(), sd64[k].numpy(), rtol=1e-5, atol=1e-5 * 3 * 0.2,
                                      adagrad=(sums[k].numpy(), 0.02, 3) if k in sums else None)


def test_fm_cfg1_full_size_matches_cpu_oracle():
    """BASELINE cfg1 at full size: FM, 26 tables x 1e5 rows x D16 (+26 first-order), 13 dense, batch 4096 —
    two SGD train steps on the GPU vs the CPU oracle twin (identical seeded init)."""
    from pytorchrec_b200.data import criteo_batch, criteo_columns
    sparse, dense, label = criteo_columns(26, 13, 100_000)
    prod = FM(sparse, dense, label, 16, random_seed=2020)
    ref = ref_models.FMRef(2020, sparse, dense, label, 16)
    prod.compile(SparseSGD(prod.get_parameters(), lr=0.1), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
    ref.compile(torch.optim.SGD(ref.get_parameters(), lr=0.1), torch.nn.BCEWithLogitsLoss())
    for s in range(2):
        batch = criteo_batch(4096, 26, 13, 100_000, seed=77 + s, dist="zipf" if s else "uniform")
        lp = prod.train_step(batch)["loss"].item()
        lr_ = ref.train_step(batch)["loss"].item()
        np.testing.assert_allclose(lp, lr_, rtol=1e-5)
    for (k, v), (_, v2) in zip(prod.state_dict().items(), ref.state_dict().items()):
        np.testing.assert_allclose(v.cpu().numpy(), v2.numpy(), rtol=1e-5, atol=2e-7, err_msg=k)
    prod.embeddings.check_index_errors()


def test_packed_ingest_equals_per_key_transfer():
    """N1: the single pinned packed H2D copy presents the same Dict[str, Tensor] as the reference's per-key .to()."""
    from pytorchrec_b200.utils.ingest import BatchPacker
    scols, dcols, lab, rows = _ctr_setup()
    batches = [_ctr_batch(rows, len(dcols), 300, seed=s) for s in range(5)]
    packer = BatchPacker(batches[0], DEV)
    for b in batches:  # more batches than staging buffers: exercises the event guard
        dev = packer.load(b)
        assert set(dev) == set(b)
        for k, v in b.items():
            assert dev[k].dtype == v.dtype and dev[k].shape == v.shape and torch.equal(dev[k].cpu(), v)
    assert packer.h2d_bytes >= sum(v.numel() * v.element_size() for v in batches[0].values())
    # model level: identical training with and without the packer
    res = []
    for packed in (True, False):
        m = DeepFM(scols, dcols, lab, 16, [32, 16], random_seed=9)
        m.packed_ingest = packed
        m.compile(SparseAdagrad(m.get_parameters(), lr=0.05), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
        res.append(([m.train_step(b)["loss"].item() for b in batches], m.state_dict()))
    assert res[0][0] == res[1][0]
    for (k, a), (_, b2) in zip(res[0][1].items(), res[1][1].items()):
        assert torch.equal(a, b2), k


def test_evaluate_with_device_side_ranks():
    """N3: evaluate() on [B, 1+neg] candidate lists (FunkSVD.py:56-65 shape) keeps scores on the GPU."""
    from pytorchrec_b200.metric import Hit, NDCG

    class Candidates(torch.utils.data.Dataset):
        def __init__(self, n, neg):
            g = torch.Generator().manual_seed(3)
            self.uid = torch.randint(1, 50, (n,), generator=g)
            self.iid = torch.randint(1, 80, (n, 1 + neg), generator=g)

        def __len__(self):
            return len(self.uid)

        def __getitem__(self, i):
            return {"uid": self.uid[i], "iid": self.iid[i], "label": torch.tensor(1)}

    m = FunkSVD(Col(50, "uid"), Col(80, "iid"), Col(2, "label"), 8, random_seed=4)
    m.compile(SparseSGD(m.get_parameters(), lr=0.1), torch.nn.MSELoss(), [Hit(10, 3), NDCG(10, 3)], DEV)
    ds = Candidates(300, 9)
    logs = m.evaluate(ds, batch_size=64)
    # host recomputation from the same scores
    with torch.no_grad():
        pred, _ = m.test_step({"uid": ds.uid, "iid": ds.iid, "label": torch.ones(300)})
    from pytorchrec_b200.metric import MetricList
    want = MetricList([Hit(10, 3), NDCG(10, 3)])(pred.cpu().numpy(), None)
    assert logs.keys() == want.keys() and all(abs(logs[k] - want[k]) < 1e-9 for k in logs)


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("sgd", dict(lr=0.05, weight_decay=0.01)),
                                     ("adagrad", dict(lr=0.05, lr_decay=0.01, weight_decay=0.01, eps=1e-10,
                                                      initial_accumulator_value=0.1)),
                                     ("adam", dict(lr=0.01, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01))])
def test_fused_dense_parameter_step_equals_torch_optim(name, kw):
    """K7: the one-launch dense update has torch.optim's arithmetic (same operation order, same state layout)."""
    from pytorchrec_b200.optim import SparseAdagrad, SparseAdam, SparseSGD
    ours_cls = {"sgd": SparseSGD, "adagrad": SparseAdagrad, "adam": SparseAdam}[name]
    ref_cls = {"sgd": torch.optim.SGD, "adagrad": torch.optim.Adagrad, "adam": torch.optim.Adam}[name]
    gen = torch.Generator().manual_seed(7)
    shapes = [(400, 429), (400,), (1,), (3, 5, 7), (2049,)]
    a = [torch.nn.Parameter(torch.randn(*s, generator=gen).to(DEV)) for s in shapes]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    ours, ref = ours_cls(a, **kw), ref_cls(b, **kw)
    for step in range(4):
        for i, (p, q) in enumerate(zip(a, b)):
            if step == 2 and i == 3:
                p.grad = q.grad = None          # a parameter without a gradient is skipped
                continue
            g = torch.randn(*p.shape, generator=gen).to(DEV)
            p.grad, q.grad = g.clone(), g.clone()
        launches = _lib_launches()
        ours.step()
        # one launch, not the stock multi-tensor cascade (two once a parameter's own step counter lags behind)
        assert _lib_launches() == launches + (2 if step == 3 and name != "sgd" else 1)
        ref.step()
        for p, q in zip(a, b):
            np.testing.assert_allclose(p.detach().cpu().numpy(), q.detach().cpu().numpy(), rtol=2e-6, atol=1e-7)
    sd = ours.state_dict()["dense"]["state"]
    rsd = ref.state_dict()["state"]
    assert set(sd) == set(rsd)
    for k in rsd:
        for n, v in rsd[k].items():
            np.testing.assert_allclose(sd[k][n].cpu().numpy(), v.cpu().numpy(), rtol=2e-6, atol=1e-7)


def _lib_launches():
    from pytorchrec_b200 import _lib
    return _lib.load().ptrec_launch_count()


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("packed_host", [False, True])
def test_prefetched_batches_train_exactly_like_direct_ones(graph, packed_host):
    """N1: prefetch(batch k+1) during step k (copy stream, double-buffered) gives bit-identical training — from plain
    host dicts and from batches the loader assembled in one pinned buffer (``IModel.pack_host``: one DMA per step)."""
    scols, dcols, lab, rows = _ctr_setup(F=5, nd=2)
    models = []
    for _ in range(2):
        m = DeepFM(scols, dcols, lab, 8, [16, 8], random_seed=11)
        m.compile(SparseAdagrad(m.get_parameters(), lr=0.05), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
        if graph:
            m.enable_cuda_graph(True, warmup=2)
        models.append(m)
    direct, pre = models
    batches = [_ctr_batch(rows, len(dcols), 128, seed=40 + s) for s in range(7)]
    if packed_host:
        batches = [pre.pack_host(b) for b in batches]
        assert batches[0].buffer.is_pinned() and all(not v.is_cuda for v in batches[0].values())
    pre.prefetch(batches[0])
    for i, b in enumerate(batches):
        la = direct.train_step({k: v.clone() for k, v in b.items()})["loss"]
        lb = pre.train_step(b)["loss"]
        if i + 1 < len(batches):
            pre.prefetch(batches[i + 1])
        assert torch.equal(la, lb), i
    with torch.no_grad():
        pre.prefetch(batches[0])
        pa, _ = direct.test_step(batches[0])
        pb, _ = pre.test_step(batches[0])
    assert torch.equal(pa, pb)
    for (k, v), (_, w) in zip(direct.state_dict().items(), pre.state_dict().items()):
        assert torch.equal(v, w), k


@pytest.mark.gpu
def test_staged_device_batches_train_exactly_like_host_ones():
    """N1: IModel.stage() keeps a batch resident in one packed device buffer; graph and eager steps accept it."""
    scols, dcols, lab, rows = _ctr_setup(F=5, nd=2)
    models = []
    for _ in range(2):
        m = DeepFM(scols, dcols, lab, 8, [16, 8], random_seed=12)
        m.compile(SparseAdagrad(m.get_parameters(), lr=0.05), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
        m.enable_cuda_graph(True, warmup=2)
        models.append(m)
    a, b = models
    batches = [_ctr_batch(rows, len(dcols), 128, seed=60 + s) for s in range(5)]
    staged = [b.stage(x) for x in batches]
    assert all(v.is_cuda for v in staged[0].values()) and staged[0].buffer is not None
    for x, y in zip(batches, staged):
        assert torch.equal(a.train_step(x)["loss"], b.train_step(y)["loss"])
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, w), k


@pytest.mark.gpu
def test_dense_layer_size_heuristic_and_parity_of_both_paths():
    """Dense = nn.Linear -> ReLU on cuBLAS below TC_MIN_MACS, on K6 above; both give the same forward / gradients."""
    from pytorchrec_b200 import _lib
    from pytorchrec_b200.model.layer import dense
    torch.manual_seed(3)
    layer = dense.Dense(429, 400, "relu", 0.0).to(DEV)
    x = torch.randn(2048, 429, device=DEV, requires_grad=True)
    outs = []
    for min_macs in (1 << 62, 0):          # cuBLAS, then K6
        dense.TC_MIN_MACS = min_macs
        before = _lib.load().ptrec_launch_count()
        y = layer(x)
        (y * torch.linspace(-1, 1, 400, device=DEV)).sum().backward()
        used_k6 = _lib.load().ptrec_launch_count() > before
        assert used_k6 == (min_macs == 0)
        outs.append((y.detach().clone(), x.grad.clone(), layer.linear.weight.grad.clone(), layer.linear.bias.grad.clone()))
        x.grad = None
        layer.zero_grad()
    for a, b in zip(*outs):
        scale = b.abs().max().item()
        assert (a - b).abs().max().item() <= 2e-6 * max(1.0, scale) * 30, (a - b).abs().max().item()
    dense.TC_MIN_MACS = 1 << 28
    assert 2048 * 429 * 400 >= dense.TC_MIN_MACS and 8192 * 96 * 200 < dense.TC_MIN_MACS
