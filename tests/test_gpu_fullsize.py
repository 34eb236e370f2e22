"""One-step parity at the BASELINE.json config sizes ("full-size parity at cfg2 / cfg3 / cfg4").

cfg2 and cfg4 run at their full table sizes; cfg3 keeps its batch / dim / cross width / tower but cuts the 26 tables
to 1e5 rows so that the CPU oracle's three dense copies (weights, dense gradients, Adagrad sums) stay small — the
table height only changes which rows are hit, not the arithmetic.  Tolerances: fp32 1e-5 relative (segment-scaled,
with the fp64 oracle as referee where Adagrad amplifies summation-order noise: conftest.assert_as_exact_as_the_oracle),
bf16 cross layers 1e-2.  The CUDA side goes through the public API (IModel.train_step -> C ABI); the oracle is the
checker only."""
import copy

import numpy as np
import pytest
import torch

from conftest import adagrad_sums, assert_as_exact_as_the_oracle
from oracle import ref_models
from pytorchrec_b200.data import amazon_batch, amazon_columns, criteo_batch, criteo_columns
from pytorchrec_b200.metric import LogLoss
from pytorchrec_b200.model import DCN, DIN, DeepFM
from pytorchrec_b200.optim import SparseAdagrad

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
BCE = torch.nn.BCEWithLogitsLoss


def _same_init(prod, ref):
    for (k, v), (k2, v2) in zip(prod.state_dict().items(), ref.state_dict().items()):
        assert k == k2 and torch.equal(v.cpu(), v2), k


def _predict(model, batch):
    model.eval()
    with torch.no_grad():
        return model({k: v for k, v in batch.items()})[0]


def test_deepfm_cfg2_full_size_steps_match_cpu_oracle():
    """BASELINE cfg2: DeepFM, 26 tables x 1e6 rows x D16 (+26 first-order), 13 dense, DNN 400-400-400, batch 16384,
    fp32, fused sparse Adagrad — two train steps (uniform ids, then Zipf ids with heavy duplication) on the GPU against
    the CPU oracle twin (dense nn.Embedding gradients + dense torch.optim.Adagrad) from the identical seeded init."""
    rows, B, D, layers, lr = 1_000_000, 16384, 16, [400, 400, 400], 0.01
    sparse, dense, label = criteo_columns(26, 13, rows)
    prod = DeepFM(sparse, dense, label, D, layers, random_seed=2020)
    ref = ref_models.DeepFMRef(2020, sparse, dense, label, D, layers)
    _same_init(prod, ref)
    ref64 = copy.deepcopy(ref).fp64()
    prod.compile(SparseAdagrad(prod.get_parameters(), lr=lr), BCE(), [LogLoss()], DEV)
    ref.compile(torch.optim.Adagrad(ref.get_parameters(), lr=lr), BCE())
    ref64.compile(torch.optim.Adagrad(ref64.get_parameters(), lr=lr), BCE())
    touched = [torch.zeros(rows, dtype=torch.bool) for _ in range(26)]
    for s, dist in enumerate(("uniform", "zipf")):
        batch = criteo_batch(B, 26, 13, rows, seed=4100 + s, dist=dist)
        for f in range(26):
            touched[f][batch[f"C{f + 1}"]] = True
        pl = prod.test_step(batch)[0].detach().cpu().numpy()
        rl = _predict(ref, batch).numpy()
        np.testing.assert_allclose(pl, rl, rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(rl).max())))
        lp = prod.train_step(batch)["loss"].item()
        lr32 = ref.train_step(batch)["loss"].item()
        ref64.train_step(batch)
        np.testing.assert_allclose(lp, lr32, rtol=1e-5)
    sd64, sums = ref64.state_dict(), adagrad_sums(ref64)
    for (k, v), (_, v2) in zip(prod.state_dict().items(), ref.state_dict().items()):
        a, b = v.cpu(), v2
        if k.startswith(("embeddings.", "first_order.")):
            f = int(k.split(".")[1])
            # rows no lookup touched are bit-identical to the init (the dense CPU update adds exactly 0 to them)
            assert torch.equal(a[~touched[f]], b[~touched[f]]), k
            a, b, b64, s64 = a[touched[f]], b[touched[f]], sd64[k][touched[f]], sums[k][touched[f]]
        else:
            b64, s64 = sd64[k], sums[k]
        assert_as_exact_as_the_oracle(k, a.numpy(), b.numpy(), b64.numpy(), rtol=1e-5, atol=1e-5 * 2 * lr,
                                      adagrad=(s64.numpy(), lr, 2))
    prod.embeddings.check_index_errors()


def test_dcn_cfg3_shape_step_matches_cpu_oracle_within_bf16_tolerance():
    """BASELINE cfg3: DCN-v2, 26 fields x D32 + 13 dense (d = 845), 3 cross layers in bf16 on tcgen05, DNN 1024x3,
    batch 32768 (tables cut to 1e5 rows, see module docstring): logits, loss and one train step vs the fp32 oracle at
    the north star's 1e-2 tolerance for the bf16 variant."""
    rows, B, D, layers, lr = 100_000, 32768, 32, [1024, 1024, 1024], 0.01
    sparse, dense, label = criteo_columns(26, 13, rows)
    prod = DCN(sparse, dense, label, D, 3, layers, random_seed=7)
    ref = ref_models.DCNRef(7, sparse, dense, label, D, 3, layers)
    _same_init(prod, ref)
    prod.compile(SparseAdagrad(prod.get_parameters(), lr=lr), BCE(), [LogLoss()], DEV)
    ref.compile(torch.optim.Adagrad(ref.get_parameters(), lr=lr), BCE())
    for s in range(2):
        batch = criteo_batch(B, 26, 13, rows, seed=4300 + s, dist="zipf" if s else "uniform")
        pl = prod.test_step(batch)[0].detach().cpu().numpy()
        rl = _predict(ref, batch).numpy()
        np.testing.assert_allclose(pl, rl, rtol=1e-2, atol=1e-2 * max(1.0, float(np.abs(rl).max())))
        lp, lr32 = prod.train_step(batch)["loss"].item(), ref.train_step(batch)["loss"].item()
        np.testing.assert_allclose(lp, lr32, rtol=1e-2)
    # Adagrad moves a weight by at most lr per step (|g| / sqrt(sum g^2) <= 1), and a bf16-rounded gradient that is ~0
    # may take the opposite sign: two models can differ by up to 2*lr per step on isolated elements — the hard bound;
    # the bulk (median) must agree to 1e-2 of the two-step scale
    for (k, v), (_, v2) in zip(prod.state_dict().items(), ref.state_dict().items()):
        a, b = v.cpu().numpy(), v2.numpy()
        np.testing.assert_allclose(a, b, rtol=1e-2, atol=2 * 2 * lr, err_msg=k)
        d = np.abs(a - b)
        assert np.median(d) <= 1e-2 * 2 * lr + 1e-2 * np.median(np.abs(b)), (k, float(np.median(d)))
        assert (d > 1e-2 * np.abs(b) + 2 * lr).mean() <= 1e-3, (k, float((d > 1e-2 * np.abs(b) + 2 * lr).mean()))
    prod.embeddings.check_index_errors()


def test_din_cfg4_full_size_step_matches_cpu_oracle():
    """BASELINE cfg4: DIN, Amazon-Books-shaped tables (603 668 users, 367 982 items, 1 600 categories), histories of
    length 100, D16 per table (q / k = item || category = 32), unit 80-40, batch 8192, fp32: logits, loss and two
    train steps vs the CPU oracle twin, the fp64 twin refereeing the Adagrad weights."""
    B, L, D, layers, lr = 8192, 100, 16, [200, 80], 0.01
    cols = amazon_columns(L)
    prod = DIN(*cols, emb_size=D, layers=layers, random_seed=11)
    ref = ref_models.DINRef(11, *cols, D, layers)
    _same_init(prod, ref)
    ref64 = copy.deepcopy(ref).fp64()
    prod.compile(SparseAdagrad(prod.get_parameters(), lr=lr), BCE(), [LogLoss()], DEV)
    ref.compile(torch.optim.Adagrad(ref.get_parameters(), lr=lr), BCE())
    ref64.compile(torch.optim.Adagrad(ref64.get_parameters(), lr=lr), BCE())
    for s in range(2):
        batch = amazon_batch(B, L, seed=4400 + s)
        pl = prod.test_step(batch)[0].detach().cpu().numpy()
        rl = _predict(ref, batch).numpy()
        np.testing.assert_allclose(pl, rl, rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(rl).max())))
        lp, lr32 = prod.train_step(batch)["loss"].item(), ref.train_step(batch)["loss"].item()
        ref64.train_step(batch)
        np.testing.assert_allclose(lp, lr32, rtol=1e-5)
    sd64, sums = ref64.state_dict(), adagrad_sums(ref64)
    for (k, v), (_, v2) in zip(prod.state_dict().items(), ref.state_dict().items()):
        assert_as_exact_as_the_oracle(k, v.cpu().numpy(), v2.numpy(), sd64[k].numpy(), rtol=1e-5, atol=1e-5 * 2 * lr,
                                      adagrad=(sums[k].numpy(), lr, 2))
