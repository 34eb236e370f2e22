"""bf16 embedding tables (``PTREC_BF16``, include/ptrec_b200.h): rows stored as bf16, everything computed in fp32.

The reference stores its tables in fp32 (``nn.Embedding``, torchrec/model/FunkSVD.py:39-48); bf16 storage is this
library's option, so the oracle is the reference idiom applied to the same numbers: a bf16 row widened to fp32 is
exact, so the gather must be bit-exact against ``F.embedding`` on the widened table, and a fused optimizer step must
equal ``torch.optim`` on the widened table rounded back to bf16 — up to one bf16 ulp where the fp32 results of the two
summation orders straddle a rounding boundary."""
import numpy as np
import pytest
import torch

from oracle import ref_models, ref_ops
from pytorchrec_b200 import _lib, ops
from pytorchrec_b200.feature_column import CategoricalColumnWithIdentity as Col
from pytorchrec_b200.feature_column import NumericColumn
from pytorchrec_b200.metric import LogLoss
from pytorchrec_b200.model import DeepFM
from pytorchrec_b200.optim import SparseAdagrad, SparseAdam, SparseRowWiseAdagrad, SparseSGD

DEV = torch.device("cuda:0")
pytestmark = [pytest.mark.gpu]


def _ids(shape, rows, seed, pad_frac=0.0, hot=False):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, rows, size=shape)
    if hot:
        x = np.where(rng.random(shape) < 0.5, rng.integers(0, 3, size=shape), x)
    if pad_frac:
        x[rng.random(shape) < pad_frac] = 0
    return torch.from_numpy(x.astype(np.int64))


@pytest.mark.parametrize("D", [1, 2, 4, 8, 16, 32, 64, 128])
@pytest.mark.parametrize("B", [1, 33, 1024])
def test_bf16_onehot_gather_is_bit_exact(D, B):
    rows = [17, 1000, 5, 301]
    g = torch.Generator().manual_seed(D)
    weights = [torch.randn(r, D, generator=g).bfloat16() for r in rows]
    id_list = [_ids((B,), rows[t], 100 * D + t) for t in range(4)]
    layout = ops.FeatureLayout([dict(table=t, bag_len=1) for t in range(4)], D, 4)
    dw = [w.to(DEV) for w in weights]   # kept alive: the table set holds raw pointers
    tables = ops.TableSet().refresh(dw)
    assert tables.dtype == _lib.BF16
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    out, _ = ops.gather_pool_fwd(tables, layout, torch.cat(id_list).to(DEV), None, B, err_flag=err)
    ref = torch.stack([torch.nn.functional.embedding(id_list[t], weights[t].float()) for t in range(4)], 1)
    assert out.dtype == torch.float32 and torch.equal(out.view(B, 4, D).cpu(), ref) and err.item() == 0


@pytest.mark.parametrize("pooling", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("mask", ["pad", "lens"])
@pytest.mark.parametrize("B,L,D", [(130, 100, 16), (33, 50, 64), (9, 11, 1)])
def test_bf16_bag_pooling_matches_the_reference_idiom(pooling, mask, B, L, D):
    g = torch.Generator().manual_seed(L)
    w = torch.randn(211, D, generator=g).bfloat16()
    ids = _ids((B, L), 211, B + L, pad_frac=0.4)
    lens = torch.from_numpy(np.random.default_rng(L).integers(0, L + 1, size=B).astype(np.int64))
    layout = ops.FeatureLayout([dict(table=0, bag_len=L, pooling=pooling, mask=mask, lens_col=0 if mask == "lens" else -1)], D, 1)
    dw = [w.to(DEV)]
    tables = ops.TableSet().refresh(dw)
    lens_dev = lens.to(torch.int32).reshape(1, B).to(DEV) if mask == "lens" else None
    out, _ = ops.gather_pool_fwd(tables, layout, ids.reshape(-1).to(DEV), lens_dev, B)
    ref = ref_ops.pooled_lookup_ref(w.float(), ids, pooling, mask, lens)
    bound = 1e-5 * torch.nn.functional.embedding(ids, w.float()).abs().sum(1) + 1e-7
    assert ((out.cpu() - ref).abs() <= bound).all()


def _one_ulp_bf16(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """|a - b| <= one bf16 ulp of the larger magnitude (8 significand bits)."""
    return (a.float() - b.float()).abs() <= 2.0 ** -7 * torch.maximum(a.float().abs(), b.float().abs())


@pytest.mark.parametrize("opt_name", ["sgd", "adagrad", "rowwise", "adam"])
@pytest.mark.parametrize("D,B,rows", [(16, 4096, 1000), (64, 2048, 300), (1, 4096, 50), (8, 700, 100000)])
def test_bf16_fused_update_equals_torch_optim_on_the_widened_table(opt_name, D, B, rows):
    from pytorchrec_b200.model.layer import EmbeddingTable
    torch.manual_seed(D + B)
    table = EmbeddingTable(rows, D, dtype=torch.bfloat16).to(DEV)
    assert table.weight.dtype == torch.bfloat16
    w0 = table.weight.detach().cpu().clone()
    lr = 0.1
    mk = dict(sgd=lambda p: SparseSGD(p, lr=lr), adagrad=lambda p: SparseAdagrad(p, lr=lr),
              rowwise=lambda p: SparseRowWiseAdagrad(p, lr=lr), adam=lambda p: SparseAdam(p, lr=lr))[opt_name]
    opt = mk([table.weight])
    ref_w = torch.nn.Parameter(w0.float().clone())
    ref_opt = dict(sgd=lambda: torch.optim.SGD([ref_w], lr=lr), adagrad=lambda: torch.optim.Adagrad([ref_w], lr=lr),
                   rowwise=None, adam=lambda: torch.optim.SparseAdam([ref_w], lr=lr))[opt_name]
    ref_opt = ref_opt() if ref_opt else None
    row_state = torch.zeros(rows)
    for step in range(3):
        ids = _ids((B,), rows, 7 * step + D, hot=True)
        gy = torch.randn(B, D, generator=torch.Generator().manual_seed(step))
        out = table(ids.to(DEV))
        assert out.dtype == torch.float32
        out.backward(gy.to(DEV))
        opt.step()
        opt.zero_grad()
        dense_g = ref_ops.dense_embedding_grad_ref(ids, gy, rows)
        if opt_name == "rowwise":
            w = ref_w.data
            ref_ops.rowwise_adagrad_ref(w, row_state, dense_g, lr, 1e-10)
        elif opt_name == "adam":
            touched = dense_g.abs().sum(1) > 0
            idx = touched.nonzero().flatten()
            ref_w.grad = torch.sparse_coo_tensor(idx.unsqueeze(0), dense_g[idx], size=dense_g.shape).coalesce()
            ref_opt.step()
        else:
            ref_w.grad = dense_g
            ref_opt.step()
        ref_w.data = ref_w.data.bfloat16().float()       # the table is stored in bf16 between steps
        got = table.weight.detach().cpu()
        want = ref_w.data.bfloat16()
        same = got == want
        assert same.float().mean().item() >= 0.995, (opt_name, step, same.float().mean().item())
        assert bool(_one_ulp_bf16(got, want).all()), (opt_name, step)
        untouched = dense_g.abs().sum(1) == 0
        assert torch.equal(got[untouched], w0[untouched]) or step > 0   # rows without a lookup keep their bits
        ref_w.data = got.float()                          # carry the device bits forward: errors do not compound
        w0 = got.clone()


def test_deepfm_with_bf16_tables_matches_the_oracle_twin_on_rounded_tables(monkeypatch):
    from pytorchrec_b200.model.layer import dense
    monkeypatch.setattr(dense, "TC_MIN_MACS", 0)
    F, nd, D, B = 6, 3, 16, 512
    rows = [50 + 13 * f for f in range(F)]
    scols = [Col(rows[f], f"C{f}") for f in range(F)]
    dcols = [NumericColumn(f"I{j}", 0.0, 1.0, 0.5, 0.25) for j in range(nd)]
    lab = Col(2, "label")
    prod = DeepFM(scols, dcols, lab, D, [32, 16], random_seed=7, table_dtype=torch.bfloat16)
    ref = ref_models.DeepFMRef(7, scols, dcols, lab, D, [32, 16])
    emb_keys = [k for k in ref.state_dict() if k.startswith(("embeddings.", "first_order."))]
    sd = {k: v.clone() for k, v in prod.state_dict().items()}
    for k in emb_keys:
        assert sd[k].dtype == torch.bfloat16
    ref.load_state_dict({k: v.float() for k, v in sd.items()})   # same numbers: bf16 values are fp32 values
    # SGD: the step is linear in the gradient, so plain tolerances hold (Adagrad's g / sqrt(sum g^2) needs the fp64
    # referee of tests/conftest.py; its bf16 arithmetic is covered kernel by kernel above)
    prod.compile(SparseSGD(prod.get_parameters(), lr=0.3), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
    ref.compile(torch.optim.SGD(ref.get_parameters(), lr=0.3), torch.nn.BCEWithLogitsLoss())
    rng = np.random.default_rng(3)
    for step in range(3):
        batch = {f"C{f}": torch.from_numpy(rng.integers(0, rows[f], size=B).astype(np.int32)) for f in range(F)}
        batch.update({f"I{j}": torch.from_numpy(rng.random(B).astype(np.float32)) for j in range(nd)})
        batch["label"] = torch.from_numpy(rng.integers(0, 2, size=B).astype(np.int32))
        lp = prod.train_step(batch)["loss"].item()
        lr_ = ref.train_step(batch)["loss"].item()
        np.testing.assert_allclose(lp, lr_, rtol=2e-5)
        got = prod.state_dict()
        rsd = ref.state_dict()
        for k in emb_keys:                                  # the twin's tables are rounded like the stored ones ...
            want = rsd[k].bfloat16()
            assert bool(_one_ulp_bf16(got[k].cpu(), want).all()), (step, k)
            rsd[k].copy_(got[k].cpu().float())              # ... and follow the device bits (no compounding)
        for k in rsd:
            if k not in emb_keys:
                np.testing.assert_allclose(got[k].cpu().numpy(), rsd[k].numpy(), rtol=1e-4, atol=1e-5, err_msg=k)
    prod.embeddings.check_index_errors()
