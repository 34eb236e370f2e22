"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (read-only at /root/reference).

Run here (CPU container), never on the GPU box:
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

What is recorded, per case: the seeded inputs, the reference model's initial ``state_dict`` (so seed
parity is checkable), its predictions / losses over a few ``IModel.train_step`` calls driven through
the reference's own ``compile`` (torchrec/model/IModel.py:94-125), and the final ``state_dict``.
Those files pin ``oracle/`` (tests/test_oracle_golden.py) and, on the GPU, the CUDA path itself.
Cases for components the reference lacks (FM / DeepFM / optimizers) are produced from the oracle
restatement and are labelled ``pinned_by_reference = 0``.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)


def _import_reference():
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):  # utils/const.py prints the hostname at import
        import torchrec  # noqa: F401
        from torchrec.feature_column import CategoricalColumnWithIdentity, CrossedColumn
        from torchrec.metric.NDCG import NDCG
        from torchrec.model.FunkSVD import FunkSVD
        from torchrec.model.NCF import NCF
        from torchrec.model.SVDPP import SVDPP
        from torchrec.model.utils import get_valid_his_index
        from torchrec.loss.BPRLoss import BPRLoss
        from torchrec.optim.optimizers import get_optimizer
    return dict(Col=CategoricalColumnWithIdentity, CrossedColumn=CrossedColumn, NDCG=NDCG, FunkSVD=FunkSVD,
                NCF=NCF, SVDPP=SVDPP, get_valid_his_index=get_valid_his_index, BPRLoss=BPRLoss,
                get_optimizer=get_optimizer)


def _sd(model, prefix):
    return {f"{prefix}/{k}": v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def _run_steps(model, opt, loss, metric, batches, rec, tag):
    model.compile(optimizer=opt, loss=loss, metrics=[metric], device=torch.device("cpu"))
    rec.update(_sd(model, f"{tag}/init"))
    for s, batch in enumerate(batches):
        model.eval()
        with torch.no_grad():
            pred, target = model(batch)
        rec[f"{tag}/pred{s}"] = pred.numpy().copy()
        rec[f"{tag}/target{s}"] = target.numpy().copy()
        logs = model.train_step(batch)
        rec[f"{tag}/loss{s}"] = np.float32(logs["loss"].item())
    rec.update(_sd(model, f"{tag}/final"))


def make_batches(rng, n_steps, B, n_u, n_i, L, pairwise):
    batches = []
    for _ in range(n_steps):
        iids = rng.integers(1, n_i, size=(B, L))
        cut = rng.integers(0, L + 1, size=B)  # right-padded with 0 = PAD; keep >= 1 valid for SVD++
        cut = np.maximum(cut, 1)
        iids[np.arange(L)[None, :] >= cut[:, None]] = 0
        b = {
            "uid": torch.from_numpy(rng.integers(1, n_u, size=B).astype(np.int32)),
            "iid": torch.from_numpy((rng.integers(1, n_i, size=(B, 2)) if pairwise else rng.integers(1, n_i, size=B)).astype(np.int32)),
            "iids": torch.from_numpy(iids.astype(np.int32)),
            "label": torch.from_numpy(rng.integers(0, 2, size=B).astype(np.int32)),
        }
        batches.append(b)
    return batches


def main():
    os.makedirs(OUT, exist_ok=True)
    R = _import_reference()
    Col = R["Col"]
    meta = dict(torch_version=torch.__version__, numpy_version=np.__version__)
    rng = np.random.default_rng(20201018)

    n_u, n_i, D, B, L, steps = 37, 53, 8, 24, 7, 3
    uid, iid, iids, label = Col(n_u, "uid"), Col(n_i, "iid"), Col(n_i, "iids"), Col(2, "label")
    ndcg = R["NDCG"](user_sample_n=2, k=1)

    # ---- SVD++ / FunkSVD through the reference's IModel.train_step, several optimizers -------------
    rec = {}
    cases = []
    for model_name, pairwise, opt_name, opt_kw in [
        ("svdpp", False, "sgd", dict(lr=0.5)),
        ("svdpp", True, "sgd", dict(lr=0.5)),
        ("svdpp", False, "adam", dict(lr=0.01)),
        ("funksvd", False, "sgd", dict(lr=0.5)),
        ("funksvd", True, "adamw", dict(lr=0.01, weight_decay=0.1)),
    ]:
        tag = f"{model_name}_{'pair' if pairwise else 'point'}_{opt_name}"
        batches = make_batches(rng, steps, B, n_u, n_i, L, pairwise)
        if model_name == "svdpp":
            model = R["SVDPP"](random_seed=2020, uid_column=uid, iid_column=iid, iids_column=iids,
                               label_column=label, emb_size=D)
        else:
            model = R["FunkSVD"](random_seed=2020, uid_column=uid, iid_column=iid, label_column=label, emb_size=D)
        opt = R["get_optimizer"](opt_name)(params=model.get_parameters(), **opt_kw)
        loss = R["BPRLoss"]() if pairwise else torch.nn.MSELoss()
        for s, b in enumerate(batches):
            for k, v in b.items():
                rec[f"{tag}/batch{s}/{k}"] = v.numpy().copy()
        _run_steps(model, opt, loss, ndcg, batches, rec, tag)
        cases.append(tag)
    rec["cases"] = np.array(cases)
    rec["dims"] = np.array([n_u, n_i, D, B, L, steps])
    rec["pinned_by_reference"] = np.int32(1)
    np.savez_compressed(os.path.join(OUT, "reference_mf.npz"), **rec, **{f"meta/{k}": np.array(v) for k, v in meta.items()})

    # ---- idioms: valid-history mask, masked mean pooling (SASRec.py:83,109-110), CrossedColumn --------
    rec = {}
    his = rng.integers(0, n_i, size=(B, L))
    his[rng.random((B, L)) < 0.4] = 0
    his_t = torch.from_numpy(his)
    valid = R["get_valid_his_index"](his_t)  # reference function, model/utils.py:5-10
    w = torch.randn(n_i, D, generator=torch.Generator().manual_seed(7))
    vec = torch.nn.functional.embedding(his_t, w)
    his_len = valid.long().sum(-1)
    pooled = (vec * valid.unsqueeze(-1).float()).sum(1) / his_len.unsqueeze(-1).float()  # SASRec.py:109-110
    rec.update(his=his, valid=valid.numpy(), weight=w.numpy(), pooled_mean=pooled.numpy(), his_len=his_len.numpy())
    cols = [Col(5, "a"), Col(7, "b"), Col(3, "c")]
    cross = R["CrossedColumn"](cols)
    batch = {c.feature_name: torch.from_numpy(rng.integers(0, c.category_num, size=B).astype(np.int32)) for c in cols}
    rec.update({f"cross/{k}": v.numpy() for k, v in batch.items()})
    rec["cross/out"] = cross.get_feature_data(batch).numpy()
    rec["cross/category_num"] = np.int64(cross.category_num)
    rec["cross/coefficients"] = np.array(cross.coefficients)
    rec["pinned_by_reference"] = np.int32(1)
    np.savez_compressed(os.path.join(OUT, "reference_idioms.npz"), **rec, **{f"meta/{k}": np.array(v) for k, v in meta.items()})

    # ---- oracle-only cases (components the reference lacks): regression pins, NOT reference pins ---------
    from oracle import ref_models
    from pytorchrec_b200.feature_column import CategoricalColumnWithIdentity as PCol, NumericColumn as PNum
    rec = {}
    F, nd, rows, Dm, Bm = 5, 3, [11, 23, 17, 29, 13], 8, 32
    scols = [PCol(rows[f], f"C{f}") for f in range(F)]
    dcols = [PNum(f"I{j}", 0.0, 1.0, 0.5, 0.25) for j in range(nd)]
    lab = PCol(2, "label")
    for name, opt_name in [("fm", "sgd"), ("deepfm", "adagrad")]:
        if name == "fm":
            model = ref_models.FMRef(2020, scols, dcols, lab, Dm)
        else:
            model = ref_models.DeepFMRef(2020, scols, dcols, lab, Dm, [16, 8])
        opt = (torch.optim.SGD(model.get_parameters(), lr=0.5) if opt_name == "sgd"
               else torch.optim.Adagrad(model.get_parameters(), lr=0.1))
        model.compile(opt, torch.nn.BCEWithLogitsLoss())
        tag = f"{name}_{opt_name}"
        rec.update(_sd(model, f"{tag}/init"))
        for s in range(3):
            b = {f"C{f}": torch.from_numpy(rng.integers(0, rows[f], size=Bm)) for f in range(F)}
            b.update({f"I{j}": torch.from_numpy(rng.random(Bm).astype(np.float32)) for j in range(nd)})
            b["label"] = torch.from_numpy(rng.integers(0, 2, size=Bm))
            for k, v in b.items():
                rec[f"{tag}/batch{s}/{k}"] = v.numpy().copy()
            with torch.no_grad():
                pred, _ = model(b)
            rec[f"{tag}/pred{s}"] = pred.numpy().copy()
            rec[f"{tag}/loss{s}"] = np.float32(model.train_step(b)["loss"].item())
        rec.update(_sd(model, f"{tag}/final"))
    rec["rows"] = np.array(rows)
    rec["dims"] = np.array([F, nd, Dm, Bm])
    rec["pinned_by_reference"] = np.int32(0)
    np.savez_compressed(os.path.join(OUT, "oracle_ctr.npz"), **rec, **{f"meta/{k}": np.array(v) for k, v in meta.items()})
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
