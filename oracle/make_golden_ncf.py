"""Generate tests/golden/reference_ncf.npz by EXECUTING THE UNMODIFIED REFERENCE NCF (torchrec/model/NCF.py).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run here (CPU container), never on the GPU box:
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_ncf.py

NCF is the one reference model whose forward runs the ``MLP`` / ``Dense`` tower (torchrec/model/layer/MLP.py:8-23,
Dense.py:4-24) and the concat -> ``Linear(., 1, bias=False)`` head (NCF.py:51,68-74): its run pins the oracle's
restatement of those layers, and — on the GPU — the product's tensor-core Linear path (K6), by the reference's own
numbers.  Recorded per case: inputs, seeded initial ``state_dict``, predictions / targets / losses over the
reference's ``IModel.compile`` / ``train_step`` (IModel.py:94-125) and the final ``state_dict``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import OUT, _import_reference, _run_steps  # noqa: E402


def main():
    R = _import_reference()
    Col = R["Col"]
    rng = np.random.default_rng(20201019)
    n_u, n_i, D, B, steps = 41, 59, 8, 32, 3
    layers = [16, 8]
    uid, iid, label = Col(n_u, "uid"), Col(n_i, "iid"), Col(2, "label")
    ndcg = R["NDCG"](user_sample_n=2, k=1)
    rec, cases = {}, []
    for n_cand, opt_name, opt_kw in [(2, "sgd", dict(lr=0.5)), (2, "adam", dict(lr=0.01)), (5, "sgd", dict(lr=0.5))]:
        tag = f"ncf_n{n_cand}_{opt_name}"
        batches = [{"uid": torch.from_numpy(rng.integers(1, n_u, size=B).astype(np.int32)),
                    "iid": torch.from_numpy(rng.integers(1, n_i, size=(B, n_cand)).astype(np.int32)),
                    "label": torch.from_numpy(rng.integers(0, 2, size=B).astype(np.int32))} for _ in range(steps)]
        model = R["NCF"](random_seed=2020, uid_column=uid, iid_column=iid, label_column=label, emb_size=D,
                         layers=layers, dropout=0.0)
        opt = R["get_optimizer"](opt_name)(params=model.get_parameters(), **opt_kw)
        loss = R["BPRLoss"]() if n_cand == 2 else torch.nn.BCEWithLogitsLoss()
        for s, b in enumerate(batches):
            for k, v in b.items():
                rec[f"{tag}/batch{s}/{k}"] = v.numpy().copy()
        _run_steps(model, opt, loss, ndcg, batches, rec, tag)
        cases.append(tag)
    rec["cases"] = np.array(cases)
    rec["dims"] = np.array([n_u, n_i, D, B, steps] + layers)
    rec["pinned_by_reference"] = np.int32(1)
    rec["meta/torch_version"] = np.array(torch.__version__)
    path = os.path.join(OUT, "reference_ncf.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
