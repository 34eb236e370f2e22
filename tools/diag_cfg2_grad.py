"""Per-sample error of dL/dv (the gradient entering the embedding backward) in step 2 of the cfg2 parity test:
CUDA vs the fp64 oracle, and the fp32 oracle vs the fp64 oracle."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ref_models
from pytorchrec_b200.data import criteo_batch, criteo_columns
from pytorchrec_b200.metric import LogLoss
from pytorchrec_b200.model import DeepFM
from pytorchrec_b200.optim import SparseAdagrad

DEV = torch.device("cuda:0")
rows, B, D, layers, lr = 1_000_000, 16384, 16, [400, 400, 400], 0.01
sparse, dense, label = criteo_columns(26, 13, rows)
prod = DeepFM(sparse, dense, label, D, layers, random_seed=2020)
ref = ref_models.DeepFMRef(2020, sparse, dense, label, D, layers)
ref64 = copy.deepcopy(ref).fp64()
BCE = torch.nn.BCEWithLogitsLoss
prod.compile(SparseAdagrad(prod.get_parameters(), lr=lr), BCE(), [LogLoss()], DEV)
ref.compile(torch.optim.Adagrad(ref.get_parameters(), lr=lr), BCE())
ref64.compile(torch.optim.Adagrad(ref64.get_parameters(), lr=lr), BCE())
batches = [criteo_batch(B, 26, 13, rows, seed=4100 + s, dist=d) for s, d in enumerate(("uniform", "zipf"))]
prod.train_step(batches[0]); ref.train_step(batches[0]); ref64.train_step(batches[0])
grads = {}
def _fwd_hook(m, i, o):
    if o.requires_grad:
        o.register_hook(lambda g: grads.__setitem__("cuda", g.detach().cpu().double()))
    return None
h = prod.embeddings.register_forward_hook(_fwd_hook)
def hook_ref(model, key):
    orig = model._parts
    def parts(data):
        v, x, logit = orig(data)
        if v.requires_grad:
            v.register_hook(lambda g: grads.__setitem__(key, g.detach().double()))
        return v, x, logit
    model._parts = parts
hook_ref(ref, "cpu32"); hook_ref(ref64, "cpu64")
zc = prod.test_step(batches[1])[0].detach().cpu().double()
prod.train_step(batches[1]); ref.train_step(batches[1]); ref64.train_step(batches[1])
g64 = grads["cpu64"].reshape(B, -1)
for k in ("cuda", "cpu32"):
    g = grads[k].reshape(B, -1)
    err = (g - g64).norm(dim=1) / g64.norm(dim=1)
    top = torch.topk(err, 6)
    print(k, "per-sample relative error of dL/dv: median %.2e  p99 %.2e  max %.2e" % (err.median(), err.quantile(0.99), err.max()))
    for e, i in zip(top.values.tolist(), top.indices.tolist()):
        print(f"   sample {i}: rel err {e:.2e}  |g64| {g64[i].norm():.3e}  logit {zc[i]:.4f}  label {int(batches[1]['label'][i])}  C4 id {int(batches[1]['C4'][i])}")
print("median |g64| per sample: %.3e" % g64.norm(dim=1).median())
