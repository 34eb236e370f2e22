"""Which elements of the cfg2 full-size parity test deviate most from the fp64 oracle, and why (hits per step, gradient
scale)?  Diagnostic for tests/test_gpu_fullsize.py."""
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
snap = []
for b in batches:
    prod.train_step(b); ref.train_step(b); ref64.train_step(b)
    snap.append({k: v.detach().cpu().clone() for k, v in prod.state_dict().items() if k.startswith("embeddings.3.")})
k = "embeddings.3.weight"
a, b32, b64 = prod.state_dict()[k].cpu().double(), ref.state_dict()[k].double(), ref64.state_dict()[k]
s64 = {n: ref64.opt.state[p]["sum"] for n, p in ref64.named_parameters()}[k]
e = (a - b64).abs()
g_elem = (s64 / 2).sqrt()
g_rms = (s64[s64 > 0].mean() / 2).sqrt()
bound = torch.where(s64 > 0, torch.clamp(lr * 2 * 16e-5 * g_rms / (g_elem + 1e-30), max=2 * lr * 2), torch.zeros_like(e)) + 1e-5 * b64.abs() + 1e-5 * 2 * lr
idx = torch.topk((e / bound).flatten(), 12).indices
print("ranked by error / bound")
ids = [bb["C4"] for bb in batches]
for i in idx.tolist():
    r, c = divmod(i, D)
    print(f"row {r} col {c}: err cuda {e[r, c]:.3e} cpu32 {(b32 - b64).abs()[r, c]:.3e}  hits step1 {(ids[0] == r).sum().item()} step2 {(ids[1] == r).sum().item()}"
          f"  bound {bound[r, c]:.3e}  sqrt(sum64) {s64[r, c].sqrt():.3e}  w64 {b64[r, c]:.5f}  after step1: cuda {snap[0][k][r, c]:.6f}")
print("rms sqrt(sum64/2) over touched:", (s64[s64 > 0] / 2).mean().sqrt().item())
