import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200.data import criteo_batch, criteo_columns
from pytorchrec_b200.metric import LogLoss
from pytorchrec_b200.optim import SparseAdagrad
from pytorchrec_b200.model import DeepFM
dev = torch.device("cuda:0")
sparse, dense, label = criteo_columns(26, 13, 1_000_000)
model = DeepFM(sparse, dense, label, 16, [400, 400, 400], random_seed=2020, table_device=dev)
model.compile(SparseAdagrad(params=model.get_parameters(), lr=0.01), torch.nn.BCEWithLogitsLoss(), [LogLoss()], dev)
model.enable_cuda_graph(True)
host = [criteo_batch(16384, 26, 13, 1_000_000, seed=1000 + i, pin=True) for i in range(8)]
for i in range(6):
    model.train_step(host[i % 8])
torch.cuda.synchronize()
def loop(prefetch, steps=50):
    tp = tt = ti = 0.0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if prefetch: model.prefetch(host[0])
    for i in range(steps):
        a = time.perf_counter(); logs = model.train_step(host[i % 8]); b = time.perf_counter()
        if prefetch and i + 1 < steps: model.prefetch(host[(i + 1) % 8])
        c = time.perf_counter(); logs["loss"].item(); d = time.perf_counter()
        tt += b - a; tp += c - b; ti += d - c
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"prefetch={prefetch}: {1e3 * (t1 - t0) / steps:.3f} ms/step  cpu: train_step {1e3 * tt / steps:.3f} prefetch {1e3 * tp / steps:.3f} item {1e3 * ti / steps:.3f}")
for pf in (False, True, False, True):
    loop(pf)
