"""Train-step timing of the other BASELINE.json configs (cfg3 DCN-v2, cfg4 DIN) on one B200: samples/s with
device-resident batches, whole-step CUDA graph.  Not the driver's bench line (that is bench.py / cfg2)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pytorchrec_b200.data import criteo_batch, criteo_columns
from pytorchrec_b200.feature_column import CategoricalColumnWithIdentity as Col
from pytorchrec_b200.metric import LogLoss
from pytorchrec_b200.model import DCN, DIN
from pytorchrec_b200.optim import SparseAdagrad

dev = torch.device("cuda:0")


def run(model, batches, steps=30, warm=6, graph=True):
    model.compile(SparseAdagrad(model.get_parameters(), lr=0.01), torch.nn.BCEWithLogitsLoss(), [LogLoss()], dev)
    if graph:
        model.enable_cuda_graph(True)
    for i in range(warm):
        model.train_step(batches[i % len(batches)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        model.train_step(batches[i % len(batches)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def dcn():
    B, rows, D = 32768, 1_000_000, 32
    sparse, dense, label = criteo_columns(26, 13, rows)
    m = DCN(sparse, dense, label, D, 3, [1024, 1024, 1024], random_seed=1, table_device=dev)
    batches = [{k: v.to(dev) for k, v in criteo_batch(B, 26, 13, rows, seed=i).items()} for i in range(4)]
    ms = run(m, batches)
    return {"config": "cfg3 DCN-v2: 26x1e6xD32, d=845, 3 cross layers bf16 tcgen05, DNN 1024x3 fp32, B=32768", "ms_per_step": ms,
            "samples_per_s": B / ms * 1e3}


def din():
    B, L, D = 8192, 100, 16
    c = dict(uid=Col(603668, "uid"), iid=Col(367982, "iid"), cid=Col(1600, "cid"), hi=Col(367982, "his_iid"),
             hc=Col(1600, "his_cid"), hl=Col(L + 1, "his_len"), label=Col(2, "label"))
    m = DIN(c["uid"], c["iid"], c["cid"], c["hi"], c["hc"], c["hl"], c["label"], emb_size=D, layers=[200, 80],
            random_seed=1, table_device=dev)
    batches = []
    for s in range(4):
        rng = np.random.default_rng(s)
        lens = rng.integers(1, L + 1, size=B)
        pad = np.arange(L)[None, :] >= lens[:, None]
        hi = rng.integers(1, 367982, size=(B, L)); hi[pad] = 0
        hc = rng.integers(1, 1600, size=(B, L)); hc[pad] = 0
        b = {"uid": rng.integers(1, 603668, size=B), "iid": rng.integers(1, 367982, size=B), "cid": rng.integers(1, 1600, size=B),
             "his_iid": hi, "his_cid": hc, "his_len": lens, "label": rng.integers(0, 2, size=B)}
        batches.append({k: torch.from_numpy(v).to(dev) for k, v in b.items()})
    ms = run(m, batches)
    return {"config": "cfg4 DIN: Amazon-Books-shaped, L=100, D=16 (q/k 32), unit 80-40, B=8192", "ms_per_step": ms,
            "samples_per_s": B / ms * 1e3}


if __name__ == "__main__":
    out = {}
    for name in (sys.argv[1:] or ["dcn", "din"]):
        out[name] = {"dcn": dcn, "din": din}[name]()
        print(name, json.dumps(out[name]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/bench_models.json", "w"), indent=1)
