"""K1 / K2a / K2b alone at the cfg2 (D16, B 16384) and cfg5-per-GPU (D64, B 65536) shapes, every selectable variant:
   python tools/bench_emb_kernels.py [cfg2] [cfg5]
CUDA events on the launching stream, a different id batch per launch, tables >> L2.  Writes
gpurun_out/emb_kernels.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
peak = 6535.4
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]


def timeit(fn, reps=40, warm=4):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def run(name, F, R, D, B, zipf=False):
    g = torch.Generator(device=dev).manual_seed(1)
    bufs = [torch.randn(R, 2 * D, device=dev, generator=g) for _ in range(F)]
    for b in bufs:
        b[:, D:].abs_()
    tables = [b[:, :D] for b in bufs]
    state = [b[:, D:] for b in bufs]
    lay = ops.FeatureLayout([dict(table=f, bag_len=1) for f in range(F)], D, F)
    ts = ops.TableSet().refresh(tables)
    p1 = ops.make_ptr_array(state)
    nb = 6
    if zipf:
        import numpy as np
        rng = np.random.default_rng(0)
        id_batches = [torch.from_numpy((1 + (rng.zipf(1.05, size=F * B) - 1) % (R - 1)).astype("int64")).to(dev) for _ in range(nb)]
    else:
        id_batches = [torch.randint(1, R, (F * B,), device=dev, generator=g) for _ in range(nb)]
    go = torch.randn(B, F * D, device=dev, generator=g)
    out = torch.empty(B, F * D, device=dev)
    args = _lib.OptimArgs(kind=_lib.OPT_ADAGRAD, step=1, lr=0.0, eps=1e-10, beta1=0, beta2=0, weight_decay=0, lr_decay=0)
    lookups = F * B
    res = {}
    t = timeit(lambda i: ops.gather_pool_fwd(ts, lay, id_batches[i % nb], None, B, out=out))
    by = lookups * 8 + 2 * lookups * D * 4
    res["gather"] = dict(us=t * 1e6, GBps=by / t / 1e9, frac=by / t / 1e9 / peak)
    for mode, label in ((0, "sort_3P+3_launches"), (1, "sort_one_sweep")):
        lib.ptrec_set_one_sweep_sort(mode)
        t = timeit(lambda i: ops.sort_dedup(ts, lay, id_batches[i % nb], None, B))
        res[label] = dict(us=t * 1e6)
    srts = [ops.sort_dedup(ts, lay, ids, None, B) for ids in id_batches]
    U = sum(int(s.n_seg.item()) for s in srts) / nb
    by = lookups * (4 + D * 4) + U * (8 + 4 * D * 4)
    t = timeit(lambda i: ops.bwd_fused(ts, p1, None, lay, B, srts[i % nb], go, None, args))
    res["update"] = dict(us=t * 1e6, GBps=by / t / 1e9, frac=by / t / 1e9 / peak, unique=U)
    print(name, json.dumps({k: {a: round(b, 3) for a, b in v.items()} for k, v in res.items()}), flush=True)
    return res


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg5"]
    out = {}
    if "cfg2" in which:
        out["cfg2_D16_B16384"] = run("cfg2", 26, 1_000_000, 16, 16384)
        out["cfg2_D16_B16384_zipf"] = run("cfg2_zipf", 26, 1_000_000, 16, 16384, zipf=True)
        out["cfg2_firstorder_D1"] = run("cfg2_D1", 26, 1_000_000, 1, 16384)
    if "cfg5" in which:
        out["cfg5_D64_B65536"] = run("cfg5", 26, 2_000_000, 64, 65536)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/emb_kernels.json", "w"), indent=1)
