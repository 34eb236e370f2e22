"""GPU probe: effect of cudaLimitMaxL2FetchGranularity on the gather / fused-update kernels (cfg2 shape)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
torch.zeros(1, device=dev)
F, R, B = 26, 1_000_000, 16384
res = {}
for D in (16, 1, 64):
    rows = R if D != 64 else 400_000
    tables = [torch.randn(rows, D, device=dev) for _ in range(F)]
    state = [torch.zeros_like(t) for t in tables]
    lay = ops.FeatureLayout([dict(table=f, bag_len=1) for f in range(F)], D, F)
    ts = ops.TableSet().refresh(tables)
    p1 = ops.make_ptr_array(state)
    idb = [torch.randint(0, rows, (F * B,), device=dev) for _ in range(8)]
    go = torch.randn(B, F * D, device=dev)
    out = torch.empty(B, F * D, device=dev)
    args = _lib.OptimArgs(kind=_lib.OPT_ADAGRAD, step=1, lr=0.0, eps=1e-10, beta1=0, beta2=0, weight_decay=0, lr_decay=0)
    srts = [ops.sort_dedup(ts, lay, i, None, B) for i in idb]
    for gran in (128, 64, 32):
        print("set", gran, lib.ptrec_set_l2_fetch_granularity(gran), "now", lib.ptrec_get_l2_fetch_granularity())
        def t(fn, reps=50):
            for i in range(5): fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps): fn(i)
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps * 1e3
        tg = t(lambda i: ops.gather_pool_fwd(ts, lay, idb[i % 8], None, B, out=out))
        tu = t(lambda i: ops.bwd_fused(ts, p1, None, lay, B, srts[i % 8], go, None, args))
        res[f"D{D}_gran{gran}"] = dict(gather_us=tg, update_us=tu)
        print(D, gran, "gather %.1f us  update %.1f us" % (tg, tu), flush=True)
    del tables, state
json.dump(res, open("gpurun_out/l2_granularity.json", "w"), indent=1)
