"""Minimal driver for ncu captures of the embedding kernels (cfg2 shape by default)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import _lib, ops
D = int(os.environ.get("D", "16")); F = 26; R = int(os.environ.get("R", "1000000")); B = int(os.environ.get("B", "16384"))
dev = torch.device("cuda:0")
torch.zeros(1, device=dev)
if os.environ.get("L2G"):
    print("set l2 gran", _lib.load().ptrec_set_l2_fetch_granularity(int(os.environ["L2G"])), _lib.load().ptrec_get_l2_fetch_granularity())
bufs = [torch.randn(R, 2 * D, device=dev) for _ in range(F)]   # weight | Adagrad sum interleaved (product layout)
tables = [b[:, :D] for b in bufs]
state = [b[:, D:].abs_() for b in bufs]
lay = ops.FeatureLayout([dict(table=f, bag_len=1) for f in range(F)], D, F)
ts = ops.TableSet().refresh(tables)
p1 = ops.make_ptr_array(state)
go = torch.randn(B, F * D, device=dev)
out = torch.empty(B, F * D, device=dev)
args = _lib.OptimArgs(kind=_lib.OPT_ADAGRAD, step=1, lr=0.0, eps=1e-10, beta1=0, beta2=0, weight_decay=0, lr_decay=0)
for it in range(4):
    ids = torch.randint(0, R, (F * B,), device=dev)
    ops.gather_pool_fwd(ts, lay, ids, None, B, out=out)
    srt = ops.sort_dedup(ts, lay, ids, None, B)
    ops.bwd_fused(ts, p1, None, lay, B, srt, go, None, args)
if os.environ.get("DCN", "0") == "1":
    Bc, d = 32768, 848
    xs = [torch.randn(Bc, d, device=dev).bfloat16() * 0.5 for _ in range(2)]
    W = (torch.randn(d, d, device=dev) / d ** 0.5).bfloat16()
    b = torch.randn(d, device=dev) * 0.1
    for _ in range(3):
        ops.dcn_cross_fwd(xs[0], xs[1], W, b)
torch.cuda.synchronize()
print("ok")
