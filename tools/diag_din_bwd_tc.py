"""K4 backward: tensor-core build (ptrec_set_din_tc mode 3) against the fp32 CUDA-core build (mode 0), relative to each tensor's maximum."""
import sys, os
sys.path.insert(0, "/root/repo")
import torch
from pytorchrec_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
B, L, DQ, H1, H2 = 64, 100, 32, 80, 40
g = torch.Generator(device=dev).manual_seed(0)
P = [torch.randn(H1, 4 * DQ, device=dev, generator=g) * 0.2, torch.randn(H1, device=dev, generator=g) * 0.1,
     torch.randn(H2, H1, device=dev, generator=g) * 0.2, torch.randn(H2, device=dev, generator=g) * 0.1,
     torch.randn(1, H2, device=dev, generator=g) * 0.2, torch.randn(1, device=dev, generator=g) * 0.1]
seq = torch.randn(B, 1 + L, DQ, device=dev, generator=g)
lens = torch.randint(1, L + 1, (B,), device=dev, generator=g).int()
go = torch.randn(B, DQ, device=dev, generator=g)
q, keys = seq[:, 0], seq[:, 1:]
res = {}
for mode in (0, 3):
    lib.ptrec_set_din_tc(mode)
    try:
        gq, gk, gp = ops.din_attn_pool_bwd(q, keys, lens, P, go)
        torch.cuda.synchronize()
        res[mode] = (gq, gk, gp)
    except Exception as e:
        print("mode", mode, "failed:", e)
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
for mode in (3,):
    if mode in res:
        r0, r = res[0], res[mode]
        print("mode", mode, "gq %.2e gk %.2e" % (rel(r[0], r0[0]), rel(r[1], r0[1])),
              " ".join("%s %.2e" % (n, rel(a, b)) for n, a, b in zip(["W1", "b1", "W2", "b2", "W3", "b3"], r[2], r0[2])))
