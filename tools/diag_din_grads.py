"""Where does K4's backward differ from an fp64 reference?  Per parameter block: error of the CUDA gradient and of a
plain fp32 torch (GPU) autograd gradient against fp64 autograd, relative to the block's RMS gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import ops

dev = torch.device("cuda:0")
B, L, DQ, H1, H2 = 8192, 100, 32, 80, 40
g = torch.Generator(device=dev).manual_seed(0)
std = 0.01
P = [torch.randn(H1, 4 * DQ, device=dev, generator=g) * std, torch.randn(H1, device=dev, generator=g) * std,
     torch.randn(H2, H1, device=dev, generator=g) * std, torch.randn(H2, device=dev, generator=g) * std,
     torch.randn(1, H2, device=dev, generator=g) * std, torch.randn(1, device=dev, generator=g) * std]
seq = torch.randn(B, 1 + L, DQ, device=dev, generator=g) * std
lens = torch.randint(1, L + 1, (B,), device=dev, generator=g).int()
go = torch.randn(B, DQ, device=dev, generator=g) * 1e-3


def ref(dtype):
    p = [x.to(dtype).requires_grad_(True) for x in P]
    s = seq.to(dtype).requires_grad_(True)
    q, k = s[:, 0], s[:, 1:]
    qe = q.unsqueeze(1).expand(B, L, DQ)
    z = torch.cat([qe, k, qe - k, qe * k], -1)
    F = torch.nn.functional
    a = F.linear(torch.relu(F.linear(torch.relu(F.linear(z, p[0], p[1])), p[2], p[3])), p[4], p[5]).squeeze(-1)
    a = a * (torch.arange(L, device=dev).unsqueeze(0) < lens.unsqueeze(1)).to(dtype)
    out = (a.unsqueeze(-1) * k).sum(1)
    out.backward(go.to(dtype))
    return out.detach(), [x.grad for x in p], s.grad


o64, g64, s64 = ref(torch.float64)
o32, g32, s32 = ref(torch.float32)
q, keys = seq[:, 0], seq[:, 1:]
oc, _ = ops.din_attn_pool_fwd(q, keys, lens, P)
gq, gk, gp = ops.din_attn_pool_bwd(q, keys, lens, P, go)
sc = torch.cat([gq.unsqueeze(1), gk], 1)


def rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().pow(2).mean().sqrt()).item()


print(f"forward: CUDA {rel(oc, o64):.2e}  torch fp32 {rel(o32, o64):.2e}")
names = ["W1", "b1", "W2", "b2", "W3", "b3"]
for n, c, t32, t64 in zip(names, gp, g32, g64):
    print(f"grad {n}: CUDA {rel(c.view_as(t64), t64):.2e}  torch fp32 {rel(t32, t64):.2e}")
    if n == "W1":
        for bi, bn in enumerate(["W1q", "W1k", "W1d", "W1p"]):
            sl = slice(bi * DQ, (bi + 1) * DQ)
            print(f"   {bn}: CUDA {rel(c.view_as(t64)[:, sl], t64[:, sl]):.2e}  torch fp32 {rel(t32[:, sl], t64[:, sl]):.2e}  rms {t64[:, sl].pow(2).mean().sqrt().item():.2e}")
print(f"grad seq (q | keys): CUDA {rel(sc, s64):.2e}  torch fp32 {rel(s32, s64):.2e}")
