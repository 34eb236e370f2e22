"""Minimal driver for ncu captures of the K4 kernels (cfg4 shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import ops
dev = torch.device("cuda:0")
B, L, DQ, H1, H2 = 8192, 100, 32, 80, 40
g = torch.Generator(device=dev).manual_seed(0)
P = [torch.randn(H1, 4 * DQ, device=dev, generator=g) * 0.1, torch.randn(H1, device=dev, generator=g) * 0.1,
     torch.randn(H2, H1, device=dev, generator=g) * 0.1, torch.randn(H2, device=dev, generator=g) * 0.1,
     torch.randn(1, H2, device=dev, generator=g) * 0.1, torch.randn(1, device=dev, generator=g) * 0.1]
seq = torch.randn(B, 1 + L, DQ, device=dev, generator=g)
lens = torch.randint(1, L + 1, (B,), device=dev, generator=g).int()
go = torch.randn(B, DQ, device=dev, generator=g)
for _ in range(3):
    ops.din_attn_pool_fwd(seq[:, 0], seq[:, 1:], lens, P)
    ops.din_attn_pool_bwd(seq[:, 0], seq[:, 1:], lens, P, go)
torch.cuda.synchronize()
print("ok")
