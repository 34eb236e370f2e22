"""Minimal driver for ncu captures of K5 (DCN-v2 cross layer GEMMs on the CTA-pair kernel) at the cfg3 shape
(B 32768, d 848): forward, input gradient, weight gradient, three times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import ops
dev = torch.device("cuda:0")
B, d = int(os.environ.get("B", 32768)), int(os.environ.get("D", 848))
g = torch.Generator(device=dev).manual_seed(0)
xs = [(torch.randn(B, d, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(4)]
W = (torch.randn(d, d, device=dev, generator=g) / d ** 0.5).to(torch.bfloat16)
bias = torch.randn(d, device=dev, generator=g) * 0.1
for it in range(3):
    out, u = ops.dcn_cross_fwd(xs[0], xs[1], W, bias)
    gx, prev = ops.dcn_cross_dgrad(xs[2], W, xs[3], xs[1])
    gw = ops.dcn_cross_wgrad(xs[2], xs[0])
torch.cuda.synchronize()
print("ok", float(out.float().abs().max()), float(gw.abs().max()))
