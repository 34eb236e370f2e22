"""Minimal driver for ncu captures of K6 in its default (fp16 x 2) operand format at the cfg2 first-layer shape:
per iteration one |x| maximum pass, one split and one forward GEMM (B 16384, 429 -> 400)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import ops
dev = torch.device("cuda:0")
B, K, N = int(os.environ.get("B", 16384)), int(os.environ.get("K", 429)), int(os.environ.get("N", 400))
xs = [torch.rand(B, K, device=dev) for _ in range(4)]
w = 0.01 * torch.randn(N, K, device=dev)
b = 0.01 * torch.randn(N, device=dev)
pw, _, _, sw = ops.tc_split2h(w)
for it in range(4):
    px, _, _, sx = ops.tc_split2h(xs[it])
    y = ops.tc_gemm_split2h(px, sx, pw, sw, K, bias=b, relu=True)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
