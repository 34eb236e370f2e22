"""Minimal driver for ncu captures of the K6 fused-tower GEMMs at the cfg2 hidden-layer shape (B 16384, 400 -> 400):
per iteration one forward GEMM (bias + ReLU -> planes + bit mask) and one input-gradient GEMM (bit mask -> planes +
column sums), plus the plain fp32 form."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import ops
dev = torch.device("cuda:0")
B, K, N = int(os.environ.get("B", 16384)), int(os.environ.get("K", 400)), int(os.environ.get("N", 400))
x = torch.rand(B, K, device=dev)
w = 0.05 * torch.randn(N, K, device=dev)
b = 0.01 * torch.randn(N, device=dev)
g = 1e-5 * torch.randn(B, N, device=dev)
px, _, _, sx = ops.tc_split2h(x)
pw, pwt, _, sw = ops.tc_split2h(w, want_t=True)
pg, _, _, sg = ops.tc_split2h(g)
one = torch.tensor([2.0 ** 4], device=dev)
big = torch.tensor([2.0 ** 20], device=dev)
mx = torch.zeros(1, device=dev)
for it in range(3):
    _, py, mask, _ = ops.tc_gemm_split2h_fused(px, sx, pw, sw, K, bias=b, relu=True, want_out=False, out_scale=one,
                                               want_mask=True, max_out=mx)
    _, pgp, _, cs = ops.tc_gemm_split2h_fused(pg, sg, pwt, sw, N, want_out=False, out_scale=big, mask_in=mask,
                                              want_colsum=True, max_out=mx)
    y = ops.tc_gemm_split2h(px, sx, pw, sw, K, bias=b, relu=True)
torch.cuda.synchronize()
print("ok", float(y.abs().max()), float(cs.abs().max()))
