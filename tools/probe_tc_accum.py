"""How exact is the tcgen05 fp32 accumulation?  bf16 x bf16 products are exact in fp32, so any error against an fp64
reference of the SAME bf16 inputs is accumulation error.  Decides whether a 3-way bf16 split can carry fp32 GEMMs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import ops
dev = torch.device("cuda:0")
for K in (448, 2560, 16384):
    d = 256
    g = torch.randn(K, d, device=dev).bfloat16()
    x = torch.randn(K, d, device=dev).bfloat16()
    out = ops.dcn_cross_wgrad(g, x)                      # g^T x, fp32 accumulate in TMEM (split-K partials in fp32)
    ref = g.double().t() @ x.double()
    f32 = (g.float().t() @ x.float()).double()           # cuBLAS fp32 SIMT on the same values
    scale = (g.double().abs().t() @ x.double().abs())    # sum of |terms|
    e_tc = ((out.double() - ref).abs() / scale).max().item()
    e_f32 = ((f32 - ref).abs() / scale).max().item()
    r_tc = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"K={K}: tcgen05 max err/sum|terms| {e_tc:.2e} (cuBLAS fp32 {e_f32:.2e}); max abs err / max|ref| {r_tc:.2e}")
