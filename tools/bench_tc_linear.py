"""K6 timing at the cfg2 DNN shapes (B 16384, 429/400 -> 400): split-3 tcgen05 GEMMs vs the fp32 cuBLAS path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
B = int(os.environ.get("B", 16384))
from pytorchrec_b200 import _lib
_lib.load().ptrec_tc_set_2sm(int(os.environ.get("TWO_SM", "1")))
_lib.load().ptrec_tc_set_bk(int(os.environ.get("BK", "32")))
for K, N in ((429, 400), (400, 400), (1024, 1024)):
    x = torch.randn(B, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    g = torch.randn(B, N, device=dev)
    px, pxt, _ = ops.tc_split3(x, want_t=True)
    pw, pwt, _ = ops.tc_split3(w, want_t=True)
    pg, pgt, _ = ops.tc_split3(g, want_t=True)
    fl = 2.0 * B * N * K
    t = {}
    t["split3(x) rm+t"] = timeit(lambda: ops.tc_split3(x, want_t=True))
    t["split3(g) relu+colsum rm+t"] = timeit(lambda: ops.tc_split3(g, relu_ref=g, want_t=True, want_colsum=True))
    t["fwd  tc"] = timeit(lambda: ops.tc_gemm_split3(px, pw, K, bias=b, relu=True))
    t["fwd  cublas fp32 linear+relu"] = timeit(lambda: torch.relu(torch.nn.functional.linear(x, w, b)))
    t["dgrad tc"] = timeit(lambda: ops.tc_gemm_split3(pg, pwt, N))
    t["dgrad cublas fp32"] = timeit(lambda: g @ w)
    t["wgrad tc"] = timeit(lambda: ops.tc_gemm_split3(pgt, pxt, B, splits=0))
    t["wgrad cublas fp32"] = timeit(lambda: g.t() @ x)
    # fp16 x 2 operand format (3 MMAs per product, two planes)
    hx, _, _, sx = ops.tc_split2h(x)
    hw, hwt, _, sw = ops.tc_split2h(w, want_t=True)
    hg, _, _, sg = ops.tc_split2h(g)
    t["split2h(x) absmax+rm"] = timeit(lambda: ops.tc_split2h(x))
    t["split2h(g) relu+colsum rm"] = timeit(lambda: ops.tc_split2h(g, relu_ref=g, want_colsum=True))
    t["fwd  tc h2"] = timeit(lambda: ops.tc_gemm_split2h(hx, sx, hw, sw, K, bias=b, relu=True))
    t["dgrad tc h2"] = timeit(lambda: ops.tc_gemm_split2h(hg, sg, hwt, sw, N))
    t["wgrad tc h2 (tn)"] = timeit(lambda: ops.tc_gemm_split2h_tn(hg, sg, N, hx, sx, K))
    t["wgrad tc (tn)"] = timeit(lambda: ops.tc_gemm_split3_tn(pg, N, px, K))
    ref = torch.relu(x.double() @ w.double().t() + b.double())
    den = x.double().abs() @ w.double().abs().t() + b.double().abs()
    e3 = ((ops.tc_gemm_split3(px, pw, K, bias=b, relu=True).double() - ref).abs() / den).max().item()
    e2 = ((ops.tc_gemm_split2h(hx, sx, hw, sw, K, bias=b, relu=True).double() - ref).abs() / den).max().item()
    ef = ((torch.relu(torch.nn.functional.linear(x, w, b)).double() - ref).abs() / den).max().item()
    print(f"B={B} K={K} N={N}  ({fl / 1e9:.1f} GFLOP fp32-equivalent per GEMM)  max err / sum|a||b|: "
          f"bf16x3 {e3:.2e}  fp16x2 {e2:.2e}  cublas fp32 {ef:.2e}")
    for k, v in t.items():
        extra = f"  {fl / v / 1e6:7.1f} TFLOP/s fp32-equiv" if "tc" in k or "cublas" in k else ""
        print(f"  {k:34s} {v:8.1f} us{extra}")
