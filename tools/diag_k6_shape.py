"""Where does a K6 GEMM differ from fp64?  Prints the error pattern (rows / columns) for one shape."""
import sys
import torch
from pytorchrec_b200 import ops

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (2048, 64, 256)
dev = torch.device("cuda:0")
torch.manual_seed(0)
a = (1e-5 * torch.randn(M, K, device=dev)) * (torch.rand(M, K, device=dev) > 0.5)
b = (torch.rand(N, K, device=dev) - 0.5) / 4
pa, _, _, sa = ops.tc_split2h(a)
pb, _, _, sb = ops.tc_split2h(b)
ref = a.double() @ b.double().t()
for name, fn in (("plain", lambda: ops.tc_gemm_split2h(pa, sa, pb, sb, K)),
                 ("absmax", lambda: ops.tc_gemm_split2h(pa, sa, pb, sb, K, want_absmax=True)[0]),
                 ("fused", lambda: ops.tc_gemm_split2h_fused(pa, sa, pb, sb, K)[0])):
    out = fn()
    err = (out.double() - ref).abs()
    bad = err > 1e-5 * ref.abs().max()
    print(name, "max err", err.max().item(), "ref max", ref.abs().max().item(), "bad", int(bad.sum()))
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("  bad rows", rows[:20].tolist(), "... n =", rows.numel())
        print("  bad cols", cols[:40].tolist(), "... n =", cols.numel())
