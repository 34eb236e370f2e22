"""K5 kernel timings at the cfg3 shape (B 32768, d 848) on one B200: forward / input gradient / weight gradient on the
CTA-pair GEMM and on the single-CTA kernel of round 1, CUDA events around 20 launches over rotating operands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import ops

dev = torch.device("cuda:0")
B, d = 32768, 848
g = torch.Generator(device=dev).manual_seed(0)
xs = [(torch.randn(B, d, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(6)]
W = (torch.randn(d, d, device=dev, generator=g) / d ** 0.5).to(torch.bfloat16)
Wt = W.t().contiguous()
bias = torch.randn(d, device=dev, generator=g) * 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


fl = 2.0 * B * d * d
for pair in (True, False):
    ops.set_dcn_2sm(pair)
    t_f = timeit(lambda i: ops.dcn_cross_fwd(xs[i % 6], xs[(i + 1) % 6], W, bias))
    t_d = timeit(lambda i: ops.dcn_cross_dgrad(xs[i % 6], Wt, xs[(i + 1) % 6], xs[(i + 2) % 6]))
    t_w = timeit(lambda i: ops.dcn_cross_wgrad(xs[i % 6], xs[(i + 1) % 6]))
    print("%s: fwd %.1f us (%.0f TFLOP/s)  dgrad %.1f us (%.0f)  wgrad %.1f us (%.0f)" % (
        "CTA pair 256x256" if pair else "single CTA 128x128", t_f, fl / t_f / 1e6, t_d, fl / t_d / 1e6, t_w, fl / t_w / 1e6))
ops.set_dcn_2sm(True)
a = torch.randn(B, d, device=dev).to(torch.bfloat16)
t_c = timeit(lambda i: torch.nn.functional.linear(xs[i % 6], W))
print("cuBLAS bf16 linear alone (no bias, no cross arithmetic): %.1f us (%.0f TFLOP/s)" % (t_c, fl / t_c / 1e6))
