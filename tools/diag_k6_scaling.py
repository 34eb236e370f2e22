"""K6 fp16 x 2 GEMM time against K, N and the stage width (fixed cost vs per-K-block cost):
    python tools/diag_k6_scaling.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import _lib, ops
dev = torch.device("cuda:0")
lib = _lib.load()


def timeit(fn, n=20):
    """GPU time per call: n calls captured in one CUDA graph (no host enqueue time between the launches)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n) * 1e3


M = int(os.environ.get("B", 16384))
for bk in (32, 64):
    lib.ptrec_tc_set_bk(bk)
    for N in (256, 400, 512):
        row = []
        for K in (32, 128, 432, 864, 1728):
            x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
            hx, _, _, sx = ops.tc_split2h(x)
            hw, _, _, sw = ops.tc_split2h(w)
            t = timeit(lambda: ops.tc_gemm_split2h(hx, sx, hw, sw, K, bias=b, relu=True))
            row.append(f"K={K}: {t:6.1f} us")
        print(f"BK={bk} M={M} N={N}  " + "  ".join(row), flush=True)
lib.ptrec_tc_set_bk(32)
# wgrad form (MN-major, split-K): N x K = 400 x 432 over the batch
for Bt in (16384, 65536):
    g = torch.randn(Bt, 400, device=dev); x = torch.randn(Bt, 432, device=dev)
    hg, _, _, sg = ops.tc_split2h(g); hx, _, _, sx = ops.tc_split2h(x)
    for splits in (0, 8, 16, 32, 64):
        t = timeit(lambda: ops.tc_gemm_split2h_tn(hg, sg, 400, hx, sx, 432, splits=splits))
        print(f"wgrad B={Bt} splits={splits}: {t:6.1f} us", flush=True)
