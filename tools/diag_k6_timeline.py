"""Per-CTA milestones of one K6 CTA-pair GEMM launch (diagnostic build: make -C pytorchrec_b200/csrc EXTRA=-DPTREC_K6_TIMELINE
after touching tc_linear.cu).  Slots: 0 entry, per tile it: 1+4it before the accumulator wait, 2+4it accumulator ready,
3+4it TMEM handed back, 4+4it tile stored; 13 loop done, 14 stores landed."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import _lib, ops
dev = torch.device("cuda:0")
lib = _lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
B, K, N = int(os.environ.get("B", 16384)), int(os.environ.get("K", 429)), int(os.environ.get("N", 400))
x = torch.rand(B, K, device=dev); w = 0.05 * torch.randn(N, K, device=dev); b = 0.01 * torch.randn(N, device=dev)
px, _, _, sx = ops.tc_split2h(x); pw, _, _, sw = ops.tc_split2h(w)
buf = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
for _ in range(3):
    ops.tc_gemm_split2h(px, sx, pw, sw, K, bias=b, relu=True)
torch.cuda.synchronize()
assert raw.ptrec_debug_k6_timeline(ctypes.c_void_p(buf.data_ptr())) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.tc_gemm_split2h(px, sx, pw, sw, K, bias=b, relu=True); e1.record()
torch.cuda.synchronize()
raw.ptrec_debug_k6_timeline(ctypes.c_void_p(0))
t = buf.cpu().view(148, 32).double()
t0 = t[:, 0][t[:, 0] > 0].min()
names = {0: "entry", 1: "t0 wait", 2: "t0 acc ready", 16: "  pass0 drained", 17: "  chunk0 out", 24: "   c1 computed", 25: "   c1 box free", 26: "   c1 staged", 27: "   c1 fenced", 18: "  chunk1 out",
         20: "  pass1 drained", 3: "t0 tmem free", 21: "  chunk2 out", 22: "  chunk3 out", 4: "t0 stored", 5: "t1 wait",
         6: "t1 acc ready", 7: "t1 tmem free", 8: "t1 stored", 30: "loop done", 31: "stores landed"}
print(f"launch {e0.elapsed_time(e1) * 1e3:.1f} us (event); SM cycles after each CTA's own entry: min / median / max over CTAs")
for s, n in names.items():
    ok = (t[:, s] > 0) & (t[:, 0] > 0)
    if ok.any():
        v = (t[ok, s] - t[ok, 0])
        print(f"  {n:16s} n={int(ok.sum()):3d}  {v.min():8.0f} {v.median():8.0f} {v.max():8.0f}")
