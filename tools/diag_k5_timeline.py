"""Per-CTA milestones of one K5 launch on the CTA-pair GEMM (diagnostic library libptrec_b200_tl.so: tc_linear.cu built
with -DPTREC_K6_TIMELINE).  Slots per tile it: 1+4it before the accumulator wait, 2+4it accumulator ready, 3+4it TMEM
handed back, 4+4it tile stored; 30 loop done, 31 stores landed."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libptrec_b200_tl.so")
from pytorchrec_b200 import ops
dev = torch.device("cuda:0")
lib = _lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
B, d = int(os.environ.get("B", 32768)), int(os.environ.get("D", 848))
which = os.environ.get("WHICH", "fwd")
g = torch.Generator(device=dev).manual_seed(0)
xs = [(torch.randn(B, d, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(4)]
W = (torch.randn(d, d, device=dev, generator=g) / d ** 0.5).to(torch.bfloat16)
bias = torch.randn(d, device=dev, generator=g) * 0.1
fn = {"fwd": lambda: ops.dcn_cross_fwd(xs[0], xs[1], W, bias), "dgrad": lambda: ops.dcn_cross_dgrad(xs[0], W, xs[1], xs[2]),
      "wgrad": lambda: ops.dcn_cross_wgrad(xs[0], xs[1])}[which]
buf = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
for _ in range(3):
    fn()
torch.cuda.synchronize()
assert raw.ptrec_debug_k6_timeline(ctypes.c_void_p(buf.data_ptr())) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record()
torch.cuda.synchronize()
raw.ptrec_debug_k6_timeline(ctypes.c_void_p(0))
t = buf.cpu().view(148, 32).double()
print(f"{which}: launch {e0.elapsed_time(e1) * 1e3:.1f} us (event); SM cycles after each CTA's own entry: min / median / max over CTAs")
names = {0: "entry"}
for it in range(7):
    names.update({1 + 4 * it: f"t{it} wait", 2 + 4 * it: f"t{it} acc ready", 3 + 4 * it: f"t{it} tmem free", 4 + 4 * it: f"t{it} stored"})
names.update({30: "loop done", 31: "stores landed"})
for s, n in names.items():
    ok = (t[:, s] > 0) & (t[:, 0] > 0)
    if ok.any():
        v = (t[ok, s] - t[ok, 0])
        print(f"  {n:16s} n={int(ok.sum()):3d}  {v.min():8.0f} {v.median():8.0f} {v.max():8.0f}")
