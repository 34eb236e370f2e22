"""Per-kernel time table of one eager train step of a BASELINE config on one B200, from the torch profiler (CUPTI):
    python tools/profile_step.py deepfm|dcn|din [out_tag]
Durations under the profiler are for the SHARE of each kernel in the step, not bench numbers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from pytorchrec_b200.data import amazon_batch, amazon_columns, criteo_batch, criteo_columns
from pytorchrec_b200.metric import LogLoss
from pytorchrec_b200.model import DCN, DIN, DeepFM
from pytorchrec_b200.optim import SparseAdagrad

which = sys.argv[1] if len(sys.argv) > 1 else "deepfm"
tag = sys.argv[2] if len(sys.argv) > 2 else which
dev = torch.device("cuda:0")
steps = 10
if which == "deepfm":
    sparse, dense, label = criteo_columns(26, 13, 1_000_000)
    model = DeepFM(sparse, dense, label, 16, [400, 400, 400], random_seed=1, table_device=dev)
    mk = lambda i: criteo_batch(16384, 26, 13, 1_000_000, seed=i)
elif which == "dcn":
    sparse, dense, label = criteo_columns(26, 13, 1_000_000)
    model = DCN(sparse, dense, label, 32, 3, [1024, 1024, 1024], random_seed=1, table_device=dev)
    mk = lambda i: criteo_batch(32768, 26, 13, 1_000_000, seed=i)
else:
    model = DIN(*amazon_columns(100), emb_size=16, layers=[200, 80], random_seed=1, table_device=dev)
    mk = lambda i: amazon_batch(8192, 100, seed=i)
model.compile(SparseAdagrad(params=model.get_parameters(), lr=0.01), torch.nn.BCEWithLogitsLoss(), [LogLoss()], dev)
batches = [{k: v.to(dev) for k, v in mk(i).items()} for i in range(4)]
for i in range(5):
    model.train_step(batches[i % 4])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        model.train_step(batches[i % 4])
    torch.cuda.synchronize()
evs = [e for e in prof.key_averages() if e.device_time_total > 0]
evs.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in evs)
ours = sum(e.device_time_total for e in evs if "ptrec::" in e.key)
out = [f"{which}: sum of kernel time per eager step {tot / steps:.1f} us in {sum(e.count for e in evs) / steps:.0f} launches; "
       f"libptrec kernels {ours / steps:.1f} us ({100 * ours / tot:.1f} %)",
       f"{'us/step':>9} {'%':>5} {'calls/step':>10} {'us/call':>8}  kernel"]
for e in evs[:60]:
    out.append(f"{e.device_time_total / steps:9.1f} {100 * e.device_time_total / tot:5.1f} {e.count / steps:10.1f} "
               f"{e.device_time_total / e.count:8.1f}  {e.key[:120]}")
txt = "\n".join(out)
print(txt, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/step_kernels_{tag}.txt", "w").write(txt + "\n")
