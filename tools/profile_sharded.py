"""Per-kernel time table of the sharded DeepFM step (cfg2 shape) on rank 0, from the torch profiler (CUPTI), eager mode.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29577 tools/profile_sharded.py
The absolute step time under the profiler is not a bench number; the per-kernel durations are what this is for."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from pytorchrec_b200.data import criteo_batch, criteo_columns
from pytorchrec_b200.metric import LogLoss
from pytorchrec_b200.optim import SparseAdagrad

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device(f"cuda:{local}"); torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rows, B, steps = 1_000_000, 16384, 10
sparse, dense, label = criteo_columns(26, 13, rows)
if world == 1:
    from pytorchrec_b200.model import DeepFM as M
else:
    from pytorchrec_b200.distributed import ShardedDeepFM as M
model = M(sparse, dense, label, 16, [400, 400, 400], random_seed=2020, table_device=dev)
model.compile(SparseAdagrad(params=model.get_parameters(), lr=0.01), torch.nn.BCEWithLogitsLoss(), [LogLoss()], dev)
batches = [{k: v.to(dev) for k, v in criteo_batch(B, 26, 13, rows, seed=1000 * (rank + 1) + i).items()} for i in range(4)]
for i in range(5):
    model.train_step(batches[i % 4])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        model.train_step(batches[i % 4])
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.key_averages() if e.device_time_total > 0]
    evs.sort(key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in evs)
    out = [f"world={world} peer={os.environ.get('PTREC_PEER_GATHER', '1')}  sum of kernel time per step: {tot / steps:.1f} us",
           f"{'us/step':>9} {'calls/step':>10} {'us/call':>8}  kernel"]
    for e in evs[:45]:
        out.append(f"{e.device_time_total / steps:9.1f} {e.count / steps:10.1f} {e.device_time_total / e.count:8.1f}  {e.key[:110]}")
    txt = "\n".join(out)
    print(txt, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    open(f"gpurun_out/kernels_sharded_n{world}_peer{os.environ.get('PTREC_PEER_GATHER', '1')}.txt", "w").write(txt + "\n")
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
os._exit(0)
