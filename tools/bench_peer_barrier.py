"""Latency of ops.PeerSync.barrier (csrc/peer_sync.cu) against a 1-element NCCL all_reduce, back to back on one stream:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29571 tools/bench_peer_barrier.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pytorchrec_b200 import ops

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}"); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
sync = ops.PeerSync(None, dev)
buf = torch.zeros(1, device=dev)


def timeit(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


t_b = timeit(lambda: sync.barrier(ops.PeerSync.FENCE))
t_n = timeit(lambda: dist.all_reduce(buf))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(50):
        sync.barrier(ops.PeerSync.FENCE)
t_g = timeit(lambda: g.replay(), 20) / 50
if rank == 0:
    txt = (f"world={world}: peer_barrier_kernel {t_b:.2f} us per call (eager launches), {t_g:.2f} us inside a CUDA graph; "
           f"1-element ncclAllReduce {t_n:.2f} us per call")
    print(txt, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    open(f"gpurun_out/peer_barrier_n{world}.txt", "w").write(txt + "\n")
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
