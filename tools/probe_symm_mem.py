"""2-GPU probe: torch symmetric memory gives peer pointers that the gather kernel can read over NVLink."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
from pytorchrec_b200 import ops
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}"); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
R, D = 1_000_000, 16
t = symm.empty(R, 2 * D, dtype=torch.float32, device=dev)
t.copy_(torch.arange(R, device=dev, dtype=torch.float32).unsqueeze(1).expand(R, 2 * D) + 0.5 * rank)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hdl.has_multicast_support if hasattr(hdl, "has_multicast_support") else None, flush=True)
dist.barrier(); torch.cuda.synchronize()
peer_rank = (rank + 1) % world
peer = hdl.get_buffer(peer_rank, (R, 2 * D), torch.float32)
assert abs(peer[5, 0].item() - (5 + 0.5 * peer_rank)) < 1e-6
# K1 reading the PEER table through its raw pointer (row stride 2*D: interleaved layout)
B = 16384
ts = ops.TableSet()
ts.ptrs = torch.tensor([hdl.buffer_ptrs[peer_rank]], dtype=torch.int64).to(dev)
ts.rows = torch.tensor([R], dtype=torch.int64).to(dev)
ts.max_rows, ts.row_stride = R, 2 * D
lay = ops.FeatureLayout([dict(table=0, bag_len=1)], D, 1)
ids = torch.randint(0, R, (B,), device=dev)
out, _ = ops.gather_pool_fwd(ts, lay, ids, None, B)
torch.cuda.synchronize()
want = ids.float().unsqueeze(1).expand(B, D) + 0.5 * peer_rank
assert torch.equal(out, want), (out[:2], want[:2])
# bandwidth of a peer gather: 26 "fields" worth of lookups
n = 26 * B
ids = torch.randint(0, R, (n,), device=dev)
for _ in range(3):
    ops.gather_pool_fwd(ts, lay, ids, None, n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.gather_pool_fwd(ts, lay, ids, None, n)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
print(rank, f"peer gather of {n} x 64 B rows: {us:.1f} us  ({n * 64 / us / 1e3:.0f} GB/s over NVLink)", flush=True)
dist.barrier()
if rank == 0:
    print("SYMM_OK", flush=True)
torch.cuda.synchronize()
os._exit(0)
