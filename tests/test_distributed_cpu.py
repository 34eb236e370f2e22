"""world_size-2 gloo tests of the multi-GPU host logic (CPU only): the row-wise shard plan
(pack -> all_to_all -> owner lookup -> all_to_all -> unpack) reproduces the unsharded lookup bit-exactly,
and the dense-gradient allreduce averages."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_sharding
from pytorchrec_b200.distributed import allreduce_dense_grads, shard_rows
from pytorchrec_b200.distributed.sharded import list_capacity


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        F, D, B = 4, 8, 96
        rows = [101, 37, 1000, 64]
        g = torch.Generator().manual_seed(5)
        full = [torch.randn(r, D, generator=g) for r in rows]                      # same on every rank
        local = [t[rank::world].contiguous() for t in full]                        # row-wise shard
        for f in range(F):
            assert local[f].shape[0] == shard_rows(rows[f], world, rank)
        ids = torch.stack([torch.randint(0, rows[f], (B,), generator=torch.Generator().manual_seed(100 * rank + f))
                           for f in range(F)])
        C = list_capacity(B, world)
        send_ids, ret_pos, longest = ref_sharding.pack_by_owner_ref(ids, world, C)
        assert longest <= C
        recv_ids = torch.empty_like(send_ids)
        dist.all_to_all_single(recv_ids, send_ids)
        rows_out = ref_sharding.owner_lookup_ref(recv_ids, local)
        recv_rows = torch.empty_like(rows_out)
        dist.all_to_all_single(recv_rows, rows_out)
        out = ref_sharding.unpack_ref(recv_rows, ret_pos)
        want = torch.stack([full[f][ids[f]] for f in range(F)], dim=1)
        assert torch.equal(out, want), "sharded lookup differs from the unsharded one"
        # dense allreduce: mean over ranks, one flat collective
        p = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))]
        p[0].grad = torch.full((3, 2), float(rank + 1))
        p[1].grad = torch.arange(5.0) * (rank + 1)
        allreduce_dense_grads(p)
        mean = sum(range(1, world + 1)) / world
        assert torch.allclose(p[0].grad, torch.full((3, 2), mean))
        assert torch.allclose(p[1].grad, torch.arange(5.0) * mean)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_shard_plan_and_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_rows_partition():
    for total in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            assert sum(shard_rows(total, world, r) for r in range(world)) == total


def test_capacity_covers_uniform_hashing():
    rng = np.random.default_rng(0)
    for B, G in ((16384, 8), (65536, 8), (4096, 2), (100, 4)):
        C = list_capacity(B, G)
        ids = rng.integers(0, 10**6, size=(50, B))
        worst = max(np.bincount(row % G, minlength=G).max() for row in ids)
        assert worst <= C and C % 16 == 0
