"""torchrun worker (NCCL, >= 2 GPUs): ShardedDeepFM vs the unsharded fused DeepFM on the concatenated batch.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_sharded_worker.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_sharding  # noqa: E402
from pytorchrec_b200 import ops  # noqa: E402
from pytorchrec_b200.distributed import ShardedDeepFM  # noqa: E402
from pytorchrec_b200.distributed.sharded import list_capacity  # noqa: E402
from pytorchrec_b200.feature_column import CategoricalColumnWithIdentity as Col, NumericColumn  # noqa: E402
from pytorchrec_b200.metric import LogLoss  # noqa: E402
from pytorchrec_b200.model import DeepFM  # noqa: E402
from pytorchrec_b200.optim import SparseAdagrad  # noqa: E402


def full_capacity_check(B, world):
    from pytorchrec_b200.distributed.sharded import owner_share
    return list_capacity(B, world, 1.25, owner_share([103, 57, 1000, 64, 31, 5000], world))


def low_cardinality_and_overflow(rank, world, dev):
    """(1) a column with fewer categories than ranks and Zipf ids trains exactly like the unsharded
    model (lists sized from the cardinalities); (2) a list that does overflow is FATAL: the sticky word reaches the
    host through pinned memory and the next train_step raises, also when the step is a replayed CUDA graph."""
    import warnings
    F, nd, D, B = 4, 2, 16, 256
    rows = [1, world - 1 if world > 2 else 1, 40, 3000]
    scols = [Col(rows[f], f"C{f}") for f in range(F)]
    dcols = [NumericColumn(f"I{j}", 0.0, 1.0, 0.5, 0.25) for j in range(nd)]
    lab = Col(2, "label")
    full = DeepFM(scols, dcols, lab, D, [32, 16], random_seed=4)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        shard = ShardedDeepFM(scols, dcols, lab, D, [32, 16], random_seed=4)
    if world > 2:  # at 2 ranks a one-category column costs no more than any other (owner share 1 vs 1/2 * 2)
        assert any("fewer categories" in str(x.message) for x in w)
    assert shard.sharded.capacity(B) == (B + 15) // 16 * 16  # category_num 1: one owner takes the whole batch
    shard.load_full_state_dict(full.state_dict())
    full.compile(SparseAdagrad(full.get_parameters(), lr=0.05), torch.nn.BCEWithLogitsLoss(), [LogLoss()], dev)
    shard.compile(SparseAdagrad(shard.get_parameters(), lr=0.05), torch.nn.BCEWithLogitsLoss(), [LogLoss()], dev)
    shard.enable_cuda_graph(True, warmup=2)  # pull mode re-exchanges its pointers in step 2 (tables re-housed in step 1)

    def batch(step, hot=False):
        rng = np.random.default_rng(2000 + step)
        gb = {f"C{f}": torch.from_numpy(((rng.zipf(1.05, size=world * B) % rows[f]) if not hot else
                                        np.full(world * B, min(1, rows[f] - 1))).astype(np.int64)) for f in range(F)}
        gb.update({f"I{j}": torch.from_numpy(rng.random(world * B).astype(np.float32)) for j in range(nd)})
        gb["label"] = torch.from_numpy(rng.integers(0, 2, size=world * B))
        return gb, {k: v[rank * B:(rank + 1) * B] for k, v in gb.items()}

    for step in range(4):
        gb, lb = batch(step)
        lf = full.train_step(gb)["loss"].item()
        ls = shard.train_step(lb)["loss"]
        t = ls.detach().clone()
        dist.all_reduce(t)
        np.testing.assert_allclose(t.item() / world, lf, rtol=2e-5)
    shard.sharded.check_errors()
    gathered = shard.full_state_dict()
    for k, v in full.state_dict().items():
        np.testing.assert_allclose(gathered[k].numpy(), v.cpu().numpy(), rtol=0, atol=3e-4, err_msg=k)

    # (2) force an overflow: lists far below the fullest owner's share (the one-category column sends its whole batch
    # to one owner).  The sticky word reaches the host through pinned memory: a train_step within the next few steps
    # raises (eager or replayed graph alike), and check_errors raises too.
    shard.sharded.owner_share, shard.sharded.capacity_factor = 1.0 / world, 0.25
    assert shard.sharded.capacity(B) < B
    shard._graphed.entries.clear()  # new capacity = new buffers: re-warm and re-capture
    raised = False
    for step in range(4):
        try:
            shard.train_step(batch(10 + step)[1])
        except RuntimeError as e:
            assert "overflowed" in str(e), e
            raised = True
            break
        torch.cuda.synchronize()
    assert raised, "an overflowed exchange list must be fatal"
    flags = torch.tensor([1.0 if raised else 0.0], device=dev)
    dist.all_reduce(flags)               # every rank saw it (each packs the same one-category column)
    assert flags.item() == world
    try:
        shard.sharded.check_errors()
        raise AssertionError("check_errors must raise too")
    except RuntimeError as e:
        assert "overflowed" in str(e), e
    dist.barrier()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    F, nd, D, B = 6, 3, 16, 256
    rows = [103, 57, 1000, 64, 31, 5000]
    scols = [Col(rows[f], f"C{f}") for f in range(F)]
    dcols = [NumericColumn(f"I{j}", 0.0, 1.0, 0.5, 0.25) for j in range(nd)]
    lab = Col(2, "label")

    # the pack kernel against its CPU restatement (bit-exact)
    ids = torch.stack([torch.randint(0, rows[f], (B,), generator=torch.Generator().manual_seed(f + 10 * rank)) for f in range(F)])
    C = list_capacity(B, world)
    assert C <= full_capacity_check(B, world) <= (B + 15) // 16 * 16  # odd heights: one owner holds ceil(R / G) rows
    ovf = torch.zeros(1, dtype=torch.int32, device=dev)
    send_ids, ret_pos = ops.a2a_pack_by_owner(ids.to(dev), F, B, world, C, ovf)
    rs, rp, _ = ref_sharding.pack_by_owner_ref(ids, world, C)
    assert torch.equal(send_ids.cpu(), rs) and torch.equal(ret_pos.cpu(), rp) and ovf.item() == 0

    full = DeepFM(scols, dcols, lab, D, [32, 16], random_seed=3)
    shard = ShardedDeepFM(scols, dcols, lab, D, [32, 16], random_seed=3)
    sd = full.state_dict()
    shard.load_full_state_dict(sd)  # N4: an unsharded checkpoint loads into the sharded model (rows rank::G)
    full.compile(SparseAdagrad(full.get_parameters(), lr=0.05), torch.nn.BCEWithLogitsLoss(), [LogLoss()], dev)
    shard.compile(SparseAdagrad(shard.get_parameters(), lr=0.05), torch.nn.BCEWithLogitsLoss(), [LogLoss()], dev)
    if os.environ.get("PTREC_TEST_GRAPH", "1") == "1":
        shard.enable_cuda_graph(True, warmup=2)  # NCCL all-to-all + all-reduce captured in the step graph

    for step in range(4):
        rng = np.random.default_rng(1000 + step)  # same global batch on every rank
        gb = {f"C{f}": torch.from_numpy((rng.zipf(1.3, size=world * B) % rows[f]).astype(np.int64)) for f in range(F)}
        gb.update({f"I{j}": torch.from_numpy(rng.random(world * B).astype(np.float32)) for j in range(nd)})
        gb["label"] = torch.from_numpy(rng.integers(0, 2, size=world * B))
        lb = {k: v[rank * B:(rank + 1) * B] for k, v in gb.items()}
        with torch.no_grad():  # a live autograd graph would pin AccumulateGrad nodes to the default stream
            pf, _ = full.test_step(gb)
            ps, _ = shard.test_step(lb)
        np.testing.assert_allclose(ps.detach().cpu().numpy(), pf.detach().cpu().numpy()[rank * B:(rank + 1) * B],
                                   rtol=1e-5, atol=2e-5)
        full.train_step(gb)
        shard.train_step(lb)
    shard.sharded.check_errors()
    want = os.environ.get("PTREC_EXCHANGE") or {"1": "pull", "0": "a2a"}.get(os.environ.get("PTREC_PEER_GATHER", ""), "push")
    assert shard.sharded.exchange == want, (shard.sharded.exchange, want)
    fsd, ssd = full.state_dict(), shard.state_dict()
    inv = {0: "embeddings", 1: "first_order"}
    for k, v in ssd.items():
        if k.startswith("sharded.groups."):
            _, _, gk, f, _ = k.split(".")
            ref = fsd[f"{inv[int(gk)]}.{f}.weight"][rank::world]
        else:
            ref = fsd[k]
        a, b = v.cpu().numpy(), ref.cpu().numpy()
        tight = np.abs(a - b) <= 1e-5 * np.abs(b) + 1.5e-5
        assert tight.mean() >= 0.995, (k, tight.mean())
        np.testing.assert_allclose(a, b, rtol=0, atol=3e-4, err_msg=k)
    # N4: the gathered checkpoint is the unsharded model's state_dict (keys, shapes, values)
    gathered = shard.full_state_dict()
    assert set(gathered) == set(fsd), set(gathered) ^ set(fsd)
    for k, v in gathered.items():
        assert v.shape == fsd[k].shape, k
        np.testing.assert_allclose(v.numpy(), fsd[k].cpu().numpy(), rtol=0, atol=3e-4, err_msg=k)
    probe = DeepFM(scols, dcols, lab, D, [32, 16], random_seed=99)
    probe.load_state_dict(gathered)
    # N4, optimizer state: the gathered Adagrad sums are the unsharded optimizer's, and they load back row-wise
    ost = shard.full_optimizer_state_dict()
    fopt = full.compiled_optimizers
    assert ost["step"] == fopt._step_count_fused and ost["dense"] is not None
    for gname, mod in (("embeddings", full.embeddings), ("first_order", full.first_order)):
        for f, t in enumerate(mod):
            ref = fopt.table_state(t.weight)["sum"].cpu().numpy()
            got = ost["tables"][f"{gname}.{f}.weight"]["sum"].numpy()
            assert got.shape == ref.shape, (gname, f, got.shape, ref.shape)
            np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-7, err_msg=f"{gname}.{f}.sum")
    for t in shard.sharded.tables:  # wipe, reload, gather again: identical
        shard.compiled_optimizers.table_state(t.weight)["sum"].zero_()
    shard.load_full_optimizer_state_dict(ost)
    ost2 = shard.full_optimizer_state_dict()
    for k, d in ost["tables"].items():
        assert torch.equal(d["sum"], ost2["tables"][k]["sum"]), k
    dist.barrier()
    low_cardinality_and_overflow(rank, world, dev)
    if rank == 0:
        print("DIST_SHARDED_OK world=%d exchange=%s" % (world, shard.sharded.exchange), flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)  # graphs that captured NCCL kernels make destroy_process_group() hang


if __name__ == "__main__":
    main()
