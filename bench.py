#!/usr/bin/env python
"""bench.py — train samples/sec of the DeepFM hot path on B200 (BASELINE.json metric).

Own arm (default): cfg2 = DeepFM, Criteo-shaped synthetic, 26 tables x 1e6 rows x dim 16 (+26 first-order
tables dim 1), 13 dense, DNN 400-400-400, batch 16384 per GPU, fp32, fused sparse Adagrad.  A "step" is
one `IModel.train_step` (forward, BCE loss, backward with the fused sort/dedup/scatter/update, dense
optimizer step).  `value` = samples/s with batches resident in HBM; `e2e` = the same step fed from pinned
HOST batches (H2D inside the timed region) with the loss read back every step.

The same JSON line also carries (rank 0): `cfg5` — the north star's row-wise sharded DeepFM config (26 tables x
5e7*N/8 rows x dim 64, batch 65536 per GPU; weak-scaling series of SURVEY.md 8d) measured at this N right after
the cfg2 measurement, so that the driver's N = 1, 2, 4, 8 runs carry the cfg5 curve; and at N = 1 `models` — train
step times of cfg3 (DCN-v2) and cfg4 (DIN) with the rooflines of their interaction kernels.

`--impl reference`: the reference's CPU path (oracle port of the reference idiom: nn.Embedding per column,
dense autograd gradients, dense torch.optim.Adagrad, driven by the reference's five-line train_step) timed
on this box's host cores: same workload / metric / unit; its line says what it ran (one CPU process, the steps
and warm-up it really executed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: everything libraries print (NCCL banner, warnings) is sent to stderr by
# pointing fd 1 at fd 2 for the whole run; the result line is written to the saved descriptor at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

CFG = dict(n_sparse=26, n_dense=13, rows=1_000_000, dim=16, batch=16384, layers=[400, 400, 400])


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--id-dist", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--batch", type=int, default=CFG["batch"])
    ap.add_argument("--rows", type=int, default=CFG["rows"])
    ap.add_argument("--dim", type=int, default=CFG["dim"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager train_step instead of the whole-step CUDA graph")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg5"],
                    help="cfg2 (default, the bench line): 26x1e6xD16, B 16384/GPU.  cfg5: 26 x (5e7*G/8) rows x D64, "
                         "B 65536/GPU, row-wise sharded (weak-scaling series of SURVEY 8d)")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cfg5", action="store_true", help="skip the cfg5 sub-record")
    ap.add_argument("--no-models", action="store_true", help="skip the cfg3 / cfg4 model records (N = 1)")
    ap.add_argument("--sub-steps", type=int, default=20, help="timed steps of the cfg5 / cfg3 / cfg4 sub-records")
    ap.add_argument("--per-key-h2d", action="store_true",
                    help="e2e: host batches as separately pinned tensors (one H2D copy per key) instead of one packed pinned buffer")
    ap.add_argument("--sync-loss", action="store_true",
                    help="e2e: report the variant that calls loss.item() (a host synchronisation) after every step; by "
                         "default step k's loss is copied to a pinned word every step and read by the host once step "
                         "k+1 is enqueued (both variants are measured and printed)")
    return ap.parse_args()


def config_dict(a, world):
    from pytorchrec_b200 import ops
    return {"dnn_gemm_operands": ops.tc_mode(),  # K6: fp32 Linear operands as bf16x3 or fp16x2 planes
            "workload": "%s DeepFM Criteo-shaped synthetic: 26 tables x %d rows x dim %d (+26 first-order dim 1), "
                        "13 dense, DNN 400-400-400, batch %d per process, fp32, Adagrad" % (a.workload, a.rows, a.dim, a.batch),
            "table_update": "fused sparse Adagrad (rows touched by the batch)",
            "global_batch": a.batch * world, "id_dist": a.id_dist, "cuda_graph": not a.no_graph,
            "parallelism": "single GPU" if world == 1 else (
                f"row-wise sharded tables x{world} ("
                + {"push": "owners push rows / requesters push gradients over NVLink inside the kernels, barrier kernels, "
                           "dense all-reduce fused into the optimizer launch: no NCCL collective in the step",
                   "pull": "NVLink peer-memory gather / gradient push inside the kernels",
                   "a2a": "NCCL all-to-all"}[getattr(a, "exchange", "push")] + ")"),
            "l2": "tables %.2f GB >> 126 MB L2; a different random id batch every step" %
                  (26 * a.rows * (a.dim + 1) * 4 / 1e9)}


# ------------------------------------------------------------------------------------------------ CPU arm
def build_cpu_model(a):
    import torch
    from oracle import ref_models
    from pytorchrec_b200.data import criteo_columns
    sparse, dense, label = criteo_columns(CFG["n_sparse"], CFG["n_dense"], a.rows)
    model = ref_models.DeepFMRef(2020, sparse, dense, label, a.dim, CFG["layers"])
    model.compile(torch.optim.Adagrad(model.get_parameters(), lr=0.01), torch.nn.BCEWithLogitsLoss())
    return model


def time_cpu(a, steps, warmup):
    import torch
    from pytorchrec_b200.data import criteo_batch
    torch.set_num_threads(os.cpu_count() or 1)
    model = build_cpu_model(a)
    batches = [criteo_batch(a.batch, CFG["n_sparse"], CFG["n_dense"], a.rows, seed=900 + i, dist=a.id_dist)
               for i in range(min(4, steps + warmup))]
    for i in range(warmup):
        model.train_step(batches[i % len(batches)])
    t0 = time.perf_counter()
    for i in range(steps):
        model.train_step(batches[(warmup + i) % len(batches)])
    dt = time.perf_counter() - t0
    return a.batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(a):
    """The CPU arm.  One process on rank 0 using every host core, at this arm's workload (cfg2 shape, batch
    ``a.batch``) whatever N is: the reference has no multi-device mode (torchrec/task/Task.py:187-190) and one
    process already owns the whole box's cores, so its samples/s IS the box's CPU throughput at every N.  The line
    reports the steps / warm-up really executed (bounded to ~2 minutes of CPU time) and the configuration that ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from pytorchrec_b200.data import criteo_batch
    torch.set_num_threads(os.cpu_count() or 1)
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    model = build_cpu_model(a)
    batches = [criteo_batch(a.batch, CFG["n_sparse"], CFG["n_dense"], a.rows, seed=900 + i, dist=a.id_dist)
               for i in range(min(4, steps + warmup))]
    t0 = time.perf_counter()
    model.train_step(batches[0])                      # first (cold) step: also the probe that bounds the run
    probe = time.perf_counter() - t0
    budget = 120.0                                    # seconds of CPU stepping, warm-up included
    warm_run = max(0, min(warmup - 1, int(0.25 * budget / probe)))
    steps_run = max(1, min(steps, int(0.75 * budget / probe)))
    for i in range(warm_run):
        model.train_step(batches[(1 + i) % len(batches)])
    t0 = time.perf_counter()
    for i in range(steps_run):
        model.train_step(batches[(1 + warm_run + i) % len(batches)])
    dt = time.perf_counter() - t0
    sps, ms, cores = a.batch * steps_run / dt, dt / steps_run * 1e3, torch.get_num_threads()
    cfg = {"workload": "%s DeepFM Criteo-shaped synthetic: 26 tables x %d rows x dim %d (+26 first-order dim 1), "
                       "13 dense, DNN 400-400-400, batch %d per process, fp32, Adagrad" % (a.workload, a.rows, a.dim, a.batch),
           "table_update": "dense torch.optim.Adagrad (every row, as the reference's optimizers do)",
           "global_batch": a.batch, "id_dist": a.id_dist, "cuda_graph": False,
           "parallelism": f"one CPU process, {cores} threads (the reference is single-device; the same CPU run is "
                          f"the baseline at every N)",
           "gpu_arm_n_gpus": a.gpus}
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": sps, "unit": "samples/s",
            "n_gpus": a.gpus, "steps": steps_run, "warmup": warm_run + 1, "steps_requested": steps,
            "warmup_requested": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{steps_run} full-size train steps (batch {a.batch}) after {warm_run + 1} warm-up, "
                                       "oracle port of the reference idiom (nn.Embedding per column, dense autograd "
                                       "gradients, dense torch.optim.Adagrad, reference train_step) on host cores"},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        pk = json.load(open(path))
        return pk["hbm_gbs"], pk.get("bf16_tflops", 2250.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 2250.0, "fallback (B200_PROFILING.md / nominal)"


def measure_deepfm(a, dev, rank, world, lib, steps, warmup, with_e2e, sampler=None):
    """Build the DeepFM of workload ``a`` (unsharded at N = 1, row-wise sharded otherwise), warm it up and time
    ``steps`` train steps on HBM-resident batches (and, with ``with_e2e``, on pinned host batches with the loss read
    back every step).  Device-timed, barrier + synchronize on both sides, max over ranks."""
    import torch
    import torch.distributed as dist

    from pytorchrec_b200.data import criteo_batch, criteo_columns
    from pytorchrec_b200.metric import LogLoss
    from pytorchrec_b200.optim import SparseAdagrad

    sparse, dense, label = criteo_columns(CFG["n_sparse"], CFG["n_dense"], a.rows)
    if world == 1:
        from pytorchrec_b200.model import DeepFM
        model = DeepFM(sparse, dense, label, a.dim, CFG["layers"], random_seed=2020, table_device=dev)
    else:
        from pytorchrec_b200.distributed import ShardedDeepFM
        model = ShardedDeepFM(sparse, dense, label, a.dim, CFG["layers"], random_seed=2020, table_device=dev)
    a.exchange = getattr(getattr(model, "sharded", None), "exchange", "none")
    opt = SparseAdagrad(params=model.get_parameters(), lr=0.01)
    model.compile(opt, torch.nn.BCEWithLogitsLoss(), [LogLoss()], dev)
    if not a.no_graph:
        model.enable_cuda_graph(True)

    n_pool = 8
    host = [criteo_batch(a.batch, CFG["n_sparse"], CFG["n_dense"], a.rows, seed=1000 * (rank + 1) + i,
                         dist=a.id_dist, pin=True) for i in range(n_pool)]
    resident = [model.stage(b) for b in host]  # HBM-resident batches, each in one packed buffer
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    # end-to-end batches: assembled by the "loader" in ONE pinned buffer each (IModel.pack_host), so that a step's input
    # moves with a single H2D copy instead of 40 (one per key: ~0.25 ms of host time per step, more than the GPU's idle
    # margin); a plain dict of separately pinned tensors (--per-key-h2d) takes the 40-copy path
    if not a.per_key_h2d:
        host = [model.pack_host(b) for b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, n, read_loss, pipelined=False):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        if read_loss:  # end to end: the data-loader pattern of IModel.fit (prefetch batch k+1 while step k runs)
            model.prefetch(batches[0])
        if read_loss and pipelined:
            # every step's loss still reaches the host inside the timed region, but through a pinned buffer read AFTER
            # the next step has been enqueued, so the device does not idle while the host issues the next launch
            pin = torch.empty(2, dtype=torch.float32).pin_memory()
            evs = [None, None]
            for i in range(n):
                logs = model.train_step(batches[i % len(batches)])
                pin[i & 1:(i & 1) + 1].copy_(logs["loss"].detach().reshape(1), non_blocking=True)
                evs[i & 1] = torch.cuda.Event()
                evs[i & 1].record()
                if i + 1 < n:
                    model.prefetch(batches[(i + 1) % len(batches)])
                if i > 0:
                    evs[(i - 1) & 1].synchronize()
                    float(pin[(i - 1) & 1])
            evs[(n - 1) & 1].synchronize()
            float(pin[(n - 1) & 1])
            n = 0  # the loop below is skipped
        for i in range(n):
            logs = model.train_step(batches[i % len(batches)])
            if read_loss:
                if i + 1 < n:
                    model.prefetch(batches[(i + 1) % len(batches)])
                logs["loss"].item()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for i in range(warmup):
        model.train_step(resident[i % n_pool])
    if sampler is not None:
        sampler.start()

    def launch_total():
        g = getattr(model, "_graphed", None)
        return lib.ptrec_launch_count() + (g.replayed_launches if g is not None else 0)

    l0 = launch_total()
    ms = timed(resident, steps, read_loss=False)
    launches = launch_total() - l0
    ms_e2e = ms_e2e_sync = None
    if with_e2e:
        model.prefetch(host[0])
        for i in range(max(warmup, 3)):  # warm the end-to-end path too (pinned staging, copy stream, prefetch buffers)
            logs = model.train_step(host[i % n_pool])
            model.prefetch(host[(i + 1) % n_pool])
            logs["loss"].item()
        model.train_step(host[max(warmup, 3) % n_pool])
        # (a) every step: H2D of the batch (prefetched on the copy stream), train_step, D2H of the loss into a pinned
        #     word that the host reads once the NEXT step is enqueued (asynchronous logging: the device never waits for
        #     the host); (b) the same with `loss.item()` — a full host synchronisation — after every step
        ms_e2e = timed(host, steps, read_loss=True, pipelined=not a.sync_loss)
        model.prefetch(host[0])
        ms_e2e_sync = timed(host, steps, read_loss=True, pipelined=False)
    for m in model.modules():  # overflowed exchange lists / out-of-range ids invalidate the run: raise, don't report
        chk = getattr(m, "check_errors", None) or getattr(m, "check_index_errors", None)
        if chk is not None and m is not model:
            chk()
    return dict(model=model, resident=resident, host=host, ms=ms, ms_e2e=ms_e2e, ms_e2e_sync=ms_e2e_sync,
                launches=int(launches), h2d=h2d)


def _release(res, world):
    """Drop a measured model (graphs first: they pin NCCL kernels and every static buffer) and return its HBM."""
    import gc

    import torch
    res["model"].enable_cuda_graph(False)
    res.clear()
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def run_b200(a):
    import torch
    import torch.distributed as dist

    from pytorchrec_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    sampler = ClockSampler(local) if rank == 0 else None
    res = measure_deepfm(a, dev, rank, world, lib, a.steps, a.warmup, with_e2e=True, sampler=sampler)
    clocks = sampler.stop() if rank == 0 else None
    ms, ms_e2e = res["ms"], res["ms_e2e"]
    total = a.batch * world * a.steps
    line = {"metric": "train_samples_per_sec", "value": total / (ms / 1e3), "unit": "samples/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(a, world),
            "e2e": {"value": total / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": res["h2d"],
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / a.steps,
                    "host_batches": ("dict of separately pinned tensors: one H2D copy per key" if a.per_key_h2d else
                                     "one pinned packed buffer per batch (IModel.pack_host): one H2D copy per step, issued "
                                     "on the copy stream while the previous step runs"),
                    "loss_read": ("loss.item() after every step (host synchronisation per step)" if a.sync_loss else
                                  "every step's loss is copied to a pinned host word inside the step's stream and read by "
                                  "the host once the next step is enqueued (asynchronous logging; the last one "
                                  "synchronously, inside the timed region)"),
                    "sync_each_step": {"value": total / (res["ms_e2e_sync"] / 1e3), "ms_per_step": res["ms_e2e_sync"] / a.steps,
                                       "loss_read": "loss.item() after every step"}},
            "gpu_launches": res["launches"], "clocks": clocks}
    if rank == 0:
        roof, line["kernels"], roof2, line["roofline_hbm_all"] = kernel_roofline(a, res["model"], res["resident"], dev)
        line["roofline"] = roof
        if roof2 is not None:
            line["roofline_hbm"] = roof2  # the slower of the two HBM-bound kernels of the embedding path
    _release(res, world)

    # ---- cfg5: the north star's sharded config at THIS N (weak scaling: 26 x 5e7*N/8 rows x D64, B 65536 per GPU) ----
    if a.workload == "cfg2" and not a.no_cfg5:
        c5 = argparse.Namespace(**vars(a))
        c5.workload, c5.rows, c5.dim, c5.batch = "cfg5", int(5e7 * world / 8), 64, 65536
        try:
            r5 = measure_deepfm(c5, dev, rank, world, lib, a.sub_steps, max(3, min(a.warmup, 5)), with_e2e=False)
            sub = {"workload": "cfg5 DeepFM: 26 tables x %d rows x dim 64 (+26 first-order), B 65536 per GPU, fp32, sparse "
                               "Adagrad; weak-scaling series rows = 5e7*N/8" % c5.rows,
                   "n_gpus": world, "steps": a.sub_steps, "ms_per_step": r5["ms"] / a.sub_steps,
                   "value": c5.batch * world * a.sub_steps / (r5["ms"] / 1e3), "unit": "samples/s",
                   "global_batch": c5.batch * world, "gpu_launches": r5["launches"],
                   "table_bytes_per_gpu": 26 * ((c5.rows + world - 1) // world) * (64 + 1) * 4 * 2,
                   "parallelism": config_dict(c5, world)["parallelism"]}
            if rank == 0:
                _, k5, _, hbm5 = kernel_roofline(c5, r5["model"], r5["resident"], dev, tensor=False)
                sub["roofline_hbm_all"] = hbm5
                sub["kernels"] = k5
            _release(r5, world)
        except Exception as e:  # noqa  (reported, never hidden: the cfg2 line above is still valid)
            sub = {"error": f"{type(e).__name__}: {e}"[:400]}
        line["cfg5"] = sub

    if rank == 0:
        if world == 1 and not a.no_models:
            try:
                line["models"] = model_records(a, dev)
            except Exception as e:  # noqa
                line["models"] = {"error": f"{type(e).__name__}: {e}"[:400]}
        if world == 1 and not a.no_cpu_baseline:
            try:
                sps, cms, cores = time_cpu(a, a.cpu_steps, 1)
                line["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                                        "ms_per_step": cms,
                                        "sample": f"{a.cpu_steps} full-size train steps (batch {a.batch}) after 1 warm-up; "
                                                  "oracle port of the reference idiom (nn.Embedding, dense grads, dense Adagrad)"}
            except Exception as e:  # noqa
                line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {e}"}
        emit(line)
    if world > 1:
        # NCCL kernels captured in the step graph keep the communicator busy at teardown: synchronise and leave
        # without the (hanging) communicator destructor.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def _time_steps(model, batches, steps, warm):
    import torch
    for i in range(warm):
        model.train_step(batches[i % len(batches)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        model.train_step(batches[i % len(batches)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _time_kernel(fn, reps=20, warm=3):
    import torch
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def model_records(a, dev):
    """cfg3 (DCN-v2) and cfg4 (DIN) at their BASELINE.json sizes on one GPU: whole train step (HBM-resident batches,
    step graph) and the interaction kernels alone against their rooflines."""
    import gc

    import torch

    from pytorchrec_b200 import ops
    from pytorchrec_b200.data import amazon_batch, amazon_columns, criteo_batch, criteo_columns
    from pytorchrec_b200.metric import LogLoss
    from pytorchrec_b200.model import DCN, DIN
    from pytorchrec_b200.optim import SparseAdagrad
    hbm_peak, tc_peak, src = _peaks()
    out = {"peak_source": src}

    def compiled(m):
        m.compile(SparseAdagrad(m.get_parameters(), lr=0.01), torch.nn.BCEWithLogitsLoss(), [LogLoss()], dev)
        if not a.no_graph:
            m.enable_cuda_graph(True)
        return m

    # ---- cfg3
    B, rows, D = 32768, 1_000_000, 32
    sparse, dense, label = criteo_columns(26, 13, rows)
    m = compiled(DCN(sparse, dense, label, D, 3, [1024, 1024, 1024], random_seed=1, table_device=dev))
    batches = [m.stage(criteo_batch(B, 26, 13, rows, seed=i)) for i in range(4)]
    ms = _time_steps(m, batches, a.sub_steps, 6)
    d = 26 * D + 13
    dp = (d + 7) // 8 * 8
    xs = [torch.randn(B, dp, device=dev).bfloat16() * 0.5 for _ in range(4)]
    W = (torch.randn(dp, dp, device=dev) / dp ** 0.5).bfloat16()
    Wt = W.t().contiguous()
    bias = torch.randn(dp, device=dev) * 0.1
    fl = 2.0 * B * dp * dp
    t_f = _time_kernel(lambda i: ops.dcn_cross_fwd(xs[i % 4], xs[(i + 1) % 4], W, bias))
    t_d = _time_kernel(lambda i: ops.dcn_cross_dgrad(xs[i % 4], Wt, xs[(i + 1) % 4], xs[(i + 2) % 4]))
    t_w = _time_kernel(lambda i: ops.dcn_cross_wgrad(xs[i % 4], xs[(i + 1) % 4]))
    out["cfg3_dcn"] = {
        "workload": "cfg3 DCN-v2: 26 x 1e6 rows x dim 32 + 13 dense (d = 845 -> %d), 3 cross layers bf16 on tcgen05, "
                    "DNN 1024x3 fp32 (K6), batch 32768, sparse Adagrad" % dp,
        "ms_per_step": ms, "value": B / ms * 1e3, "unit": "samples/s",
        "roofline": {"bound": "tensor", "kernel": "dcn_cross_fwd (K5)", "achieved": fl / t_f / 1e12, "peak": tc_peak,
                     "unit": "TFLOP/s", "frac": fl / t_f / 1e12 / tc_peak, "seconds_per_launch": t_f,
                     "algorithmic_flops_per_launch": fl},
        "kernels": {"dcn_cross_fwd": {"seconds": t_f, "TFLOPs": fl / t_f / 1e12},
                    "dcn_cross_dgrad": {"seconds": t_d, "TFLOPs": fl / t_d / 1e12},
                    "dcn_cross_wgrad": {"seconds": t_w, "TFLOPs": fl / t_w / 1e12}}}
    m.enable_cuda_graph(False)
    del m, batches, xs, W, Wt
    gc.collect()
    torch.cuda.empty_cache()

    # ---- cfg4
    B, L, D = 8192, 100, 16
    cols = amazon_columns(L)
    m = compiled(DIN(*cols, emb_size=D, layers=[200, 80], random_seed=1, table_device=dev))
    batches = [m.stage(amazon_batch(B, L, seed=i)) for i in range(4)]
    ms = _time_steps(m, batches, a.sub_steps, 6)
    DQ, H1, H2 = 2 * D, 80, 40
    g = torch.Generator(device=dev).manual_seed(0)
    att = m.attention
    params = [p.detach() for p in (att.fc1.weight, att.fc1.bias, att.fc2.weight, att.fc2.bias, att.fc3.weight, att.fc3.bias)]
    seqs = [torch.randn(B, 1 + L, DQ, device=dev, generator=g) for _ in range(3)]
    lens = torch.randint(1, L + 1, (B,), device=dev, generator=g).int()
    go = torch.randn(B, DQ, device=dev, generator=g)
    n_pos = float(lens.sum().item())
    t_f = _time_kernel(lambda i: ops.din_attn_pool_fwd(seqs[i % 3][:, 0], seqs[i % 3][:, 1:], lens, params))
    t_b = _time_kernel(lambda i: ops.din_attn_pool_bwd(seqs[i % 3][:, 0], seqs[i % 3][:, 1:], lens, params, go))
    flops = n_pos * 2 * (H1 * 4 * DQ + H2 * H1 + H2)      # the activation unit as the paper states it
    byts = n_pos * DQ * 4
    out["cfg4_din"] = {
        "workload": "cfg4 DIN: Amazon-Books-shaped (603668 users, 367982 items, 1600 categories), histories of 100, "
                    "dim 16 per table (q / k = 32), unit 80-40, DNN 200-80, batch 8192, sparse Adagrad",
        "ms_per_step": ms, "value": B / ms * 1e3, "unit": "samples/s",
        "roofline": {"bound": "hbm", "kernel": "din_attn_pool_fwd (K4)", "achieved": byts / t_f / 1e9, "peak": hbm_peak,
                     "unit": "GB/s", "frac": byts / t_f / 1e9 / hbm_peak, "seconds_per_launch": t_f,
                     "algorithmic_bytes_per_launch": byts, "TFLOPs": flops / t_f / 1e12,
                     "note": "keys bytes B*L*DQ*4 over valid positions; FLOPs = positions * 2 * (4DQ*80 + 80*40 + 40)"},
        "kernels": {"din_attn_pool_fwd": {"seconds": t_f, "GBps": byts / t_f / 1e9, "TFLOPs": flops / t_f / 1e12},
                    "din_attn_pool_bwd": {"seconds": t_b, "GBps": 2 * byts / t_b / 1e9, "TFLOPs": 3 * flops / t_b / 1e12}}}
    m.enable_cuda_graph(False)
    del m, batches, seqs
    gc.collect()
    torch.cuda.empty_cache()
    return out


def kernel_roofline(a, model, resident, dev, tensor=True):
    """Time this library's hot kernels alone with CUDA events on the launching stream, each launch on a
    different id batch (tables >> L2), and convert with the algorithmic byte counts of DESIGN.md.
    Returns (roofline of the dominant kernel, per-kernel timings, roofline of the slower HBM kernel, rooflines of
    BOTH HBM kernels of the embedding path)."""
    import torch

    from pytorchrec_b200 import _lib, ops
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    emb = model.embeddings
    emb_tables = [t for t in emb.groups[0]] if hasattr(emb, "groups") else [t for t in emb]
    F, D, B = len(emb_tables), a.dim, a.batch
    R = min(t.weight.shape[0] for t in emb_tables)
    # private copies in the layout the optimizer uses (weight | Adagrad sum interleaved in one 2*D-float row) ...
    if R * F * D * 8 < 8e9:
        bufs = [torch.randn(t.weight.shape[0], 2 * D, device=dev) for t in emb_tables]
        for b in bufs:
            b[:, D:].abs_()
    else:  # ... except at cfg5 size (83 GB per GPU): the measured model's own interleaved buffers, in place — the
        # update below runs with lr = 0 (weights unchanged) and the model is discarded after this call
        opt = model.compiled_optimizers
        bufs = [opt._interleaved[id(t.weight)] for t in emb_tables]
    tables = [b[:, :D] for b in bufs]
    lay = ops.FeatureLayout([dict(table=f, bag_len=1) for f in range(F)], D, F)
    ts = ops.TableSet().refresh(tables)
    id_batches = [(torch.cat([b[c.feature_name].reshape(-1) for c in emb.columns]) % R).contiguous() for b in resident]
    state = [b[:, D:2 * D] for b in bufs]
    p1 = ops.make_ptr_array(state)
    go = torch.randn(B, F * D, device=dev)
    args = _lib.OptimArgs(kind=_lib.OPT_ADAGRAD, step=1, lr=0.0, eps=1e-10, beta1=0, beta2=0, weight_decay=0, lr_decay=0)
    srts = [ops.sort_dedup(ts, lay, ids, None, B) for ids in id_batches]
    n_seg = [int(s.n_seg.item()) for s in srts]
    lookups = F * B

    def time_it(fn, reps):
        """Average GPU time of one call: `reps` calls (rotating operands) captured in ONE CUDA graph and replayed, CUDA
        events around the replays on the launching stream — no host enqueue time between the launches, which for these
        20 us kernels is as long as the kernels themselves.  Falls back to an eager launch loop if capture fails."""
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        graph = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(2):
                    fn(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for i in range(reps):
                    fn(i)
            graph.replay()
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001 — a kernel wrapper that cannot be captured is timed eagerly
            graph = None
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if graph is not None:
            e0.record()
            for _ in range(3):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / (3 * reps) * 1e-3
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    out = torch.empty(B, F * D, device=dev)
    nb = len(id_batches)
    t_gather = time_it(lambda i: ops.gather_pool_fwd(ts, lay, id_batches[i % nb], None, B, out=out), 40)
    t_sort = time_it(lambda i: ops.sort_dedup(ts, lay, id_batches[i % nb], None, B), 40)
    t_upd = time_it(lambda i: ops.bwd_fused(ts, p1, None, lay, B, srts[i % nb], go, None, args), 40)
    v3 = out.view(B, F, D)
    t_fm = time_it(lambda i: ops.fm2_fwd(v3), 40)
    gy = torch.randn(B, device=dev)
    t_fmb = time_it(lambda i: ops.fm2_bwd(v3, gy), 40)
    U = sum(n_seg) / len(n_seg)
    bytes_gather = lookups * 8 + lookups * D * 4 + lookups * D * 4
    bytes_upd = lookups * (4 + D * 4) + U * (8 + 2 * D * 4 + 2 * D * 4)
    bytes_sort = lookups * 8 + 3 * lookups * 16 + lookups * 8
    kernels = {
        "gather_pool_fwd": {"seconds": t_gather, "bytes": bytes_gather, "GBps": bytes_gather / t_gather / 1e9},
        "fused_update_adagrad": {"seconds": t_upd, "bytes": bytes_upd, "GBps": bytes_upd / t_upd / 1e9, "unique_rows": U},
        "sort_dedup(9 launches)": {"seconds": t_sort, "bytes": bytes_sort, "GBps": bytes_sort / t_sort / 1e9},
        "fm2_fwd": {"seconds": t_fm, "bytes": B * F * D * 4 + B * 4, "GBps": (B * F * D * 4 + B * 4) / t_fm / 1e9},
        "fm2_bwd": {"seconds": t_fmb, "bytes": 2 * B * F * D * 4, "GBps": 2 * B * F * D * 4 / t_fmb / 1e9},
    }
    hbm = {k: v for k, v in kernels.items() if k in ("gather_pool_fwd", "fused_update_adagrad")}
    dom = max(hbm, key=lambda k: hbm[k]["seconds"])
    traffic_all = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and a.workload == "cfg2":
        traffic_all = json.load(open(tpath))

    def hbm_roof(k):
        return {"kernel": k, "bound": "hbm", "achieved": kernels[k]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kernels[k]["GBps"] / peak, "traffic": traffic_all.get(k), "peak_source": peak_src,
                "frac_of_spec_8000": kernels[k]["GBps"] / 8000.0,
                "algorithmic_bytes_per_launch": kernels[k]["bytes"], "seconds_per_launch": kernels[k]["seconds"]}
    roof_hbm = hbm_roof(dom)
    hbm_all = {k: hbm_roof(k) for k in hbm}
    if not tensor:
        return roof_hbm, kernels, roof_hbm, hbm_all

    # K6: the DNN-tower GEMMs (9 launches per step: the largest share of the step).  Tensor-bound: each fp32 product
    # is 6 bf16 plane-pair MMAs, so a launch issues 12*M*N*K bf16 FLOPs.  Operands rotate over 4
    # buffer sets (260 MB > L2).
    from pytorchrec_b200.model.layer.dense import tc_linear_enabled
    mlp = getattr(model, "mlp", None)
    if not (tc_linear_enabled() and mlp is not None):
        return roof_hbm, kernels, None, hbm_all
    lin = mlp.mlp[0].linear
    N, K = lin.weight.shape
    xs = [torch.randn(B, K, device=dev) for _ in range(4)]
    gs = [torch.randn(B, N, device=dev) for _ in range(4)]  # the gradient entering the layer: dW = g^T x is [N, K]
    bias = lin.bias.detach()
    h2 = ops.tc_mode() == "fp16x2"
    if h2:   # fp16 x 2 operands: 3 MMAs per product, two planes, an |x|-maximum pass inside the split call
        hxs = [ops.tc_split2h(x) for x in xs]
        hgs = [ops.tc_split2h(g) for g in gs]
        hw = ops.tc_split2h(lin.weight.detach())
        t_plain = time_it(lambda i: ops.tc_gemm_split2h(hxs[i % 4][0], hxs[i % 4][3], hw[0], hw[3], K, bias=bias, relu=True), 40)
        # the form the fused tower launches for a hidden layer: bias + ReLU, the result written as the next layer's
        # fp16 planes + the ReLU bit mask + its maximum (no fp32 output, no split pass afterwards)
        oscale = torch.full((1,), 2.0 ** 4, device=dev)
        omax = torch.zeros(1, device=dev)
        t_gemm = time_it(lambda i: ops.tc_gemm_split2h_fused(hxs[i % 4][0], hxs[i % 4][3], hw[0], hw[3], K, bias=bias, relu=True,
                                                             want_out=False, out_scale=oscale, want_mask=True, max_out=omax), 40)
        t_wgrad = time_it(lambda i: ops.tc_gemm_split2h_tn(hgs[i % 4][0], hgs[i % 4][3], N, hxs[i % 4][0],
                                                           hxs[i % 4][3], K), 20)
        t_split = time_it(lambda i: ops.tc_split2h(xs[i % 4]), 40)
        pairs, planes, fmt, kname = 3, 2, "fp16", "gemm_split2h_fused"
        kernels["tc_linear_fwd(gemm_split2h, fp32 out)"] = {"seconds": t_plain, "M": B, "N": N, "K": K,
                                                             "TFLOPs_fp16": 2.0 * 3 * B * N * K / t_plain / 1e12}
    else:
        pxs = [ops.tc_split3(x)[0] for x in xs]
        pgs = [ops.tc_split3(g)[0] for g in gs]
        pw = ops.tc_split3(lin.weight.detach())[0]
        t_gemm = time_it(lambda i: ops.tc_gemm_split3(pxs[i % 4], pw, K, bias=bias, relu=True), 40)
        t_wgrad = time_it(lambda i: ops.tc_gemm_split3_tn(pgs[i % 4], N, pxs[i % 4], K), 20)
        t_split = time_it(lambda i: ops.tc_split3(xs[i % 4]), 40)
        pairs, planes, fmt, kname = 6, 3, "bf16", "gemm_split3"
    flops = 2.0 * pairs * B * N * K
    kernels[f"tc_linear_fwd({kname})"] = {"seconds": t_gemm, f"flops_{fmt}_issued": flops, f"TFLOPs_{fmt}": flops / t_gemm / 1e12,
                                          "TFLOPs_fp32_equivalent": flops / pairs / t_gemm / 1e12, "M": B, "N": N, "K": K}
    kernels[f"tc_linear_wgrad({kname}_tn)"] = {"seconds": t_wgrad, "M": N, "N": K, "K": B,
                                               f"TFLOPs_{fmt}": 2.0 * pairs * B * N * K / t_wgrad / 1e12,
                                               "TFLOPs_fp32_equivalent": 2.0 * B * N * K / t_wgrad / 1e12}
    sp_bytes = B * K * 4 * (2 if h2 else 1) + planes * B * ((K + 7) // 8 * 8) * 2
    kernels["tc_split2h" if h2 else "tc_split3"] = {"seconds": t_split, "bytes": sp_bytes, "GBps": sp_bytes / t_split / 1e9}
    pk = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    tpeak = pk.get("bf16_tflops", 2250.0)
    roof = {"kernel": f"{kname} (K6, DNN tower Linear fwd)", "bound": "tensor", "achieved": flops / t_gemm / 1e12,
            "peak": tpeak, "unit": "TFLOP/s", "frac": flops / t_gemm / 1e12 / tpeak,
            # the same launch counted in the FLOPs of the fp32 product it computes (2*M*N*K), not in issued MMAs
            "achieved_algorithmic": flops / pairs / t_gemm / 1e12, "frac_algorithmic": flops / pairs / t_gemm / 1e12 / tpeak,
            "traffic": traffic_all.get(kname),
            "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone)" if pk else "nominal",
            "algorithmic_flops_per_launch": flops, "seconds_per_launch": t_gemm,
            "note": f"{fmt} FLOPs issued = {pairs} plane pairs x 2*M*N*K; fp32-equivalent rate = achieved / {pairs}"}
    return roof, kernels, roof_hbm, hbm_all


if __name__ == "__main__":
    args = parse()
    if args.workload == "cfg5":
        args.rows = int(5e7 * max(args.gpus, 1) / 8)
        args.dim, args.batch = 64, 65536
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
