#!/usr/bin/env python
"""bench.py — train samples/sec of the DeepFM hot path on B200 (BASELINE.json metric).

Own arm (default): cfg2 = DeepFM, Criteo-shaped synthetic, 26 tables x 1e6 rows x dim 16 (+26 first-order
tables dim 1), 13 dense, DNN 400-400-400, batch 16384 per GPU, fp32, fused sparse Adagrad.  A "step" is
one `IModel.train_step` (forward, BCE loss, backward with the fused sort/dedup/scatter/update, dense
optimizer step).  `value` = samples/s with batches resident in HBM; `e2e` = the same step fed from pinned
HOST batches (H2D inside the timed region) with the loss read back every step.

`--impl reference`: the reference's CPU path (oracle port of the reference idiom: nn.Embedding per column,
dense autograd gradients, dense torch.optim.Adagrad, driven by the reference's five-line train_step) timed
on this box's host cores, same config / metric / unit.
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
    ap.add_argument("--pipelined-loss", action="store_true",
                    help="e2e: read step k's loss from a pinned buffer after step k+1 is enqueued (experimental)")
    return ap.parse_args()


def config_dict(a, world):
    from pytorchrec_b200 import ops
    return {"dnn_gemm_operands": ops.tc_mode(),  # K6: fp32 Linear operands as bf16x3 or fp16x2 planes
            "workload": "%s DeepFM Criteo-shaped synthetic: 26 tables x %d rows x dim %d (+26 first-order dim 1), "
                        "13 dense, DNN 400-400-400, batch %d per GPU, fp32, sparse Adagrad" % (a.workload, a.rows, a.dim, a.batch),
            "global_batch": a.batch * world, "id_dist": a.id_dist, "cuda_graph": not a.no_graph,
            "parallelism": "single GPU" if world == 1 else (
                f"row-wise sharded tables x{world} ("
                + ("NVLink peer-memory gather / push inside the kernels" if getattr(a, "peer_path", True)
                   else "NCCL all-to-all") + ") + dense allreduce"),
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    # bounded: a CPU step at cfg2 moves ~10 GB (dense grads + dense Adagrad over 1.8 GB of tables)
    steps_run, warm_run = min(steps, 5), min(warmup, 1)
    sps, ms, cores = time_cpu(a, steps_run, warm_run)
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": sps, "unit": "samples/s",
            "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(a, max(a.gpus, 1)),
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{steps_run} full-size train steps (batch {a.batch}) after {warm_run} warm-up, "
                                       "oracle port of the reference idiom on host cores"},
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
def run_b200(a):
    import torch
    import torch.distributed as dist

    from pytorchrec_b200 import _lib, ops
    from pytorchrec_b200.data import criteo_batch, criteo_columns
    from pytorchrec_b200.metric import LogLoss
    from pytorchrec_b200.optim import SparseAdagrad

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

    sparse, dense, label = criteo_columns(CFG["n_sparse"], CFG["n_dense"], a.rows)
    if world == 1:
        from pytorchrec_b200.model import DeepFM
        model = DeepFM(sparse, dense, label, a.dim, CFG["layers"], random_seed=2020, table_device=dev)
    else:
        from pytorchrec_b200.distributed import ShardedDeepFM
        model = ShardedDeepFM(sparse, dense, label, a.dim, CFG["layers"], random_seed=2020, table_device=dev)
    a.peer_path = bool(getattr(getattr(model, "sharded", None), "peer", False))
    opt = SparseAdagrad(params=model.get_parameters(), lr=0.01)
    model.compile(opt, torch.nn.BCEWithLogitsLoss(), [LogLoss()], dev)
    if not a.no_graph:
        model.enable_cuda_graph(True)

    n_pool = 8
    host = [criteo_batch(a.batch, CFG["n_sparse"], CFG["n_dense"], a.rows, seed=1000 * (rank + 1) + i,
                         dist=a.id_dist, pin=True) for i in range(n_pool)]
    resident = [model.stage(b) for b in host]  # HBM-resident batches, each in one packed buffer
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, steps, read_loss):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        if read_loss:  # end to end: the data-loader pattern of IModel.fit (prefetch batch k+1 while step k runs)
            model.prefetch(batches[0])
        if read_loss and a.pipelined_loss:
            # experimental (off by default, not yet measured): every step's loss still reaches the host inside the
            # timed region, but through a pinned buffer read AFTER the next step has been enqueued, so the device
            # does not idle while the host issues the next prefetch + graph launch
            pin = torch.empty(2, dtype=torch.float32).pin_memory()
            evs = [None, None]
            for i in range(steps):
                logs = model.train_step(batches[i % len(batches)])
                pin[i & 1:(i & 1) + 1].copy_(logs["loss"].detach().reshape(1), non_blocking=True)
                evs[i & 1] = torch.cuda.Event()
                evs[i & 1].record()
                if i + 1 < steps:
                    model.prefetch(batches[(i + 1) % len(batches)])
                if i > 0:
                    evs[(i - 1) & 1].synchronize()
                    float(pin[(i - 1) & 1])
            evs[(steps - 1) & 1].synchronize()
            float(pin[(steps - 1) & 1])
            steps = 0  # the loop below is skipped
        for i in range(steps):
            logs = model.train_step(batches[i % len(batches)])
            if read_loss:
                if i + 1 < steps:
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

    for i in range(a.warmup):
        model.train_step(resident[i % n_pool])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    def launch_total():
        g = getattr(model, "_graphed", None)
        return lib.ptrec_launch_count() + (g.replayed_launches if g is not None else 0)

    l0 = launch_total()
    ms = timed(resident, a.steps, read_loss=False)
    launches = launch_total() - l0
    model.prefetch(host[0])
    for i in range(max(a.warmup, 3)):  # warm the end-to-end path too (pinned staging, copy stream, prefetch buffers)
        logs = model.train_step(host[i % n_pool])
        model.prefetch(host[(i + 1) % n_pool])
        logs["loss"].item()
    model.train_step(host[max(a.warmup, 3) % n_pool])
    ms_e2e = timed(host, a.steps, read_loss=True)
    clocks = sampler.stop() if rank == 0 else None
    if hasattr(model, "embeddings") and hasattr(model.embeddings, "check_index_errors"):
        model.embeddings.check_index_errors()

    total = a.batch * world * a.steps
    line = {"metric": "train_samples_per_sec", "value": total / (ms / 1e3), "unit": "samples/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(a, world),
            "e2e": {"value": total / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": int(launches), "clocks": clocks}

    if rank == 0:
        roof, line["kernels"], roof2 = kernel_roofline(a, model, resident, dev)
        line["roofline"] = roof
        if roof2 is not None:
            line["roofline_hbm"] = roof2  # the dominant HBM-bound kernel of the embedding path
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
        # NCCL kernels captured in the step graph keep the communicator busy at teardown: drop the graphs,
        # synchronise, and leave without the (hanging) communicator destructor.
        dist.barrier()
        model.enable_cuda_graph(False)
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def kernel_roofline(a, model, resident, dev):
    """Time this library's hot kernels alone with CUDA events on the launching stream, each launch on a
    different id batch (tables >> L2), and convert with the algorithmic byte counts of DESIGN.md."""
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
    # private copies in the layout the optimizer uses (weight | Adagrad sum interleaved in one 2*D-float row)
    bufs = [torch.randn(t.weight.shape[0], 2 * D, device=dev) for t in emb_tables]
    tables = [b[:, :D] for b in bufs]
    lay = ops.FeatureLayout([dict(table=f, bag_len=1) for f in range(F)], D, F)
    ts = ops.TableSet().refresh(tables)
    id_batches = [(torch.cat([b[c.feature_name].reshape(-1) for c in emb.columns]) % R).contiguous() for b in resident]
    state = [b[:, D:].abs_() for b in bufs]
    p1 = ops.make_ptr_array(state)
    go = torch.randn(B, F * D, device=dev)
    args = _lib.OptimArgs(kind=_lib.OPT_ADAGRAD, step=1, lr=0.0, eps=1e-10, beta1=0, beta2=0, weight_decay=0, lr_decay=0)
    srts = [ops.sort_dedup(ts, lay, ids, None, B) for ids in id_batches]
    n_seg = [int(s.n_seg.item()) for s in srts]
    lookups = F * B

    def time_it(fn, reps):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
    if os.path.exists(tpath):
        traffic_all = json.load(open(tpath))
    roof_hbm = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["GBps"] / peak, "traffic": traffic_all.get(dom), "peak_source": peak_src,
                "frac_of_spec_8000": kernels[dom]["GBps"] / 8000.0,
                "algorithmic_bytes_per_launch": kernels[dom]["bytes"], "seconds_per_launch": kernels[dom]["seconds"]}

    # K6: the DNN-tower GEMMs (9 launches per step: the largest share of the step).  Tensor-bound: each fp32 product
    # is 6 bf16 plane-pair MMAs, so a launch issues 12*M*N*K bf16 FLOPs.  Operands rotate over 4
    # buffer sets (260 MB > L2).
    from pytorchrec_b200.model.layer.dense import tc_linear_enabled
    mlp = getattr(model, "mlp", None)
    if not (tc_linear_enabled() and mlp is not None):
        return roof_hbm, kernels, None
    lin = mlp.mlp[0].linear
    N, K = lin.weight.shape
    xs = [torch.randn(B, K, device=dev) for _ in range(4)]
    bias = lin.bias.detach()
    h2 = ops.tc_mode() == "fp16x2"
    if h2:   # fp16 x 2 operands: 3 MMAs per product, two planes, an |x|-maximum pass inside the split call
        hxs = [ops.tc_split2h(x) for x in xs]
        hw = ops.tc_split2h(lin.weight.detach())
        t_gemm = time_it(lambda i: ops.tc_gemm_split2h(hxs[i % 4][0], hxs[i % 4][3], hw[0], hw[3], K, bias=bias, relu=True), 40)
        t_wgrad = time_it(lambda i: ops.tc_gemm_split2h_tn(hxs[i % 4][0], hxs[i % 4][3], K, hxs[(i + 1) % 4][0],
                                                           hxs[(i + 1) % 4][3], K), 20)
        t_split = time_it(lambda i: ops.tc_split2h(xs[i % 4]), 40)
        pairs, planes, fmt, kname = 3, 2, "fp16", "gemm_split2h"
    else:
        pxs = [ops.tc_split3(x)[0] for x in xs]
        pw = ops.tc_split3(lin.weight.detach())[0]
        t_gemm = time_it(lambda i: ops.tc_gemm_split3(pxs[i % 4], pw, K, bias=bias, relu=True), 40)
        t_wgrad = time_it(lambda i: ops.tc_gemm_split3_tn(pxs[i % 4], K, pxs[(i + 1) % 4], K), 20)
        t_split = time_it(lambda i: ops.tc_split3(xs[i % 4]), 40)
        pairs, planes, fmt, kname = 6, 3, "bf16", "gemm_split3"
    flops = 2.0 * pairs * B * N * K
    kernels[f"tc_linear_fwd({kname})"] = {"seconds": t_gemm, f"flops_{fmt}_issued": flops, f"TFLOPs_{fmt}": flops / t_gemm / 1e12,
                                          "TFLOPs_fp32_equivalent": flops / pairs / t_gemm / 1e12, "M": B, "N": N, "K": K}
    kernels[f"tc_linear_wgrad({kname}_tn)"] = {"seconds": t_wgrad, "M": K, "N": K, "K": B,
                                               f"TFLOPs_{fmt}": 2.0 * pairs * B * K * K / t_wgrad / 1e12}
    sp_bytes = B * K * 4 * (2 if h2 else 1) + planes * B * ((K + 7) // 8 * 8) * 2
    kernels["tc_split2h" if h2 else "tc_split3"] = {"seconds": t_split, "bytes": sp_bytes, "GBps": sp_bytes / t_split / 1e9}
    pk = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    tpeak = pk.get("bf16_tflops", 2250.0)
    roof = {"kernel": f"{kname} (K6, DNN tower Linear fwd)", "bound": "tensor", "achieved": flops / t_gemm / 1e12,
            "peak": tpeak, "unit": "TFLOP/s", "frac": flops / t_gemm / 1e12 / tpeak,
            "traffic": traffic_all.get(kname),
            "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone)" if pk else "nominal",
            "algorithmic_flops_per_launch": flops, "seconds_per_launch": t_gemm,
            "note": f"{fmt} FLOPs issued = {pairs} plane pairs x 2*M*N*K; fp32-equivalent rate = achieved / {pairs}"}
    return roof, kernels, roof_hbm


if __name__ == "__main__":
    args = parse()
    if args.workload == "cfg5":
        args.rows = int(5e7 * max(args.gpus, 1) / 8)
        args.dim, args.batch = 64, 65536
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
