"""GPU timing of individual kernels against their rooflines (CUDA events, different inputs per launch).
   python tools/bench_kernels.py dcn|din|emb"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchrec_b200 import _lib, ops

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timeit(fn, reps=20, warm=3):
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


def dcn():
    B, d = 32768, 848
    xs = [torch.randn(B, d, device=dev).bfloat16() * 0.5 for _ in range(4)]
    W = (torch.randn(d, d, device=dev) / d ** 0.5).bfloat16()
    b = torch.randn(d, device=dev) * 0.1
    res = {}
    t = timeit(lambda i: ops.dcn_cross_fwd(xs[i % 4], xs[(i + 1) % 4], W, b))
    res["cross_fwd"] = dict(seconds=t, tflops=2 * B * d * d / t / 1e12)
    Wt = W.t().contiguous()
    t = timeit(lambda i: ops.dcn_cross_dgrad(xs[i % 4], Wt, xs[(i + 1) % 4], xs[(i + 2) % 4]))
    res["cross_dgrad"] = dict(seconds=t, tflops=2 * B * d * d / t / 1e12)
    t = timeit(lambda i: ops.dcn_cross_wgrad(xs[i % 4], xs[(i + 1) % 4]))
    res["cross_wgrad(+2 transposes)"] = dict(seconds=t, tflops=2 * B * d * d / t / 1e12)
    xf = [x.float() for x in xs[:2]]
    Wf = W.float()
    t = timeit(lambda i: torch.addcmul(xf[0], xf[1], torch.nn.functional.linear(xf[i % 2], Wf, b)))
    res["torch_fp32_linear+addcmul"] = dict(seconds=t, tflops=2 * B * d * d / t / 1e12)
    t = timeit(lambda i: torch.addcmul(xs[0], xs[1], torch.nn.functional.linear(xs[i % 4], W, b.bfloat16())))
    res["torch_bf16_cublas_linear+addcmul"] = dict(seconds=t, tflops=2 * B * d * d / t / 1e12)
    for k, v in res.items():
        v["frac_of_bf16_peak"] = v["tflops"] / peaks["bf16_tflops"]
    return res


def din():
    B, L, DQ, H1, H2 = 8192, 100, 32, 80, 40
    g = torch.Generator(device=dev).manual_seed(0)
    params = [torch.randn(H1, 4 * DQ, device=dev, generator=g) * 0.1, torch.randn(H1, device=dev, generator=g) * 0.1,
              torch.randn(H2, H1, device=dev, generator=g) * 0.1, torch.randn(H2, device=dev, generator=g) * 0.1,
              torch.randn(1, H2, device=dev, generator=g) * 0.1, torch.randn(1, device=dev, generator=g) * 0.1]
    seqs = [torch.randn(B, 1 + L, DQ, device=dev, generator=g) for _ in range(3)]
    lens = torch.randint(1, L + 1, (B,), device=dev, generator=g).int()
    go = torch.randn(B, DQ, device=dev, generator=g)
    n_pos = float(lens.sum().item())
    res = {}
    t = timeit(lambda i: ops.din_attn_pool_fwd(seqs[i % 3][:, 0], seqs[i % 3][:, 1:], lens, params))
    flops = n_pos * 2 * (H1 * DQ + H2 * H1 + H2)
    bytes_ = n_pos * DQ * 4
    res["din_fwd"] = dict(seconds=t, gflops_fused=flops / 1e9, tflops=flops / t / 1e12, GBps=bytes_ / t / 1e9,
                          frac_hbm=bytes_ / t / 1e9 / peaks["hbm_gbs"], positions=n_pos)
    t = timeit(lambda i: ops.din_attn_pool_bwd(seqs[i % 3][:, 0], seqs[i % 3][:, 1:], lens, params, go))
    res["din_bwd"] = dict(seconds=t, tflops=3 * flops / t / 1e12, GBps=2 * bytes_ / t / 1e9,
                          frac_hbm=2 * bytes_ / t / 1e9 / peaks["hbm_gbs"])
    # the unfused torch path the kernel replaces (materialises [B, L, 4*DQ] and the hidden layers)
    W1, b1, W2, b2, W3, b3 = params

    def torch_fwd(i):
        q, k = seqs[i % 3][:, 0], seqs[i % 3][:, 1:]
        qe = q.unsqueeze(1).expand(B, L, DQ)
        z = torch.cat([qe, k, qe - k, qe * k], -1)
        a = torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(z, W1, b1)), W2, b2)), W3, b3).squeeze(-1)
        a = a * (torch.arange(L, device=dev).unsqueeze(0) < lens.unsqueeze(1))
        return (a.unsqueeze(-1) * k).sum(1)
    t = timeit(torch_fwd)
    res["torch_unfused_fwd"] = dict(seconds=t)
    return res


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "dcn"
    out = {"dcn": dcn, "din": din}[which]()
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/kernels_{which}.json", "w"), indent=1)
