"""K6 fused tower (include/ptrec_b200.h, "the fused tower"): carried scales, the prescaled split, and the GEMM epilogue
that writes its consumer's fp16 planes, the ReLU bit mask and the bias gradient.  The fused GEMM runs the same MMAs as
``ptrec_tc_gemm_split2h``, so its fp32 result must equal that kernel's bit for bit; everything the epilogue derives from
the result (planes, mask bits, maximum) is checked bit for bit against the restatement in oracle/ref_ops.py, the column
sums to summation-order tolerance, and the whole MLP against the per-layer path and an fp64 evaluation."""
import pytest
import torch

from oracle import ref_ops
from pytorchrec_b200 import _lib, ops

DEV = torch.device("cuda:0")
pytestmark = [pytest.mark.gpu]


@pytest.fixture(autouse=True)
def _fp16x2_operands():
    old = ops.tc_mode()
    ops.set_tc_mode("fp16x2")
    yield
    ops.set_tc_mode(old)


def _word(v: float) -> torch.Tensor:
    return torch.tensor([v], dtype=torch.float32, device=DEV)


def test_scale_roll_turns_maxima_into_scales_and_flags_overflow():
    maxima = [0.0, 1.0, 3.7e-9, 5.0e4, 47.9, 0.0]
    slots = torch.zeros(len(maxima), 2, device=DEV)
    slots[:, 1] = torch.tensor(maxima)
    slots[5, 0] = 0.25                      # a slot nobody wrote since the last roll keeps its scale
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    cs = ops.tc_scale_roll(slots, err)
    want = [1.0] + [ref_ops.h2_carried_scale_ref(m) for m in maxima[1:5]] + [0.25]
    assert cs.tolist() == want and slots[:, 0].tolist() == want and not slots[:, 1].any()
    for m, s in zip(maxima[1:5], want[1:5]):
        assert 32.0 <= m * s < 64.0
    assert err.item() == 0
    # growth within the headroom (x 255) passes, beyond the fp16 range (x 2100 -> 2^5 * 2100 > 65504) is flagged
    slots[:, 1] = torch.tensor([0.0, 255.0, 0, 0, 0, 0], device=DEV)
    ops.tc_scale_roll(slots, err)
    assert err.item() == 0
    slots[1, 0] = 32.0
    slots[1, 1] = 2100.0
    ops.tc_scale_roll(slots, err)
    assert err.item() == 1


@pytest.mark.parametrize("R,C", [(64, 64), (200, 429), (1031, 13), (4096, 400)])
@pytest.mark.parametrize("mag", [1.0, 1e-8, 3e5])
def test_prescaled_split_equals_the_restatement(R, C, mag):
    gen = torch.Generator().manual_seed(R + C)
    x = mag * torch.randn(R, C, generator=gen) * torch.exp(2 * torch.randn(R, C, generator=gen))
    y = torch.randn(R, C, generator=gen)
    s = ref_ops.h2_carried_scale_ref(float(x.abs().max()) * 3.0)   # a stale scale: the tensor shrank 3-fold
    mx = _word(0.0)
    pl, plt, cs = ops.tc_split2h_prescaled(x.to(DEV), _word(s), mx, want_t=True, want_colsum=True)
    h0, h1 = ref_ops.split2h_prescaled_ref(x, s)
    assert torch.equal(pl[0, :, :C].cpu(), h0) and torch.equal(pl[1, :, :C].cpu(), h1)
    assert torch.equal(plt[0, :, :R].cpu(), h0.t()) and torch.equal(plt[1, :, :R].cpu(), h1.t())
    assert mx.item() == x.abs().max().item()
    assert torch.allclose(cs.double().cpu(), x.double().sum(0), rtol=1e-5, atol=1e-5 * x.abs().sum(0).max().item())
    # the two planes reproduce x to 2^-22 relative (elements far below the maximum: to 2^-39 of the maximum with a stale scale)
    back = (h0.double() + h1.double() / 2048.0) / s
    assert ((back - x.double()).abs() <= 2.0 ** -22 * x.double().abs() + 2.0 ** -38 * float(x.abs().max())).all()
    # ReLU backward: the maximum is that of the masked tensor
    mx2 = _word(0.0)
    pl2, _, cs2 = ops.tc_split2h_prescaled(x.to(DEV), _word(s), mx2, relu_ref=y.to(DEV), want_colsum=True)
    m0, m1 = ref_ops.split2h_prescaled_ref(x, s, y > 0)
    assert torch.equal(pl2[0, :, :C].cpu(), m0) and torch.equal(pl2[1, :, :C].cpu(), m1)
    assert mx2.item() == (x * (y > 0)).abs().max().item()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 400, 429), (1000, 16, 40), (2048, 429, 400), (130, 1, 700),
                                   (4096, 400, 1677), (257, 290, 33)])
def test_fused_gemm_epilogue_outputs(M, N, K):
    gen = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=gen).abs().to(DEV)
    b = (0.05 * torch.randn(N, K, generator=gen)).to(DEV)
    bias = (0.3 * torch.randn(N, generator=gen)).to(DEV)
    pa, _, _, sa = ops.tc_split2h(a)
    pb, _, _, sb = ops.tc_split2h(b)
    base = ops.tc_gemm_split2h(pa, sa, pb, sb, K, bias=bias, relu=True)
    # forward form: fp32 + planes + mask of the positive entries + maximum
    s_out = ref_ops.h2_carried_scale_ref(float(base.abs().max()) * 0.5)
    mx = _word(0.0)
    y, py, mask, _ = ops.tc_gemm_split2h_fused(pa, sa, pb, sb, K, bias=bias, relu=True, want_out=True,
                                               out_scale=_word(s_out), want_mask=True, max_out=mx)
    assert torch.equal(y, base)
    h0, h1 = ref_ops.split2h_prescaled_ref(base.cpu(), s_out)
    assert torch.equal(py[0, :, :N].cpu(), h0) and torch.equal(py[1, :, :N].cpu(), h1)
    assert torch.equal(ref_ops.unpack_mask_ref(mask.cpu(), N), base.cpu() > 0)
    assert mx.item() == base.abs().max().item()
    # planes only (what a hidden layer writes)
    y2, py2, mask2, _ = ops.tc_gemm_split2h_fused(pa, sa, pb, sb, K, bias=bias, relu=True, want_out=False,
                                                  out_scale=_word(s_out), want_mask=True)
    assert y2 is None and torch.equal(py2[:, :, :N], py[:, :, :N]) and torch.equal(
        ref_ops.unpack_mask_ref(mask2.cpu(), N), base.cpu() > 0)
    # backward form: no bias / ReLU, the result multiplied by a bit mask, column sums, planes, maximum of the masked
    plain = ops.tc_gemm_split2h(pa, sa, pb, sb, K)
    keep = ref_ops.unpack_mask_ref(mask.cpu(), N)
    masked = plain.cpu() * keep
    s_g = ref_ops.h2_carried_scale_ref(float(masked.abs().max()))
    mx3 = _word(0.0)
    y3, py3, _, cs = ops.tc_gemm_split2h_fused(pa, sa, pb, sb, K, want_out=True, out_scale=_word(s_g), mask_in=mask,
                                               want_colsum=True, max_out=mx3)
    assert torch.equal(y3.cpu(), masked)
    g0, g1 = ref_ops.split2h_prescaled_ref(masked, s_g)
    assert torch.equal(py3[0, :, :N].cpu(), g0) and torch.equal(py3[1, :, :N].cpu(), g1)
    assert mx3.item() == masked.abs().max().item()
    ref_cs = masked.double().sum(0)
    assert (cs.double().cpu() - ref_cs).abs().max().item() <= 1e-6 * masked.double().abs().sum(0).max().item() + 1e-30
    # plain fp32 result through the staged epilogue (what the first layer's input gradient uses)
    y4, _, _, _ = ops.tc_gemm_split2h_fused(pa, sa, pb, sb, K)
    assert torch.equal(y4, plain)


def _fp64_tower(mlp, x, gy, masks):
    """fp64 evaluation of the tower that takes the ReLU decisions ``masks`` (bool [B, N_l] per layer) the device took:
    a pre-activation within rounding of zero is on in one arithmetic and off in the other, and the gradients of the two
    choices differ by far more than any tolerance.  Asserts that the decisions are the fp64 ones wherever the
    pre-activation is not ambiguous."""
    xd = x.detach().double().requires_grad_(True)
    ws = [l.linear.weight.detach().double().requires_grad_(True) for l in mlp.mlp]
    bs = [l.linear.bias.detach().double().requires_grad_(True) for l in mlp.mlp]
    h = xd
    for w, b, m in zip(ws, bs, masks):
        pre = h @ w.t() + b
        bound = h.detach().abs() @ w.detach().abs().t() + b.detach().abs()
        clear = pre.detach().abs() > 1e-5 * bound
        assert torch.equal((pre.detach() > 0)[clear], m[clear]), "ReLU decision differs where the pre-activation is not ~0"
        h = pre * m
    h.backward(gy.double())
    return h.detach(), xd.grad, [w.grad for w in ws], [b.grad for b in bs]


@pytest.mark.parametrize("B,dims", [(1024, (429, 400, 400, 200)), (300, (77, 130, 40)), (2048, (64, 256))])
def test_fused_mlp_matches_fp64_and_the_per_layer_path(B, dims, monkeypatch):
    from pytorchrec_b200.model.layer import dense
    monkeypatch.setattr(dense, "TC_MIN_MACS", 0)
    torch.manual_seed(5)
    mlp = dense.MLP(dims[0], list(dims[1:]), "relu", 0.0).to(DEV)
    mlp._keep_masks = True
    layer_out = {}
    for i, d in enumerate(mlp.mlp):   # the per-layer path (step 0): its decisions are its outputs' signs
        d.register_forward_hook(lambda mod, inp, out, i=i: layer_out.__setitem__(i, out.detach() > 0))
    lib = _lib.load()
    L = len(dims) - 1
    for step in range(3):   # step 0 measures the scales on the per-layer path; steps 1, 2 run fused
        x = (torch.rand(B, dims[0], device=DEV) * (1.0 + step)).requires_grad_(True)
        gy = 1e-5 * torch.randn(B, dims[-1], device=DEV)
        mlp.zero_grad()
        layer_out.clear()
        n0 = lib.ptrec_launch_count()
        y = mlp(x)
        n_fwd = lib.ptrec_launch_count() - n0
        y.backward(gy)
        n_all = lib.ptrec_launch_count() - n0
        if step > 0:
            # forward: roll + split(x) + L x (split(W) + GEMM); backward: split(g) + colsum, (L - 1) x (GEMM + colsum),
            # dx GEMM, L x (wgrad GEMM [+ split-K reduce])
            assert not layer_out and n_fwd == 2 + 2 * L, n_fwd
            assert n_all - n_fwd <= 2 + 2 * (L - 1) + 1 + 2 * L, n_all - n_fwd
            masks = [ref_ops.unpack_mask_ref(m.cpu(), dims[1 + l]).to(DEV) for l, m in enumerate(mlp._last_masks)]
            masks.append(y.detach() > 0)
        else:
            masks = [layer_out[l] for l in range(L)]
        got = (y.detach(), x.grad, [l.linear.weight.grad.clone() for l in mlp.mlp],
               [l.linear.bias.grad.clone() for l in mlp.mlp])
        ref = _fp64_tower(mlp, x, gy, masks)
        flat_g = [got[0], got[1]] + got[2] + got[3]
        flat_r = [ref[0], ref[1]] + ref[2] + ref[3]
        for i, (a, r) in enumerate(zip(flat_g, flat_r)):
            tol = 3e-6 * r.abs().max().item()
            assert (a.double() - r).abs().max().item() <= tol, (step, i, (a.double() - r).abs().max().item(), tol)
    mlp.check_errors()
    assert mlp._scales.fwd_ready and mlp._scales.bwd_ready
    # inference under no_grad takes the fused path too and agrees with the per-layer path
    with torch.no_grad():
        x = torch.rand(B, dims[0], device=DEV)
        y_eval = mlp(x)
    monkeypatch.setenv("PTREC_TC_FUSED", "0")
    with torch.no_grad():
        y_layer = mlp(x)
    assert layer_out and (y_eval - y_layer).abs().max().item() <= 3e-6 * y_layer.abs().max().item()


def test_fused_mlp_flags_a_tensor_that_outgrows_its_scale(monkeypatch):
    from pytorchrec_b200.model.layer import dense
    monkeypatch.setattr(dense, "TC_MIN_MACS", 0)
    torch.manual_seed(1)
    mlp = dense.MLP(64, [128, 32], "relu", 0.0).to(DEV)
    x = torch.rand(256, 64, device=DEV)
    with torch.no_grad():
        mlp(x)                 # measures
        mlp(x)                 # fused
        mlp(x * 100.0)         # within the headroom
        mlp(x)
        mlp.check_errors()
        mlp(x * 1.0e4)         # planes overflow: flagged at the next roll
        mlp(x)
    with pytest.raises(RuntimeError, match="256-fold"):
        mlp.check_errors()
    mlp.reset_scales()
    with torch.no_grad():
        mlp(x)
    mlp.check_errors()


def test_tower_neighbours_hand_over_planes_and_deepfm_trains_identically(monkeypatch):
    """fm_head(..., tower=mlp) writes the tower input as planes, row_dot(..., tower_handoff=True) hands the output
    gradient back as planes: a DeepFM trained that way equals one trained with the hand-offs switched off (fp32
    tensors + split passes) up to the rounding of one extra fp32 store, and the launch count drops."""
    import numpy as np
    from pytorchrec_b200.feature_column import CategoricalColumnWithIdentity as Col
    from pytorchrec_b200.feature_column import NumericColumn
    from pytorchrec_b200.metric import LogLoss
    from pytorchrec_b200.model import DeepFM
    from pytorchrec_b200.model.layer import dense, interaction
    from pytorchrec_b200.optim import SparseSGD
    monkeypatch.setattr(dense, "TC_MIN_MACS", 0)
    F, nd, D, B = 6, 3, 16, 1024
    rows = [50 + 13 * f for f in range(F)]
    scols = [Col(rows[f], f"C{f}") for f in range(F)]
    dcols = [NumericColumn(f"I{j}", 0.0, 1.0, 0.5, 0.25) for j in range(nd)]
    lab = Col(2, "label")

    def run(handoff: bool):
        if not handoff:
            monkeypatch.setattr(dense.MLP, "tower_call", lambda self, x_like: None)
            real = interaction.row_dot
            monkeypatch.setattr(interaction, "row_dot", lambda h, w, tower_handoff=False: real(h, w, False))
        m = DeepFM(scols, dcols, lab, D, [64, 32], random_seed=7)
        m.compile(SparseSGD(m.get_parameters(), lr=0.2), torch.nn.BCEWithLogitsLoss(), [LogLoss()], DEV)
        rng = np.random.default_rng(3)
        lib = _lib.load()
        losses, launches = [], []
        for step in range(5):
            batch = {f"C{f}": torch.from_numpy(rng.integers(0, rows[f], size=B).astype(np.int32)) for f in range(F)}
            batch.update({f"I{j}": torch.from_numpy(rng.random(B).astype(np.float32)) for j in range(nd)})
            batch["label"] = torch.from_numpy(rng.integers(0, 2, size=B).astype(np.int32))
            n0 = lib.ptrec_launch_count()
            losses.append(m.train_step(batch)["loss"].item())
            launches.append(lib.ptrec_launch_count() - n0)
        m.mlp.check_errors()
        monkeypatch.undo()
        monkeypatch.setattr(dense, "TC_MIN_MACS", 0)
        return losses, launches, {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}

    l1, n1, sd1 = run(True)
    l0, n0, sd0 = run(False)
    assert n1[-1] <= n0[-1] - 2, (n1, n0)        # split(x) and split(g) are gone (the column sums moved, not vanished)
    for a, b in zip(l1, l0):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (l1, l0)
    for k in sd1:
        d = (sd1[k].double() - sd0[k].double()).abs().max().item()
        assert d <= 2e-6 * max(1.0, sd0[k].abs().max().item()), (k, d)
