"""K6 in its fp16 x 2 operand format (``PTREC_TC_MODE=fp16x2``, include/ptrec_b200.h): device planes bit for bit against
the restatement in oracle/ref_ops.py, GEMM error against fp64 at the level of an fp32 GEMM, and a Dense layer against
``nn.Linear``.  fp16x2 is the default operand format (``PTREC_TC_MODE=bf16x3`` selects the three-plane one, whose kernel tests in
tests/test_gpu_kernels.py call it directly and run either way); the model parity tests of tests/test_gpu_models.py
exercise the default format end to end."""
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


@pytest.fixture(params=["2sm", "2sm_db", "1sm"])
def tc_kernel(request):
    """CTA-pair kernel with one 256-column accumulator pair, with two 128-column pairs (double-buffered), single CTA."""
    lib = _lib.load()
    bn = lib.ptrec_tc_get_bn()
    lib.ptrec_tc_set_2sm(0 if request.param == "1sm" else 1)
    lib.ptrec_tc_set_bn(128 if request.param == "2sm_db" else 256)
    yield request.param
    lib.ptrec_tc_set_2sm(1)
    lib.ptrec_tc_set_bn(bn)


@pytest.mark.parametrize("R,C", [(64, 64), (200, 429), (1031, 13), (4096, 400)])
@pytest.mark.parametrize("mag", [1.0, 1e-8, 3e5])
def test_split2h_planes_equal_the_restatement(R, C, mag):
    gen = torch.Generator().manual_seed(R + C)
    x = (mag * torch.randn(R, C, generator=gen) * torch.exp(3 * torch.randn(R, C, generator=gen)))
    y = torch.randn(R, C, generator=gen)
    h0, h1, s = ref_ops.split2h_ref(x)
    pl, plt, cs, sc = ops.tc_split2h(x.to(DEV), want_planes=True, want_t=True, want_colsum=True)
    assert sc.item() == s
    assert torch.equal(pl[0, :, :C].cpu(), h0) and torch.equal(pl[1, :, :C].cpu(), h1)
    assert torch.equal(plt[0, :, :R].cpu(), h0.t()) and torch.equal(plt[1, :, :R].cpu(), h1.t())
    assert (pl[:, :, C:] == 0).all() and (plt[:, :, R:] == 0).all()
    assert torch.allclose(cs.double().cpu(), x.double().sum(0), rtol=1e-5, atol=1e-5 * x.abs().sum(0).max().item())
    # fused ReLU backward: planes of g * (y > 0), scale from the unmasked tensor
    m0, m1, s2 = ref_ops.split2h_ref(x, mask_ref=y)
    pl2, _, cs2, sc2 = ops.tc_split2h(x.to(DEV), relu_ref=y.to(DEV), want_colsum=True)
    assert sc2.item() == s2 == s
    assert torch.equal(pl2[0, :, :C].cpu(), m0) and torch.equal(pl2[1, :, :C].cpu(), m1)
    masked = (x * (y > 0)).double()
    assert torch.allclose(cs2.double().cpu(), masked.sum(0), rtol=1e-5, atol=1e-5 * x.abs().sum(0).max().item())
    # strided, 16-byte-misaligned source (a column slice)
    wide = (mag * torch.randn(R, C + 5, generator=gen)).to(DEV)
    v0, v1, s3 = ref_ops.split2h_ref(wide[:, 3:3 + C].cpu())
    pl3, _, _, sc3 = ops.tc_split2h(wide[:, 3:3 + C])
    assert sc3.item() == s3 and torch.equal(pl3[0, :, :C].cpu(), v0) and torch.equal(pl3[1, :, :C].cpu(), v1)


def test_split2h_of_zeros_and_single_element():
    pl, _, _, sc = ops.tc_split2h(torch.zeros(5, 9, device=DEV))
    assert sc.item() == 1.0 and not pl.any()
    pl, _, _, sc = ops.tc_split2h(torch.full((1, 1), -3.0, device=DEV))
    assert sc.item() == 2.0 ** 12 and pl[0, 0, 0].item() == -3.0 * 2 ** 12 and pl[1, 0, 0].item() == 0.0


def _err(out, ref, scale):
    return ((out.double() - ref).abs() / scale).max().item()


def _chain(k_steps):
    """The tensor core aligns each MMA's products with the running fp32 sum and TRUNCATES: a chain of
    ``k_steps`` accumulations of 16 products carries up to half an ulp of bias per step relative to the running sum.
    Invisible when the sum cancels (|sum| << sum |a||b|, the usual case), visible on heavy-tailed operands whose sum is
    dominated by a few terms (measured: 7.1e-7 at K = 700 against 1.3e-7 for the CPU emulation of the same format)."""
    return k_steps * 2.0 ** -25


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 400, 429), (1000, 16, 40), (2048, 429, 400), (130, 1, 700),
                                   (4096, 400, 1677)])
@pytest.mark.parametrize("kind", ["unit", "tower", "tiny"])
def test_gemm_split2h_has_fp32_level_error(M, N, K, kind, tc_kernel):
    gen = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=gen)
    b = torch.randn(N, K, generator=gen)
    if kind == "tower":
        a, b = a.abs(), 0.01 * b
    elif kind == "tiny":          # gradient-sized operand: below fp16's range without the per-tensor scale
        a = 1e-8 * a * torch.exp(2 * torch.randn(M, K, generator=gen))
    a, b = a.to(DEV), b.to(DEV)
    bias = (b.abs().mean() * K ** 0.5 * a.abs().mean() * torch.randn(N, generator=gen).to(DEV)).contiguous()
    pa, _, _, sa = ops.tc_split2h(a)
    pb, _, _, sb = ops.tc_split2h(b)
    ref = a.double() @ b.double().t()
    scale = a.double().abs() @ b.double().abs().t()
    fp32_err = _err(a @ b.t(), ref, scale)
    out = ops.tc_gemm_split2h(pa, sa, pb, sb, K)
    assert out.shape == (M, N)
    tol = max(3e-7, 2 * fp32_err) + (_chain(K / 16) if kind == "tiny" else 0.0)
    assert _err(out, ref, scale) <= tol, (_err(out, ref, scale), fp32_err)
    # the device result equals the CPU emulation of the same arithmetic up to accumulation order
    emu = ref_ops.gemm_split2h_ref(a.cpu(), b.cpu())
    assert _err(out.cpu(), emu.double(), scale.cpu()) <= 2 * tol   # two fp32 accumulations
    out2 = ops.tc_gemm_split2h(pa, sa, pb, sb, K, bias=bias, relu=True)
    ref2 = torch.relu(ref + bias.double())
    assert _err(out2, ref2, scale + bias.abs().double()) <= tol
    out3 = ops.tc_gemm_split2h(pa, sa, pb, sb, K, splits=3)
    assert _err(out3, ref, scale) <= tol + 1e-7
    # the epilogue's max |out| word: exact, and a split that takes it produces the same planes as one that looks itself
    for kw in (dict(), dict(bias=bias, relu=True)):
        o, am = ops.tc_gemm_split2h(pa, sa, pb, sb, K, want_absmax=True, **kw)
        assert torch.equal(o, ops.tc_gemm_split2h(pa, sa, pb, sb, K, **kw))
        assert am.item() == o.abs().max().item()
        p1, _, _, s1 = ops.tc_split2h(o)
        p2, _, _, s2 = ops.tc_split2h(o, absmax_in=am)
        assert torch.equal(s1, s2) and torch.equal(p1, p2)


@pytest.mark.parametrize("B,N,K", [(4096, 400, 429), (1000, 128, 64), (777, 16, 1030), (64, 200, 13)])
def test_gemm_split2h_tn_weight_gradient_from_row_major_planes(B, N, K, tc_kernel):
    gen = torch.Generator().manual_seed(B + N)
    g = (1e-7 * torch.randn(B, N, generator=gen) * torch.exp(torch.randn(B, N, generator=gen))).to(DEV)
    x = torch.randn(B, K, generator=gen).to(DEV)
    pg, _, _, sg = ops.tc_split2h(g)
    px, _, _, sx = ops.tc_split2h(x)
    ref = g.double().t() @ x.double()
    scale = g.double().abs().t() @ x.double().abs()
    fp32_err = _err(g.t() @ x, ref, scale)
    auto = _lib.load().ptrec_tc_gemm_split3_default_splits(N, K, B)
    for splits, tol in ((0, max(4e-7, 2 * fp32_err) + _chain(B / 16 / auto)), (1, 1e-6 + _chain(B / 16))):
        dw = ops.tc_gemm_split2h_tn(pg, sg, N, px, sx, K, splits=splits)
        assert dw.shape == (N, K)
        assert _err(dw, ref, scale) <= tol, (splits, _err(dw, ref, scale), fp32_err)


def test_mlp_carries_absmax_words_between_layers_and_matches_fp64(monkeypatch):
    """Three Dense layers on the per-layer path (PTREC_TC_FUSED=0; the fused tower has tests/test_gpu_tc_fused.py):
    outputs / input gradients travel with the word holding their maximum (no extra pass), and the whole tower agrees
    with an fp64 evaluation."""
    from pytorchrec_b200.model.layer import dense
    monkeypatch.setenv("PTREC_TC_FUSED", "0")
    old, dense.TC_MIN_MACS = dense.TC_MIN_MACS, 0
    try:
        torch.manual_seed(5)
        mlp = dense.MLP(429, [400, 400, 200], "relu", 0.0).to(DEV)
        x = torch.rand(1024, 429, device=DEV).requires_grad_(True)
        h = mlp.mlp[0](x)
        assert dense._carried_absmax(h) is not None and dense._carried_absmax(h).item() == h.abs().max().item()
        assert dense._carried_absmax(h + 0) is None
        lib = _lib.load()
        n0 = lib.ptrec_launch_count()
        y = mlp(x)
        fwd_launches = lib.ptrec_launch_count() - n0
        # per layer: split(x) [+ absmax for layer 0 only], absmax(W) + split(W), GEMM  ->  3 * 4 + 1
        assert fwd_launches == 13, fwd_launches
        gy = 1e-5 * torch.randn_like(y)
        y.backward(gy)
        xd = x.detach().double().requires_grad_(True)
        hd = xd
        for layer in mlp.mlp:
            hd = torch.relu(hd @ layer.linear.weight.detach().double().t() + layer.linear.bias.detach().double())
        hd.backward(gy.double())
        assert (y.double() - hd).abs().max().item() <= 3e-6 * hd.abs().max().item()
        assert (x.grad.double() - xd.grad).abs().max().item() <= 3e-6 * xd.grad.abs().max().item()
    finally:
        dense.TC_MIN_MACS = old


def test_dense_layer_matches_nn_linear_forward_and_backward():
    from pytorchrec_b200.model.layer import dense
    old, dense.TC_MIN_MACS = dense.TC_MIN_MACS, 0
    try:
        torch.manual_seed(3)
        layer = dense.Dense(429, 400, "relu", 0.0).to(DEV)
        x = (torch.rand(2048, 429, device=DEV)).requires_grad_(True)
        gy = 1e-6 * torch.randn(2048, 400, device=DEV)
        before = _lib.load().ptrec_launch_count()
        y = layer(x)
        y.backward(gy)
        assert _lib.load().ptrec_launch_count() > before
        got = (y.detach(), x.grad, layer.linear.weight.grad, layer.linear.bias.grad)
        xd = x.detach().double().requires_grad_(True)
        w = layer.linear.weight.detach().double().requires_grad_(True)
        b = layer.linear.bias.detach().double().requires_grad_(True)
        yd = torch.relu(xd @ w.t() + b)
        yd.backward(gy.double())
        for name, a, r in zip(("y", "dx", "dW", "db"), got, (yd.detach(), xd.grad, w.grad, b.grad)):
            tol = 3e-6 * r.abs().max().item()   # 1e-5 relative asked; sums of ~400-2000 terms
            assert (a.double() - r).abs().max().item() <= tol, (name, (a.double() - r).abs().max().item(), tol)
    finally:
        dense.TC_MIN_MACS = old


def test_both_pair_tile_widths_give_identical_results():
    lib = _lib.load()
    bn = lib.ptrec_tc_get_bn()
    gen = torch.Generator().manual_seed(21)
    a = torch.randn(1000, 429, generator=gen).to(DEV)
    b = torch.randn(400, 429, generator=gen).to(DEV)
    g = torch.randn(1000, 400, generator=gen).to(DEV)
    bias = torch.randn(400, generator=gen).to(DEV)
    pa, _, _, sa = ops.tc_split2h(a)
    pb, _, _, sb = ops.tc_split2h(b)
    pg, _, _, sg = ops.tc_split2h(g)
    outs = []
    try:
        for width in (256, 128):
            lib.ptrec_tc_set_bn(width)
            outs.append((ops.tc_gemm_split2h(pa, sa, pb, sb, 429, bias=bias, relu=True),
                         ops.tc_gemm_split2h_tn(pg, sg, 400, pa, sa, 429, splits=4)))
    finally:
        lib.ptrec_tc_set_bn(bn)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
