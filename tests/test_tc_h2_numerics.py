"""Numerics of the fp16 x 2 operand format of K6 (include/ptrec_b200.h), emulated on the CPU: the representation
and the three-product formula keep fp32-level error on the magnitudes the DNN tower sees (activations O(1),
N(0, 0.01) weights, gradients of 1e-8 with heavy tails) — the property the GPU tests then check on the device."""
import numpy as np
import pytest
import torch

from oracle import ref_ops


def rel_err(out, a, b):
    ref = a.double() @ b.double().t()
    scale = a.double().abs() @ b.double().abs().t()
    return ((out.double() - ref).abs() / scale.clamp_min(1e-300)).max().item()


@pytest.mark.parametrize("case", ["unit", "tower_forward", "tiny_gradients", "wide_dynamic_range"])
def test_fp16x2_product_has_fp32_level_error(case):
    g = torch.Generator().manual_seed(3)
    M, N, K = 256, 200, 429
    if case == "unit":
        a, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    elif case == "tower_forward":
        a, b = torch.rand(M, K, generator=g), 0.01 * torch.randn(N, K, generator=g)
    elif case == "tiny_gradients":       # below fp16's subnormal range without the scale
        a = 1e-8 * torch.randn(M, K, generator=g) * torch.exp(2 * torch.randn(M, K, generator=g))
        b = 0.01 * torch.randn(N, K, generator=g)
    else:
        a = torch.randn(M, K, generator=g) * torch.exp(4 * torch.randn(M, K, generator=g))
        b = torch.randn(N, K, generator=g) * torch.exp(4 * torch.randn(N, K, generator=g))
    err = rel_err(ref_ops.gemm_split2h_ref(a, b), a, b)
    fp32 = rel_err(a @ b.t(), a, b)
    assert err <= max(3e-7, 2 * fp32), (case, err, fp32)


def test_split2h_representation_error_and_range():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(512, 300, generator=g) * torch.exp(3 * torch.randn(512, 300, generator=g))
    for mag in (1.0, 1e-9, 1e12):
        h0, h1, s = ref_ops.split2h_ref(x * mag)
        assert np.log2(s) == int(np.log2(s))                                  # a power of two
        top = float((x * mag).abs().max()) * s
        assert 2.0 ** 13 <= top < 2.0 ** 14
        assert torch.isfinite(h0.float()).all() and torch.isfinite(h1.float()).all()
        back = (h0.double() + h1.double() / 2048.0) / s
        err = (back - (x * mag).double()).abs()
        # 2^-22 relative for elements within 2^-28 of the maximum, an absolute floor of 2^-36 / s below that
        assert (err <= 2.0 ** -22 * (x * mag).abs().double() + 2.0 ** -36 / s).all()
    h0, h1, s = ref_ops.split2h_ref(torch.zeros(4, 4))
    assert s == 1.0 and not h0.any() and not h1.any()
    # ReLU mask: the scale comes from the unmasked tensor, masked entries are exactly zero in both planes
    y = torch.randn(512, 300, generator=g)
    h0, h1, s2 = ref_ops.split2h_ref(x, mask_ref=y)
    assert s2 == ref_ops.split2h_ref(x)[2]
    assert not h0[y <= 0].any() and not h1[y <= 0].any()
