import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu`")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_mf():
    return np.load(os.path.join(GOLDEN, "reference_mf.npz"))


@pytest.fixture(scope="session")
def golden_idioms():
    return np.load(os.path.join(GOLDEN, "reference_idioms.npz"))


@pytest.fixture(scope="session")
def golden_ncf():
    return np.load(os.path.join(GOLDEN, "reference_ncf.npz"))


@pytest.fixture(scope="session")
def golden_ctr():
    return np.load(os.path.join(GOLDEN, "oracle_ctr.npz"))


def state_from(npz, prefix):
    """{param name: tensor} for keys ``<prefix>/<name>``"""
    out = {}
    for k in npz.files:
        if k.startswith(prefix + "/"):
            out[k[len(prefix) + 1:]] = torch.from_numpy(np.array(npz[k]))
    return out


def batch_from(npz, prefix):
    return {k: torch.from_numpy(np.array(v)) for k, v in ((k[len(prefix) + 1:], npz[k]) for k in npz.files
                                                          if k.startswith(prefix + "/"))}


def assert_as_exact_as_the_oracle(name, a, b32, b64, rtol=1e-5, atol=1e-7, slack=3.0, max_slack=10.0):
    """fp32 parity where summation order decides the outcome.

    ``a`` = CUDA result, ``b32`` = the CPU oracle in fp32, ``b64`` = the same oracle run in fp64 ("exact").  Elements
    within ``rtol * |b32| + atol`` pass outright.  If some do not (Adagrad's g / (sqrt(sum g^2) + eps) amplifies
    reduction-order noise without bound where a row's duplicate gradients cancel), the CUDA result must be as close to
    the EXACT result as the reference's own fp32 arithmetic is: at the median and the 90th / 99th / 99.9th percentile
    |a - b64| <= slack * |b32 - b64| + the tolerance, and the same at the maximum with ``max_slack`` (the largest of
    millions of heavy-tailed errors is a noisy statistic: the worst CUDA element may be a few times off the worst
    oracle element without the distributions differing).  No fraction of elements is exempted."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b32 = np.asarray(b32, dtype=np.float64).ravel()
    b64 = np.asarray(b64, dtype=np.float64).ravel()
    tol = rtol * np.abs(b32) + atol
    if (np.abs(a - b32) <= tol).all():
        return
    e_p, e_r = np.abs(a - b64), np.abs(b32 - b64)
    floor = float(tol.max())
    for q in (0.5, 0.9, 0.99, 0.999, 1.0):
        qp, qr = np.quantile(e_p, q), np.quantile(e_r, q)
        k = max_slack if q == 1.0 else slack
        assert qp <= k * qr + floor, (f"{name}: error vs the fp64 oracle at quantile {q}: CUDA {qp:.3e}, "
                                      f"CPU fp32 oracle {qr:.3e} (slack {k}, floor {floor:.1e})")
