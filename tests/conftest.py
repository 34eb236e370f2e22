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
