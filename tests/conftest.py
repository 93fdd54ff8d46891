"""pytest configuration: the `gpu` marker, import paths and shared fixtures.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI load/export checks, gloo world_size-2 tests.
`-m gpu`       : parity tests proper, through the C ABI on a real B200 (nothing here reads /root/reference).
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = os.environ.get("RLG_REFERENCE", "/root/reference")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _load_by_path(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def ref_losses():
    """The real reference utils/losses.py, only where /root/reference is mounted (build container)."""
    if not os.path.exists(os.path.join(REFERENCE, "utils", "losses.py")):
        pytest.skip("reference checkout not present")
    return _load_by_path("utils/losses.py", "ref_losses")


@pytest.fixture(scope="session")
def ref_autoencoder():
    if not os.path.exists(os.path.join(REFERENCE, "models", "autoencoder.py")):
        pytest.skip("reference checkout not present")
    return _load_by_path("models/autoencoder.py", "ref_autoencoder")


@pytest.fixture(scope="session")
def golden_chamfer():
    return np.load(os.path.join(GOLDEN, "chamfer_ref.npz"))


@pytest.fixture(scope="session")
def golden_encoder():
    return np.load(os.path.join(GOLDEN, "encoder_ref.npz"))


@pytest.fixture(scope="session")
def golden_reward():
    return np.load(os.path.join(GOLDEN, "reward_ref.npz"))


@pytest.fixture(scope="session")
def rlg():
    import gan_rl_3d_b200
    return gan_rl_3d_b200
