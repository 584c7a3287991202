import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without CUDA skips the gpu-marked tests instead of failing in them."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the C oracle and the CUDA library exist (both are git-ignored build products)."""
    import subprocess
    if not os.path.isfile(os.path.join(REPO, "oracle", "libdwt_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")])
    if not os.path.isfile(os.path.join(REPO, "wavedm_b200", "libwavedm_b200.so")):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "wavedm_b200", "csrc"), "-j8"])
    yield


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name))
