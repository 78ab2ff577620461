import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The CUDA library and the C oracle are built in-tree (git-ignored); build them if absent."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "platymatch_b200", "libplatymatch_b200.so")):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "platymatch_b200", "csrc")])
    if not os.path.exists(os.path.join(ROOT, "oracle", "libpm_oracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


@pytest.fixture(scope="session")
def O():
    import oracle
    return oracle


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session", params=["asset02", "asset04", "synth400"])
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    return g


HYP_TAGS = ["11", "12", "13", "14", "21", "22", "23", "24"]
UNARY_KEYS = {"11": ("u11", "u21"), "12": ("u11", "u22"), "13": ("u11", "u23"), "14": ("u11", "u24"),
              "21": ("u12", "u21"), "22": ("u12", "u22"), "23": ("u12", "u23"), "24": ("u12", "u24")}
