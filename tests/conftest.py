import os
import sys

# several engines of one process share a GPU in the sharded-store tests: give every stream its own hardware queue
# (must be in the environment before the CUDA context exists)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_usable():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """a plain `pytest` on a box without a usable GPU skips the gpu-marked tests instead of failing in dfb_create"""
    if _cuda_usable():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def rcv1():
    """the reference's 100-row fixture (tests/data) as parsed by the reference reader"""
    d = np.load(os.path.join(GOLDEN, "rcv1_100.npz"))
    return dict(offset=d["offset"], label=d["label"], index=d["index"], value=d["value"])


@pytest.fixture(scope="session")
def refout():
    """outputs of the compiled reference (tests/golden/make_golden.py)"""
    return dict(np.load(os.path.join(GOLDEN, "ref_outputs.npz")))


def parse_kwargs(arr):
    out = {}
    for s in arr:
        k, v = str(s).split("=")
        out[k] = int(v) if k in ("V_dim", "V_threshold", "seed") else float(v)
    return out


def syn_batches(refout):
    bs = []
    b = 0
    while f"syn{b}_offset" in refout:
        bs.append((refout[f"syn{b}_offset"], refout[f"syn{b}_label"], refout[f"syn{b}_index"],
                   refout.get(f"syn{b}_value")))
        b += 1
    return bs
