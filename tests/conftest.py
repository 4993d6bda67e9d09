import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
