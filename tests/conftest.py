import json
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


def load_manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def load_golden(prefix, name):
    return np.load(os.path.join(GOLDEN, f"{prefix}_{name}.npz"))


def circ_dist_u8(a, b):
    """circular distance mod 256 (depth outputs wrap, quirk Q1)."""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return np.minimum(d, 256 - d)


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.lib()
    return orc
