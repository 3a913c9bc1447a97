"""pytest configuration: `gpu` marker + shared fixtures.

`-m "not gpu"`: oracle vs golden vectors / reference, host logic, C-ABI symbol export.
`-m gpu`     : parity tests proper -- CUDA path through the C-ABI vs the CPU oracle.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "huffman-codec_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    import pyoracle
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref not built (no reference sources on this box)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def samples():
    z = np.load(os.path.join(GOLDEN_DIR, "samples.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def hc():
    """The product: ctypes binding over the C-ABI of libhc_b200.so (fails loudly if missing)."""
    import hc_b200
    return hc_b200
