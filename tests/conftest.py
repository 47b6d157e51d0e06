import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_ROOT = os.environ.get("TCL_REFERENCE_ROOT", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # -m gpu tests must not silently pass on a box without a GPU
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_npz(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def tcl():
    import tcl_b200
    tcl_b200._cabi.build()
    return tcl_b200


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def reference():
    """The reference's own modules, imported by path (build container only)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "utils")):
        pytest.skip("reference checkout not present (GPU box)")
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self  # flowtools.py:25 hard-codes .cuda()
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "utils"))
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "methods", "learning-based"))
    import flowtools
    import fs_lib
    return flowtools, fs_lib
