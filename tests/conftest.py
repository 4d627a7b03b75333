import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import mkfbodytracker_pdaf_b200 as mk
        return mk.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


class Arm:
    """one arm model in its three incarnations: product handle, oracle handle, numpy restatement"""

    def __init__(self, path, gamma_path):
        import mkf_oracle as orc
        import mkfbodytracker_pdaf_b200 as mk
        import np_ref
        self.mk = mk.Model.load(path, gamma_path)
        self.arrays = self.mk.arrays()
        a = self.arrays
        args = (a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"])
        self.orc = orc.Model(*args)
        self.np = np_ref.NpModel(*args)


@pytest.fixture(scope="session")
def left_arm():
    import mkfbodytracker_pdaf_b200 as mk
    # quirk B4 (src/pfPose.cpp:52-53): gamma of BOTH arms comes from the right-arm file
    return Arm(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)


@pytest.fixture(scope="session")
def right_arm():
    import mkfbodytracker_pdaf_b200 as mk
    return Arm(mk.RIGHT_ARM_MODEL, mk.RIGHT_ARM_MODEL)


@pytest.fixture
def rng():
    return np.random.default_rng(12345)
