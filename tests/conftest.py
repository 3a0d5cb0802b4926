import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(1, os.path.join(ROOT, "tests"))     # tests/parity.py, tests/plugin_templates.py
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~30 s on CPU")


class Golden(object):
    """Lazy access to the fixtures written by tests/golden/make_golden.py."""

    def __init__(self):
        with open(os.path.join(GOLDEN, "golden_meta.json")) as f:
            self.meta = json.load(f)
        self._cache = {}

    def npz(self, name):
        if name not in self._cache:
            self._cache[name] = np.load(os.path.join(GOLDEN, name))
        return self._cache[name]

    @property
    def synthetic_dem(self):
        return np.load(os.path.join(GOLDEN, "synthetic_dem_f32.npy")).astype(np.float64)

    @property
    def faultzone_dem(self):
        return self.npz("faultzone_dem_f32.npz")["z"].astype(np.float64)

    def seeded_dem(self, which):
        return self.npz("seeded_dems_f32.npz")[which].astype(np.float64)


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library on a CUDA device (gpu tests only)."""
    if os.environ.get("SB_DEBUG_EMU") == "1":
        # developer aid: run the parity tests against the CPU emulator build of the kernels
        from tests.emu.build_emu import build
        from scarplet_b200 import _lib
        lib = _lib.open_library(build())
        _lib._use_library(lib)
        return lib
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__
    __graft_entry__.build()
    from scarplet_b200 import _lib
    _lib._use_library(None)
    return _lib.load()


@pytest.fixture()
def emu_lib():
    """Routes scarplet_b200's host layer to the CPU emulator build of the SAME kernel
    source (tests/emu) for the duration of one test.  Test infrastructure only."""
    from tests.emu.build_emu import build
    from scarplet_b200 import _lib
    lib = _lib.open_library(build())
    prev = _lib._use_library(lib)
    yield lib
    _lib._use_library(prev)


def relerr(a, b, where):
    return np.abs(a - b)[where] / np.maximum(np.abs(b)[where], 1e-300)
