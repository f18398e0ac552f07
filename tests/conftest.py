import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mpc-ekf4fastcharge_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _libmek_built():
    """The product library is built by __graft_entry__.build(); a fresh checkout that runs the tests first gets it built here
    (nvcc cross-compiles without a GPU).  Nothing is substituted if the build fails: the tests that need it then fail loudly."""
    lib = os.path.join(PKG, "mek", "libmek.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.call(["make", "-j8", "-C", os.path.join(PKG, "csrc"), "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return lib


@pytest.fixture(scope="session")
def bank():
    from mek.bank import Bank, default_bank_path
    return Bank.load(default_bank_path()).validate()


@pytest.fixture(scope="session")
def coracle_mod():
    from oracle import coracle
    coracle.build()
    return coracle


# signal scales for the relative-error criterion: |a-b| <= tol * max(|a|, scale)
SCALE = dict(u=60.0, V=4.0, SOC=1.0, phise=0.1)


def rel_err(a, b, scale):
    a = np.asarray(a, float); b = np.asarray(b, float)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "NaN patterns differ"
    m = ~nan_a
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(a[m]), scale)))
