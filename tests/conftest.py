import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Build (or find) the in-tree libraries once per session."""
    from compfinance_b200 import build
    build.build()
    build.build_host()
    return True


@pytest.fixture(scope="session")
def ref():
    """The reference itself, compiled by oracle/build_ref.py (prebuilt .so travels to the GPU box)."""
    from oracle import build_ref, refapi
    try:
        build_ref.build("/root/reference", verbose=False)
    except FileNotFoundError:
        pytest.skip("reference oracle library not available")
    r = refapi.get()
    r.start_pool(-1)
    return r


@pytest.fixture(scope="session")
def eng(built):
    from compfinance_b200 import capi
    return capi.Engine(device=0)


@pytest.fixture(scope="session")
def cf(built):
    from compfinance_b200.api import CompFinance
    return CompFinance(device=0)


# ---- canonical market data of the BASELINE configs ------------------------
def config3_surface():
    spots = np.arange(55, 201, 5.0)
    times = np.arange(1, 37) / 12.0
    vols = 0.15 + 0.10 * np.log(spots[:, None] / 100.0) ** 2 + 0.02 * times[None, :]
    return spots, times, vols


def put_config3(api, model_id="dup", product_id="uoc", max_dt=0.25):
    spots, times, vols = config3_surface()
    api.put_dupire(100.0, spots, times, vols, max_dt, model_id)
    api.put_barrier(120.0, 150.0, 3.0, 1.0 / 52, 0.01, False, product_id)
    return spots, times, vols


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
