import importlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU check, skipped unless MIRO_SLOW=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("MIRO_SLOW") == "1":
        return
    skip = pytest.mark.skip(reason="set MIRO_SLOW=1 to run")
    for it in items:
        if "slow" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("cse168-raytracer_b200")


@pytest.fixture(scope="session")
def scenes():
    return importlib.import_module("cse168-raytracer_b200.scenes")


@pytest.fixture(scope="session")
def oracle():
    import miro_driver as md
    if not os.path.exists(md.ORACLE_SO):
        pytest.skip("oracle/libmiro_oracle.so not built (run __graft_entry__.build())")
    return md.oracle()


def _ref(kind):
    import miro_driver as md
    path = {"scalar": md.REF_SO, "stats": md.REF_STATS_SO}[kind]
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs /root/reference at build time)")
    return md.reference(kind)


@pytest.fixture(scope="session")
def reference():
    return _ref("scalar")


@pytest.fixture(scope="session")
def reference_stats():
    return _ref("stats")


@pytest.fixture(scope="session")
def emu():
    import ctypes
    path = os.path.join(HERE, "cpu_emu", "libmiro_emu.so")
    if not os.path.exists(path):
        pytest.skip("tests/cpu_emu/libmiro_emu.so not built")
    return ctypes.CDLL(path)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def subsample_rays(rays, w, h, step):
    """Every step-th pixel in x and y of a row-major w x h ray image."""
    r = rays.reshape(h, w, 8)[::step, ::step]
    return np.ascontiguousarray(r.reshape(-1, 8))


def random_rays(n, lo, hi, seed):
    """Incoherent rays: origins uniform in the box [lo,hi] (grown), directions uniform on the sphere."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(lo, np.float32); hi = np.asarray(hi, np.float32)
    c, e = (lo + hi) / 2, (hi - lo) / 2
    o = (c + (rng.random((n, 3), dtype=np.float32) * 2 - 1) * e * 1.5).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = o; rays[:, 4:7] = d; rays[:, 7] = 1e12
    return rays
